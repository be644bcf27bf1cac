"""Debug aid: run the same block twice at several sizes, report cells that differ and the device counters."""
import sys
import numpy as np
sys.path.insert(0, '.')
import bench
from vstrains_b200 import pe_inference
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1_000_000, 4_000_000, 9_000_000]
opts = dict(kv.split("=") for kv in sys.argv[3].split(",")) if len(sys.argv) > 3 and "=" in sys.argv[3] else {}
cfg, g, genomes, ab = bench.make_graph(name, max(sizes))
f, r = bench.make_reads(cfg, genomes, ab, max(sizes), 0)
gfa = g.to_gfa()
ids, seqs = pe_inference.parse_gfa_nodes(gfa)
names = ["TOTAL", "N", "SHORT", "USED", "KEYS", "SPILL_CURSOR", "ERR", "FAST", "GENERIC", "WORK", "BAILED", "DEFER", "WORK2", "DEFER2", "BIG",
         "LISTS", "OVF", "PAIR_OCC", "EXP", "EXP_CURSOR"]
for n in sizes:
    fs, rs = bench.prefix_pairs(f, r, 0, n, cfg.read_len)
    with pe_inference.PEIndex(seqs, cfg.k) as ix:
        for k, v in opts.items():
            ix.set_option(k, int(v))
        res = []
        for rep in range(3):
            ix.reset()
            ix.count_host(fs, rs)
            node, short = ix.matrices()
            st = ix.stats()
            buf = np.zeros(len(names) + 4, dtype=np.uint64)
            ix.set_option("dbg_counters", buf.ctypes.data)
            res.append((node.copy(), short.copy()))
            print(n, rep, "sum", int(node.sum()) + int(short.sum()), "n_keys", st["n_keys"], {k: st[k] for k in ("total_pairs", "used_pairs")},
                  dict(zip(names, map(int, buf))), flush=True)
        for rep in (1, 2):
            dn = np.nonzero(res[0][0] != res[rep][0]); ds = np.nonzero(res[0][1] != res[rep][1])
            print("  run0 vs run%d: node cells differing %d, short cells differing %d" % (rep, dn[0].size, ds[0].size))
            if dn[0].size:
                i, j = dn[0][0], dn[1][0]
                print("   e.g. node[%d][%d] = %d vs %d" % (i, j, res[0][0][i, j], res[rep][0][i, j]))
            if ds[0].size:
                i, j = ds[0][0], ds[1][0]
                print("   e.g. short[%d][%d] = %d vs %d" % (i, j, res[0][1][i, j], res[rep][1][i, j]))
