"""Debug aid: chunked streaming variants vs single chunk (sums + stats)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import synthgen as synth
from vstrains_b200 import pe_inference
cfg = synth.CONFIGS["C1"]
g, f, r = synth.generate(cfg, pairs=9000)
ids, seqs = pe_inference.parse_gfa_nodes(g.to_gfa())
for chunk_mb, two_pass, scan_mode in ((256, 0, 0), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 0, 3), (2, 0, 0), (256, 0, 1)):
    with pe_inference.PEIndex(seqs, cfg.k) as ix:
        ix.set_option("chunk_mb", chunk_mb)
        ix.set_option("scan_two_pass", two_pass)
        ix.set_option("scan_mode", scan_mode)
        ix.count_host(f, r)
        node, short = ix.matrices()
        st = ix.stats()
        buf = np.zeros(24, dtype=np.uint64)
        ix.set_option("dbg_counters", buf.ctypes.data)
        print(chunk_mb, two_pass, scan_mode, int(node.sum()), int(short.sum()),
              {k: st[k] for k in ("total_pairs", "n_pairs", "short_pairs", "used_pairs", "n_keys")}, list(map(int, buf)))
