"""Profiling aid: per-tier read counts of the map stage (device counters after one call).
usage: python tools/dbg_map.py [opt=value,...|-] [config] [pairs]"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
from vstrains_b200 import pe_inference
opts = dict(kv.split("=") for kv in sys.argv[1].split(",")) if len(sys.argv) > 1 and "=" in sys.argv[1] else {}
cfgn = sys.argv[2] if len(sys.argv) > 2 else "C2"
pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 500000
cfg, g, f, r = bench.make_workload(cfgn, pairs, 0)
ix = pe_inference.PEIndex([bytes(s) for s in g.seqs], cfg.k)
[ix.set_option(k, int(v)) for k, v in opts.items()]
d_f = torch.from_numpy(f).cuda(); d_r = torch.from_numpy(r).cuda()
for _ in range(3):
    ix.reset(); ix.count_device(d_f.data_ptr(), f.size, d_r.data_ptr(), r.size)
names = ["TOTAL", "N", "SHORT", "USED", "KEYS", "SPILL_CURSOR", "ERR", "FAST", "GENERIC", "WORK", "BAILED", "DEFER", "WORK2", "DEFER1", "-", "LISTS", "OVF", "PAIR_OCC", "EXP", "EXP_CURSOR",
         "B_USED", "B_N", "B_SHORT", "WALK", "WALK1", "MEMO_HIT", "PAIR_LIST"]
buf = np.zeros(len(names), dtype=np.uint64)
ix.set_option("dbg_counters", buf.ctypes.data)
print("%s %s pairs=%d (per-mate counters are those of the second mate)" % (opts, cfgn, pairs))
for n, v in zip(names, buf):
    print("  %-12s %d" % (n, v))
st = ix.stats()
print({k: st[k] for k in ("ms_scan", "ms_map", "ms_count", "ms_total")})
