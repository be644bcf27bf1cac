"""Small end-to-end run for compute-sanitizer (memcheck): golden fixtures + one synthetic case."""
import glob, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synthgen as synth
from vstrains_b200 import pe_inference
for p in sorted(glob.glob("tests/golden/*.npz"))[:8] + ["tests/golden/synth_2x250_k127.npz", "tests/golden/synth_2x150_k77.npz"]:
    z = np.load(p)
    if int(z["status"]) != 0:
        continue
    for opts in ({}, {"scan_mode": 1}, {"force_generic": 1, "scan_mode": 1}, {"sparse": 1}, {"subst": 0}, {"tier_overlap": 0}, {"memo": 0}):
        ids, seqs = pe_inference.parse_gfa_nodes(z["gfa"].tobytes())
        with pe_inference.PEIndex(seqs, int(z["k"])) as ix:
            for k, v in opts.items():
                ix.set_option(k, v)
            ix.count_host(z["fwd"].tobytes(), z["rve"].tobytes())
            if not opts:
                ix.count_host(z["fwd"].tobytes(), z["rve"].tobytes())      # second call: the read memo answers
            ix.sparse() if ix.is_sparse else ix.matrices()
cfg = synth.CONFIGS["C2"]
g, f, r = synth.generate(cfg, pairs=3000)
ids, node, short, st = pe_inference.pe_inference(g.to_gfa(), f, r, cfg.k)
print("sanitize run done", st["total_pairs"], int(node.sum()))
