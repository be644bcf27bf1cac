"""Launched by tests/test_gpu_parity.py under torchrun (one process per GPU, NCCL): every rank runs
vstrains_b200.dist.run_rank on its record-aligned shard, the matrices are merged with ONE NCCL
allreduce in device memory, and every rank checks the merged result against the C oracle run on
the whole input (test infrastructure).  Exit status 0 = bit-exact on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29513 tests/nccl_rank_check.py [config] [pairs] [sparse]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import synthgen as synth
    from oracle import c_oracle
    from vstrains_b200 import dist as vdist
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = synth.CONFIGS[name]
    g, f, r = synth.generate(cfg, pairs=pairs)               # identical on every rank
    gfa = g.to_gfa()
    ids, node, short, counters = vdist.run_rank(gfa, f, r, cfg.k, rank, world, device=local)
    onode, oshort, ostats = c_oracle.run(gfa, f, r, cfg.k)
    ok = np.array_equal(node.astype(np.int64), onode) and np.array_equal(short.astype(np.int64), oshort) and counters == ostats
    flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print("rank %d/%d: %s (links %d, pairs %r)" % (rank, world, "bit-exact" if ok else "MISMATCH", int(node.sum()), counters), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
