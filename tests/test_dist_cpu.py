"""world_size-2 gloo tests of the N>1 host logic: record-aligned sharding + one allreduce.
The per-shard counts come from the oracle here (no GPU); the merged result must equal the
single-process result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle, pe_oracle
from vstrains_b200 import dist as vdist
import synthgen as synth
from vstrains_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_count(gfa, f, r, k):
    node, short, stats = c_oracle.run(gfa, f, r, k, 1)
    return node, short, stats


def _worker(rank, world, port, gfa, f, r, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids, node, short, counters = vdist.run_rank(gfa, f, r, k, rank, world, count_fn=_oracle_count)
        if rank == 0:
            np.savez(out, node=node, short=short, **{k_: np.int64(v) for k_, v in counters.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_counts_allreduce_to_the_single_process_result(tmp_path, world):
    cfg = synth.CONFIGS["C3"]
    g, f, r = synth.generate(cfg, pairs=1500)
    # make the two files disagree on record count and end without a newline
    r = r[: r.size - 700]
    f = f[:-1]
    gfa = g.to_gfa()
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(world, _free_port(), gfa, f, r, cfg.k, out), nprocs=world, join=True)
    z = np.load(out)
    node, short, stats = c_oracle.run(gfa, f, r, cfg.k, 1)
    assert np.array_equal(z["node"], node)
    assert np.array_equal(z["short"], short)
    for k in vdist.COUNTER_KEYS:
        assert int(z[k]) == stats[k]


def test_shard_ranges_cover_every_record_once():
    rng = np.random.default_rng(5)
    for nl in (b"\n", b"\r\n", b"\r"):
        recs_f = [b"@h%d" % i + nl + b"ACGT"[: rng.integers(1, 5)] * rng.integers(1, 9) + nl + b"+" + nl + b"II" + nl for i in range(37)]
        recs_r = [b"@h%d" % i + nl + b"TTGA" * rng.integers(1, 9) + nl + b"+" + nl + b"II" + nl for i in range(41)]
        f = np.frombuffer(b"".join(recs_f), dtype=np.uint8)
        r = np.frombuffer(b"".join(recs_r)[: -len(nl)], dtype=np.uint8)
        assert shard.n_records(f) == 37 and shard.n_records(r) == 41
        for world in (1, 2, 5, 40):
            ranges = shard.shard_ranges(f, r, world)
            assert ranges[0][0] == 0 and ranges[0][2] == 0
            tot = 0
            for i, (a, b, c, d) in enumerate(ranges):
                if i:
                    assert a == ranges[i - 1][1] and c == ranges[i - 1][3]
                nf, nr = shard.n_records(f[a:b]), shard.n_records(r[c:d])
                assert nf == nr
                tot += nf
            assert tot == 37
            assert pe_oracle.split_lines(f[: ranges[-1][1]].tobytes())[: 4 * 37] == pe_oracle.split_lines(f.tobytes())[: 4 * 37]
