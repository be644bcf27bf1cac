"""One-process-per-GPU execution of the path (torchrun / torch.distributed plumbing).

Read pairs are independent (reference utils/VStrains_PE_Inference.py:154-188), so rank r takes
the r-th record-aligned shard of both FASTQ files, counts it against the replicated index and
the per-rank ``[node_mat | short_mat]`` are summed with ONE allreduce (NCCL over NVLink on GPUs;
gloo in the CPU tests of the sharding/merge logic).  Integer sums make the result independent
of the world size.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import numpy as np

from . import shard

COUNTER_KEYS = ("total_pairs", "n_pairs", "short_pairs", "used_pairs")


def rank_shard(fwd: np.ndarray, rve: np.ndarray, rank: int, world: int) -> Tuple[np.ndarray, np.ndarray]:
    """The byte ranges of both files that hold record range ``rank`` of ``world``."""
    lo_f, hi_f, lo_r, hi_r = shard.shard_ranges(fwd, rve, world)[rank]
    return fwd[lo_f:hi_f], rve[lo_r:hi_r]


def merge_counts(mats, counters: Dict[str, int], group=None):
    """Sum ``mats`` (int64 tensor holding node_mat then short_mat, on the rank's device) and the
    pair counters over all ranks, in place.  One allreduce for the matrices, one tiny one for
    the counters."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mats, {k: int(counters[k]) for k in COUNTER_KEYS}
    dist.all_reduce(mats, op=dist.ReduceOp.SUM, group=group)
    c = torch.tensor([int(counters[k]) for k in COUNTER_KEYS], dtype=torch.int64, device=mats.device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return mats, dict(zip(COUNTER_KEYS, (int(x) for x in c.tolist())))


class _DeviceArray:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}


def run_rank(gfa: bytes, fwd: np.ndarray, rve: np.ndarray, kmer_size: int, rank: int, world: int,
             count_fn: Optional[Callable] = None, device: Optional[int] = None):
    """Whole path on this rank's shard + merge.  ``count_fn(gfa, f, r, k) -> (node, short, counters)``
    replaces the CUDA path in CPU-only tests (numpy int64 matrices); by default the shard goes
    through libvspe on ``device`` and the matrices are reduced in place in device memory.
    Returns (ids, node_mat, short_mat, counters) with the merged (global) values."""
    import torch
    from . import pe_inference
    f, r = rank_shard(fwd, rve, rank, world)
    ids, seqs = pe_inference.parse_gfa_nodes(gfa)
    n = len(ids)
    if count_fn is not None:
        node, short, counters = count_fn(gfa, f, r, kmer_size)
        mats = torch.from_numpy(np.concatenate([node.reshape(-1), short.reshape(-1)]).astype(np.int64))
        mats, counters = merge_counts(mats, counters)
        m = mats.numpy()
        return ids, m[: n * n].reshape(n, n), m[n * n:].reshape(n, n), counters
    dev = rank if device is None else device
    with pe_inference.PEIndex(seqs, kmer_size, device=dev) as ix:
        ix.count_host(f, r)
        st = ix.stats()
        ptr, cnt = ix.matrices_device()
        if cnt:
            mats = torch.as_tensor(_DeviceArray(ptr, cnt), device=torch.device("cuda", dev))
            _, counters = merge_counts(mats, st)
            torch.cuda.synchronize(dev)
        else:
            counters = {k: st[k] for k in COUNTER_KEYS}
        ix.set_pair_counters(*(counters[k] for k in COUNTER_KEYS))
        node, short = ix.matrices()
    return ids, node, short, counters
