"""Record-aligned sharding of a FASTQ pair (multi-GPU partitioning).

Pairing is by record index (reference utils/VStrains_PE_Inference.py:154-159), so both files
are cut at the SAME record numbers.  Cut points are found by counting line terminators with
the universal-newline rule; this is host-side plumbing (numpy), not part of the timed path."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def line_ends(buf: np.ndarray) -> np.ndarray:
    """Byte offsets one past every line terminator ('\\n', lone '\\r'; '\\r\\n' counts once)."""
    b = np.asarray(buf, dtype=np.uint8)
    nl = b == 10
    cr = b == 13
    if cr.any():
        nxt_nl = np.zeros(b.size, dtype=bool)
        nxt_nl[:-1] = nl[1:]
        term = nl | (cr & ~nxt_nl)
    else:
        term = nl
    return np.nonzero(term)[0] + 1


def n_records(buf: np.ndarray) -> int:
    b = np.asarray(buf, dtype=np.uint8)
    if b.size == 0:
        return 0
    ends = line_ends(b)
    lines = ends.size + (0 if ends.size and ends[-1] == b.size else 1)
    return lines // 4


def shard_ranges(fwd: np.ndarray, rve: np.ndarray, n_shards: int) -> List[Tuple[int, int, int, int]]:
    """[(fwd_lo, fwd_hi, rve_lo, rve_hi)] byte ranges holding the same record ranges."""
    ef, er = line_ends(fwd), line_ends(rve)
    total = min(n_records(fwd), n_records(rve))
    out = []
    for s in range(n_shards):
        a, b = total * s // n_shards, total * (s + 1) // n_shards
        lo_f = 0 if a == 0 else int(ef[4 * a - 1])
        lo_r = 0 if a == 0 else int(er[4 * a - 1])
        hi_f = lo_f if b == a else int(ef[4 * b - 1]) if 4 * b - 1 < ef.size else int(np.asarray(fwd).size)
        hi_r = lo_r if b == a else int(er[4 * b - 1]) if 4 * b - 1 < er.size else int(np.asarray(rve).size)
        out.append((lo_f, hi_f, lo_r, hi_r))
    return out
