"""CPU restatement of VStrains' paired-end link inference -- TEST INFRASTRUCTURE ONLY.

This file is the parity oracle for the CUDA path.  It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs;
nothing under ``vstrains_b200/`` may import it (the product path has no CPU fallback).

Parity status: PINNED.  The reference ships no golden vectors, so the oracle
is pinned against outputs of the reference script itself
(``/root/reference/utils/VStrains_PE_Inference.py``) generated in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/``; ``tests/test_oracle.py``
re-checks every fixture byte for byte, and (when ``/root/reference`` is present) re-runs the
reference live on fresh random cases.

Every function cites the reference lines it follows (paths relative to the reference root).
"""
from __future__ import annotations

import io
import os
import shutil
from typing import Dict, Iterable, List, Sequence, Tuple

_REV = {"A": "T", "T": "A", "C": "G", "G": "C"}


def reverse_seq(seq: str) -> str:
    """utils/VStrains_PE_Inference.py:9-13 -- reverse complement; KeyError on non-ACGT."""
    return "".join(_REV[c] for c in reversed(seq))


def split_lines(data: bytes) -> List[str]:
    """Text-mode ``readlines()`` as used at utils/VStrains_PE_Inference.py:105,149-150:
    universal newlines ('\\n', '\\r\\n' and lone '\\r' all end a line and are translated to
    '\\n'), a trailing unterminated line is still a line.  Non-UTF-8 input raises like the
    reference does under a UTF-8 locale."""
    return io.TextIOWrapper(io.BytesIO(data), encoding="utf-8", newline=None).readlines()


def parse_gfa(data: bytes) -> Tuple[List[str], List[str]]:
    """utils/VStrains_PE_Inference.py:101-112 -- every line, ``Line[:-1].split('\\t')``;
    keep lines whose first field is ``S``: id = field 1, seq = field 2."""
    ids: List[str] = []
    seqs: List[str] = []
    for line in split_lines(data):
        f = line[:-1].split("\t")
        if f[0] == "S":
            ids.append(f[1])
            seqs.append(f[2])
    return ids, seqs


def build_index(seqs: Sequence[str], split_len: int) -> Dict[str, List[Tuple[int, int]]]:
    """utils/VStrains_PE_Inference.py:117-135 -- (k+1)-mer -> postings [(node, pos)].
    Forward k-mer and its reverse complement both receive (node, pos); postings are a
    multiset (a reverse-palindromic k-mer gets the same entry twice)."""
    table: Dict[str, List[Tuple[int, int]]] = {}
    for i, seq in enumerate(seqs):
        for p in range(len(seq) - split_len + 1):
            kmer = seq[p:p + split_len]
            rev = reverse_seq(kmer)
            table.setdefault(kmer, []).append((i, p))
            table.setdefault(rev, []).append((i, p))
    return table


def map_read(seq: str, table: Dict[str, List[Tuple[int, int]]], seqlen: Sequence[int],
             split_len: int) -> List[int]:
    """utils/VStrains_PE_Inference.py:16-48 (single_end_read_mapping) with the float test
    ``v >= max(min(saturate, expected), 1)`` (:36-47) restated in integers:

        saturate = min(len, rlen - kmin) - split_len + 1            (:38-41; coords cancel)
        expected * rlen = (min(rlen, len) - split_len + 1) * (rlen - split_len)   (:42-44)
        keep  <=>  v >= saturate  or  v * rlen >= expected * rlen   (v >= 1 always holds)

    Returns ascending node indices, as the reference's ``enumerate(nodes)`` scan does."""
    rlen = len(seq)
    v: Dict[int, int] = {}
    kmin: Dict[int, int] = {}
    for i in range(rlen - split_len + 1):
        post = table.get(seq[i:i + split_len])
        if post:
            for rid, _ in post:
                if rid in v:
                    v[rid] += 1
                else:
                    v[rid] = 1
                    kmin[rid] = i
    out = []
    for rid in sorted(v):
        ln = seqlen[rid]
        sat = min(ln, rlen - kmin[rid]) - split_len + 1
        ab = (min(rlen, ln) - split_len + 1) * (rlen - split_len)
        if v[rid] >= sat or v[rid] * rlen >= ab:
            out.append(rid)
    return out


def count_pairs(fwd_lines: Sequence[str], rve_lines: Sequence[str], table, seqlen, split_len):
    """utils/VStrains_PE_Inference.py:154-188 -- records paired by index, seq = 2nd line of
    each 4-line group minus its last char; upper-case 'N' in either mate skips the pair
    (checked before the length test); either mate shorter than split_len skips it.
    Returns sparse ``node`` / ``short`` counters keyed (i, j) and the three pair counters."""
    node: Dict[Tuple[int, int], int] = {}
    short: Dict[Tuple[int, int], int] = {}
    n_reads = short_reads = used = 0
    total = min(len(fwd_lines) // 4, len(rve_lines) // 4)
    for r in range(total):
        fseq = fwd_lines[4 * r + 1][:-1]
        rseq = rve_lines[4 * r + 1][:-1]
        if "N" in fseq or "N" in rseq:
            n_reads += 1
        elif len(fseq) < split_len or len(rseq) < split_len:
            short_reads += 1
        else:
            used += 1
            lefts = map_read(fseq, table, seqlen, split_len)
            rights = map_read(rseq, table, seqlen, split_len)
            for lst in (lefts, rights):                       # :174-184
                for a in range(len(lst)):
                    for b in range(a, len(lst)):
                        key = (lst[a], lst[b])
                        short[key] = short.get(key, 0) + 1
            for i in lefts:                                   # :186-188
                for j in rights:
                    node[(i, j)] = node.get((i, j), 0) + 1
    return node, short, {"total_pairs": total, "n_pairs": n_reads,
                         "short_pairs": short_reads, "used_pairs": used}


def info_bytes(ids: Sequence[str], counts: Dict[Tuple[int, int], int]) -> bytes:
    """utils/VStrains_PE_Inference.py:194-207 -- dense row-major ``id_i:id_j:count\\n``."""
    n = len(ids)
    out = []
    for i in range(n):
        a = ids[i]
        for j in range(n):
            out.append("%s:%s:%d\n" % (a, ids[j], counts.get((i, j), 0)))
    return "".join(out).encode()


def run_bytes(gfa: bytes, fwd: bytes, rve: bytes, kmer_size: int):
    """Whole path on in-memory inputs -> (pe_info bytes, st_info bytes, stats, ids)."""
    ids, seqs = parse_gfa(gfa)
    split_len = kmer_size + 1                                  # :114
    table = build_index(seqs, split_len)
    seqlen = [len(s) for s in seqs]
    node, short, stats = count_pairs(split_lines(fwd), split_lines(rve), table, seqlen, split_len)
    return info_bytes(ids, node), info_bytes(ids, short), stats, ids


def run(gfa_path: str, fwd_path: str, rve_path: str, kmer_size: int, out_dir: str):
    """File-level equivalent of the reference ``main()`` (:51-211) incl. ``rm -rf DIR``."""
    if out_dir.endswith("/"):
        out_dir = out_dir[:-1]
    shutil.rmtree(out_dir, ignore_errors=True)                 # :93-96
    os.makedirs(out_dir, exist_ok=True)
    with open(gfa_path, "rb") as f:
        gfa = f.read()
    with open(fwd_path, "rb") as f:
        fwd = f.read()
    with open(rve_path, "rb") as f:
        rve = f.read()
    pe, st, stats, _ = run_bytes(gfa, fwd, rve, kmer_size)
    with open(os.path.join(out_dir, "pe_info"), "wb") as f:
        f.write(pe)
    with open(os.path.join(out_dir, "st_info"), "wb") as f:
        f.write(st)
    return stats


def dense(n: int, counts: Dict[Tuple[int, int], int]):
    import numpy as np
    m = np.zeros((n, n), dtype=np.int64)
    for (i, j), c in counts.items():
        m[i, j] = c
    return m
