"""Host-side mirror of the reference's paired-end link inference script.

Same names, flags, stdout lines, side effects and output files as reference
``utils/VStrains_PE_Inference.py`` (``main()`` :51-211); the compute goes through the C ABI of
``libvspe.so`` (CUDA, sm_100a).  There is no CPU path here: without the library or a B200 every
call raises / exits non-zero.

In-process API::

    ids, node_mat, short_mat, stats = pe_inference(gfa, fwd, rve, kmer_size)
"""
from __future__ import annotations

import argparse
import ctypes
import os
import sys
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import Stats, VspeError, check

rev_dict = {"A": "T", "T": "A", "C": "G", "G": "C"}


def reverse_seq(seq: str):
    """Same contract as reference utils/VStrains_PE_Inference.py:12-13 (KeyError on non-ACGT).
    Kept for API fidelity; the index build does this on the device."""
    return "".join(rev_dict[x] for x in reversed(seq))


def _u8(x) -> np.ndarray:
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x.view(np.uint8).reshape(-1))
    return np.frombuffer(bytes(x) if not isinstance(x, (bytes, bytearray, memoryview)) else x, dtype=np.uint8)


def parse_gfa_nodes(gfa: bytes) -> Tuple[List[str], List[bytes]]:
    """S lines of a GFA in file order -- reference :101-112 (text mode, ``Line[:-1].split('\\t')``)."""
    ids, seqs = [], []
    text = gfa.replace(b"\r\n", b"\n").replace(b"\r", b"\n")
    lines = text.split(b"\n")
    terminated = text.endswith(b"\n")
    if terminated:
        lines.pop()
    for n, line in enumerate(lines):
        if not terminated and n == len(lines) - 1:
            line = line[:-1]                      # `Line[:-1]` eats a real char of an unterminated last line
        f = line.split(b"\t")
        if f[0] == b"S":
            if len(f) < 3:
                raise VspeError(-4, "GFA S line with fewer than 3 fields")
            ids.append(f[1].decode("ascii"))
            seqs.append(f[2])
    return ids, seqs


class PEIndex:
    """Device-resident (k+1)-mer index + count matrices for one graph (one GPU)."""

    def __init__(self, seqs: Sequence[bytes], kmer_size: int, device: int = 0):
        self._L = _lib.lib()
        self._ctx = ctypes.c_void_p()
        check(self._L.vspe_create(device, ctypes.byref(self._ctx)))
        self.n_nodes = len(seqs)
        self.split_len = kmer_size + 1
        cat = np.frombuffer(b"".join(seqs), dtype=np.uint8)
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        if seqs:
            off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
        try:
            check(self._L.vspe_index_build(self._ctx, cat.ctypes.data if cat.size else None, off.ctypes.data,
                                           self.n_nodes, self.split_len))
        except Exception:
            self.close()
            raise

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._L.vspe_destroy(self._ctx)
            self._ctx = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, name: str, value: int):
        check(self._L.vspe_set_option(self._ctx, name.encode(), int(value)))

    def reset(self):
        check(self._L.vspe_reset(self._ctx))

    def count_host(self, fwd, rve):
        f, r = _u8(fwd), _u8(rve)
        check(self._L.vspe_count_host(self._ctx, f.ctypes.data if f.size else None, f.size,
                                      r.ctypes.data if r.size else None, r.size))

    def count_host_ptr(self, fptr: int, fn: int, rptr: int, rn: int):
        check(self._L.vspe_count_host(self._ctx, fptr, fn, rptr, rn))

    def count_device(self, fptr: int, fn: int, rptr: int, rn: int):
        """fptr/rptr: device pointers (e.g. torch ``tensor.data_ptr()``) on this context's GPU."""
        check(self._L.vspe_count_device(self._ctx, fptr, fn, rptr, rn))

    @property
    def is_sparse(self) -> bool:
        return bool(self._L.vspe_is_sparse(self._ctx))

    def sparse(self) -> Tuple[np.ndarray, np.ndarray]:
        """Sparse mode: (keys, counts), keys ascending, key = mat*N*N + i*N + j (mat 0 node_mat, 1 short_mat)."""
        n, pk, pc = ctypes.c_uint64(), ctypes.c_void_p(), ctypes.c_void_p()
        check(self._L.vspe_sparse_host(self._ctx, ctypes.byref(n), ctypes.byref(pk), ctypes.byref(pc)))
        if n.value == 0:
            return np.zeros(0, np.uint64), np.zeros(0, np.uint64)
        keys = np.ctypeslib.as_array(ctypes.cast(pk, ctypes.POINTER(ctypes.c_uint64)), (n.value,)).copy()
        counts = np.ctypeslib.as_array(ctypes.cast(pc, ctypes.POINTER(ctypes.c_uint64)), (n.value,)).copy()
        return keys, counts

    def stream(self) -> int:
        """The cudaStream_t (as an integer) this context launches on, e.g. for ``torch.cuda.ExternalStream``."""
        return self._L.vspe_stream(self._ctx) or 0

    def sparse_device(self) -> Tuple[int, int, int]:
        """Sparse mode: (n_runs, device pointer of the keys, device pointer of the counts)."""
        n, pk, pc = ctypes.c_uint64(), ctypes.c_void_p(), ctypes.c_void_p()
        check(self._L.vspe_sparse_device(self._ctx, ctypes.byref(n), ctypes.byref(pk), ctypes.byref(pc)))
        return n.value, pk.value or 0, pc.value or 0

    def sparse_merge_device(self, kptr: int, cptr: int, n: int):
        """Add runs that live in this context's device memory (exact integer sums)."""
        check(self._L.vspe_sparse_merge_device(self._ctx, kptr, cptr, n))

    def sparse_clear(self):
        """Forget the runs of this context (its pair counters stay)."""
        check(self._L.vspe_sparse_clear(self._ctx))

    def sparse_merge(self, keys: np.ndarray, counts: np.ndarray):
        """Add the runs of another context / rank (exact integer sums)."""
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        c = np.ascontiguousarray(counts, dtype=np.uint64)
        check(self._L.vspe_sparse_merge(self._ctx, k.ctypes.data if k.size else None, c.ctypes.data if c.size else None, k.size))

    def matrices(self, out: Optional[Tuple[np.ndarray, np.ndarray]] = None) -> Tuple[np.ndarray, np.ndarray]:
        """(node_mat, short_mat), uint64 [N, N].  ``out``: two C-contiguous uint64 [N, N] arrays to copy into -- e.g.
        views of pinned host memory (``vspe_alloc_pinned`` / a pinned torch tensor), which the device-to-host copy
        then reaches at PCIe speed instead of through the driver's staging of pageable memory."""
        n = self.n_nodes
        if out is not None and not self.is_sparse:
            node, short = out
            for a in (node, short):
                if a.dtype != np.uint64 or a.shape != (n, n) or not a.flags["C_CONTIGUOUS"]:
                    raise VspeError(-1, "matrices(out=...): need two C-contiguous uint64 arrays of shape (N, N)")
            check(self._L.vspe_matrices_host(self._ctx, node.ctypes.data, short.ctypes.data))
            return node, short
        if self.is_sparse:
            if n > 20000:
                raise VspeError(-1, "graph too large for dense matrices: use PEIndex.sparse()")
            keys, counts = self.sparse()
            flat = np.zeros(2 * n * n, dtype=np.uint64)
            flat[keys.astype(np.int64)] = counts
            return flat[: n * n].reshape(n, n), flat[n * n:].reshape(n, n)
        node = np.zeros((n, n), dtype=np.uint64)
        short = np.zeros((n, n), dtype=np.uint64)
        check(self._L.vspe_matrices_host(self._ctx, node.ctypes.data, short.ctypes.data))
        return node, short

    def matrices_device(self) -> Tuple[int, int]:
        p, n = ctypes.c_void_p(), ctypes.c_uint64()
        check(self._L.vspe_matrices_device(self._ctx, ctypes.byref(p), ctypes.byref(n)))
        return p.value or 0, n.value

    def stats(self) -> dict:
        s = Stats()
        check(self._L.vspe_get_stats(self._ctx, ctypes.byref(s)))
        return s.as_dict()

    def set_pair_counters(self, total, n, short, used):
        check(self._L.vspe_set_pair_counters(self._ctx, total, n, short, used))

    def map_reads(self, fq) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """-> (offsets[R+1], nodes, status[R]) for every complete record of one FASTQ buffer."""
        b = _u8(fq)
        nr, po, pn, ps = ctypes.c_uint64(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        check(self._L.vspe_map_reads(self._ctx, b.ctypes.data if b.size else None, b.size, ctypes.byref(nr),
                                     ctypes.byref(po), ctypes.byref(pn), ctypes.byref(ps)))
        R = nr.value
        off = np.ctypeslib.as_array(ctypes.cast(po, ctypes.POINTER(ctypes.c_uint64)), (R + 1,)).copy()
        nodes = np.ctypeslib.as_array(ctypes.cast(pn, ctypes.POINTER(ctypes.c_uint32)), (max(int(off[-1]), 1),)).copy()[:int(off[-1])]
        status = np.ctypeslib.as_array(ctypes.cast(ps, ctypes.POINTER(ctypes.c_uint8)), (max(R, 1),)).copy()[:R] if R else np.zeros(0, np.uint8)
        return off, nodes, status

    def split_records(self, fq) -> Tuple[int, np.ndarray, np.ndarray]:
        b = _u8(fq)
        nl, nr, ps, pl = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_void_p(), ctypes.c_void_p()
        check(self._L.vspe_split_records(self._ctx, b.ctypes.data if b.size else None, b.size, ctypes.byref(nl),
                                         ctypes.byref(nr), ctypes.byref(ps), ctypes.byref(pl)))
        R = nr.value
        if R == 0:
            return nl.value, np.zeros(0, np.uint64), np.zeros(0, np.uint32)
        st = np.ctypeslib.as_array(ctypes.cast(ps, ctypes.POINTER(ctypes.c_uint64)), (R,)).copy()
        ln = np.ctypeslib.as_array(ctypes.cast(pl, ctypes.POINTER(ctypes.c_uint32)), (R,)).copy()
        return nl.value, st, ln


def read_input(path: str) -> bytes:
    """The bytes the CLI would process for ``path``: a plain file as is, a gzip file (detected by
    its magic bytes; concatenated members included) inflated.  Host-only, needs no GPU."""
    L = _lib.lib()
    data, n = ctypes.c_void_p(), ctypes.c_uint64()
    check(L.vspe_read_input(os.fsencode(path), ctypes.byref(data), ctypes.byref(n)))
    try:
        return ctypes.string_at(data, n.value)
    finally:
        L.vspe_free_input(data)


def pe_inference(gfa: bytes, fwd, rve, kmer_size: int, device: int = 0, options: Optional[dict] = None):
    """In-memory equivalent of the reference script: -> (ids, node_mat, short_mat, stats)."""
    ids, seqs = parse_gfa_nodes(gfa)
    with PEIndex(seqs, kmer_size, device) as ix:
        for k, v in (options or {}).items():
            ix.set_option(k, v)
        ix.count_host(fwd, rve)
        node, short = ix.matrices()
        return ids, node, short, ix.stats()


def write_info(path: str, ids: Sequence[str], mat: np.ndarray):
    """Dense ``id_i:id_j:count`` writer (reference :194-207) through the C ABI."""
    n = len(ids)
    arr = (ctypes.c_char_p * max(n, 1))(*[i.encode() for i in ids])
    m = np.ascontiguousarray(mat, dtype=np.uint64)
    check(_lib.lib().vspe_write_info(path.encode(), arr, n, m.ctypes.data if n else None))


def write_info_sparse(path: str, ids: Sequence[str], keys: np.ndarray, counts: np.ndarray, mat: int):
    """Only the non-zero ``id_i:id_j:count`` lines of matrix ``mat`` (0 pe_info / node_mat, 1 st_info /
    short_mat).  ``process_pe_info`` (reference utils/VStrains_IO.py:598-612) parses it to the same
    dict as the dense file because it zero-initialises every key."""
    n = len(ids)
    arr = (ctypes.c_char_p * max(n, 1))(*[i.encode() for i in ids])
    k = np.ascontiguousarray(keys, dtype=np.uint64)
    c = np.ascontiguousarray(counts, dtype=np.uint64)
    check(_lib.lib().vspe_write_info_sparse(path.encode(), arr, n, k.ctypes.data if k.size else None,
                                            c.ctypes.data if c.size else None, k.size, mat))


def pe_info_dict(ids: Sequence[str], node_mat: np.ndarray, short_mat: np.ndarray) -> dict:
    """The dict the pipeline's consumer builds from the two files (reference
    utils/VStrains_IO.py:598-627 ``process_pe_info``), straight from the matrices: key
    ``(min(u, v), max(u, v))`` over the id STRINGS (as the reference compares them), value = sum of
    both matrices over both orientations.  Lets a caller skip the N*N-line text round trip
    (SURVEY section 8f, row 3).  Like the reference's parser, a repeated id folds onto one key."""
    n = len(ids)
    tot = np.asarray(node_mat, dtype=np.uint64).reshape(n, n) + np.asarray(short_mat, dtype=np.uint64).reshape(n, n)
    out = {}
    for a in range(n):
        u = ids[a]
        row = tot[a]
        for b in range(n):
            v = ids[b]
            key = (u, v) if u <= v else (v, u)
            out[key] = out.get(key, 0) + int(row[b])
    return out


def single_end_read_mapping(seq: str, kmer_htable, index2seqlen: list, split_len: int, len_index2id: int):
    """API-fidelity shim for reference :16-48.  ``kmer_htable`` must be a :class:`PEIndex`
    (the device index); the Python dict of the reference is not accepted because this package
    has no CPU path."""
    if not isinstance(kmer_htable, PEIndex):
        raise TypeError("kmer_htable must be a vstrains_b200.pe_inference.PEIndex (device index)")
    fq = ("@q\n%s\n+\n%s\n" % (seq, "I" * len(seq))).encode()
    off, nodes, status = kmer_htable.map_reads(fq)
    return [int(x) for x in nodes[off[0]:off[1]]] if status[0] == 0 else []


def main(argv: Optional[Sequence[str]] = None):
    print("----------------------Paired-End Information Alignment----------------------")
    parser = argparse.ArgumentParser(prog="pe_info",
                                     description="""Align Paired-End reads to nodes in graph to obtain strong links""")
    parser.add_argument("-g", "--gfa,", dest="gfa", type=str, required=True, help="graph, .gfa format")
    parser.add_argument("-o", "--output_dir", dest="dir", type=str, required=True, help="output directory")
    parser.add_argument("-f", "--forward", dest="fwd", required=True, help="forward read, .fastq")
    parser.add_argument("-r", "--reverse", dest="rve", required=True, help="reverse read, .fastq")
    parser.add_argument("-k", "--kmer_size", dest="kmer_size", type=int, default=128, help="unique kmer size")
    parser.add_argument("--gpus", dest="gpus", type=int, default=int(os.environ.get("VSPE_GPUS", "1")),
                        help="GPUs of this box to shard read pairs over (env VSPE_GPUS)")
    args = parser.parse_args(argv)
    glb_start = time.time()
    print("Start aligning reads to gfa nodes")
    st = Stats()
    L = _lib.lib()
    check(L.vspe_run(args.gfa.encode(), args.fwd.encode(), args.rve.encode(), args.kmer_size, args.dir.encode(),
                     args.gpus, ctypes.byref(st)))
    out_dir = args.dir[:-1] if args.dir.endswith("/") else args.dir
    print("Number of processed reads: ", st.total_pairs)
    print("pairs used / with N / too short: ", st.used_pairs, st.n_pairs, st.short_pairs)
    print("Global time elapsed: ", time.time() - glb_start)
    print("result stored in: ", "{0}/pe_info".format(out_dir))


if __name__ == "__main__":
    try:
        main()
    except VspeError as e:
        print(str(e), file=sys.stderr)
        sys.exit(1)
    sys.exit(0)
