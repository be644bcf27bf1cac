"""ctypes loader for the C restatement (oracle/pe_oracle.c) -- test infrastructure only."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return os.path.join(_HERE, "libpe_oracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libpe_oracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "pe_oracle.c")):
            build()
        L = ctypes.CDLL(path)
        u8p = ctypes.c_char_p
        L.vso_count_nodes.restype = ctypes.c_int64
        L.vso_count_nodes.argtypes = [u8p, ctypes.c_int64]
        L.vso_run.restype = ctypes.c_int
        L.vso_run.argtypes = [u8p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                              ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                              ctypes.c_void_p, ctypes.c_int]
        L.vso_map_reads.restype = ctypes.c_int
        L.vso_map_reads.argtypes = [u8p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                    ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64)]
        L.vso_free.argtypes = [ctypes.c_void_p]
        L.vso_write_info.restype = ctypes.c_int
        L.vso_write_info.argtypes = [ctypes.c_char_p, u8p, ctypes.c_int64, ctypes.c_void_p]
        _LIB = L
    return _LIB


def _buf(x):
    a = np.ascontiguousarray(np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray)) else x)
    return a, a.ctypes.data, a.size


def run(gfa: bytes, fwd, rve, kmer_size: int, nthreads: int = 0):
    """-> (node_mat int64[N,N], short_mat int64[N,N], stats dict).  Raises on oracle error codes."""
    L = lib()
    n = L.vso_count_nodes(gfa, len(gfa))
    if n < 0:
        raise ValueError("oracle error %d" % n)
    node = np.zeros((n, n), dtype=np.int64)
    short = np.zeros((n, n), dtype=np.int64)
    stats = np.zeros(4, dtype=np.int64)
    fa, fp, fn = _buf(fwd)
    ra, rp, rn = _buf(rve)
    rc = L.vso_run(gfa, len(gfa), fp, fn, rp, rn, kmer_size, node.ctypes.data, short.ctypes.data,
                   stats.ctypes.data, nthreads)
    if rc:
        raise ValueError("oracle error %d" % rc)
    return node, short, dict(zip(("total_pairs", "n_pairs", "short_pairs", "used_pairs"), map(int, stats)))


def map_reads(gfa: bytes, fq, kmer_size: int):
    """-> (offsets int64[R+1], nodes int32[...], status uint8[R])."""
    L = lib()
    fa, fp, fn = _buf(fq)
    off, nod, st = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    nreads = ctypes.c_int64()
    rc = L.vso_map_reads(gfa, len(gfa), fp, fn, kmer_size, ctypes.byref(off), ctypes.byref(nod),
                         ctypes.byref(st), ctypes.byref(nreads))
    if rc:
        raise ValueError("oracle error %d" % rc)
    R = nreads.value
    offsets = np.ctypeslib.as_array(ctypes.cast(off, ctypes.POINTER(ctypes.c_int64)), (R + 1,)).copy()
    nodes = np.ctypeslib.as_array(ctypes.cast(nod, ctypes.POINTER(ctypes.c_int32)), (max(int(offsets[-1]), 1),)).copy()[:offsets[-1]]
    status = np.ctypeslib.as_array(ctypes.cast(st, ctypes.POINTER(ctypes.c_uint8)), (R + 1,)).copy()[:R]
    for p in (off, nod, st):
        L.vso_free(p)
    return offsets, nodes, status


def info_bytes(ids, mat) -> bytes:
    """Dense ``id_i:id_j:count\\n`` text from a matrix (numpy-vectorised; test helper)."""
    n = len(ids)
    out = []
    for i in range(n):
        row = mat[i]
        out.append("".join("%s:%s:%d\n" % (ids[i], ids[j], row[j]) for j in range(n)))
    return "".join(out).encode()
