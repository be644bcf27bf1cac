"""ctypes binding of libvspe.so (C ABI declared in include/vspe.h).

The library is the product: if it cannot be loaded, or no sm_100 device works, every entry
point raises -- there is no CPU fallback on this path."""
from __future__ import annotations

import ctypes
import os
import re
from typing import List

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VSPE_LIB_PATH") or os.path.join(_HERE, "libvspe.so")   # (VSPE_LIB_PATH: kernel experiments)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "vspe.h")


class VspeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libvspe error %d: %s" % (code, msg))
        self.code = code


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in (
        "total_pairs", "n_pairs", "short_pairs", "used_pairs", "bytes_fwd", "bytes_rve", "n_nodes",
        "n_kmers", "table_slots", "reads_fast", "reads_generic", "n_keys", "kernel_launches")] + \
        [(n, ctypes.c_float) for n in ("ms_index", "ms_h2d", "ms_scan", "ms_map", "ms_count", "ms_total", "ms_k_scan_rows")] + \
        [("n_k_scan_rows", ctypes.c_uint32), ("ms_k_walk", ctypes.c_float), ("n_k_walk", ctypes.c_uint32),
         ("scan_redo_tiles", ctypes.c_uint32), ("reserved0", ctypes.c_uint32), ("reads_memo", ctypes.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def declared_symbols() -> List[str]:
    """Every function include/vspe.h declares (used by the no-GPU ABI test)."""
    with open(HEADER_PATH) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(vspe_[a-z_0-9]+)\s*\(", src)))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VspeError(-6, "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    P, u64, u32, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    L.vspe_last_error.restype = ctypes.c_char_p
    L.vspe_version.restype = ctypes.c_char_p
    L.vspe_create.argtypes = [i32, ctypes.POINTER(P)]
    L.vspe_destroy.argtypes = [P]
    L.vspe_destroy.restype = None
    L.vspe_read_input.argtypes = [ctypes.c_char_p, ctypes.POINTER(P), ctypes.POINTER(u64)]
    L.vspe_free_input.argtypes = [P]
    L.vspe_free_input.restype = None
    L.vspe_index_build.argtypes = [P, P, P, u32, u32]
    L.vspe_reset.argtypes = [P]
    L.vspe_count_device.argtypes = [P, P, u64, P, u64]
    L.vspe_count_host.argtypes = [P, P, u64, P, u64]
    L.vspe_matrices_device.argtypes = [P, ctypes.POINTER(P), ctypes.POINTER(u64)]
    L.vspe_matrices_host.argtypes = [P, P, P]
    L.vspe_get_stats.argtypes = [P, ctypes.POINTER(Stats)]
    L.vspe_set_pair_counters.argtypes = [P, u64, u64, u64, u64]
    L.vspe_map_reads.argtypes = [P, P, u64, ctypes.POINTER(u64), ctypes.POINTER(P), ctypes.POINTER(P), ctypes.POINTER(P)]
    L.vspe_split_records.argtypes = [P, P, u64, ctypes.POINTER(u64), ctypes.POINTER(u64), ctypes.POINTER(P), ctypes.POINTER(P)]
    L.vspe_write_info.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), u32, P]
    L.vspe_run.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, i32, ctypes.c_char_p, i32, ctypes.POINTER(Stats)]
    L.vspe_is_sparse.argtypes = [P]
    L.vspe_sparse_host.argtypes = [P, ctypes.POINTER(u64), ctypes.POINTER(P), ctypes.POINTER(P)]
    L.vspe_sparse_merge.argtypes = [P, P, P, u64]
    L.vspe_sparse_device.argtypes = [P, ctypes.POINTER(u64), ctypes.POINTER(P), ctypes.POINTER(P)]
    L.vspe_sparse_merge_device.argtypes = [P, P, P, u64]
    L.vspe_sparse_clear.argtypes = [P]
    L.vspe_stream.argtypes = [P]
    L.vspe_stream.restype = P
    L.vspe_write_info_sparse.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), u32, P, P, u64, i32]
    L.vspe_alloc_pinned.argtypes = [ctypes.c_size_t]
    L.vspe_alloc_pinned.restype = P
    L.vspe_free_pinned.argtypes = [P]
    L.vspe_free_pinned.restype = None
    L.vspe_set_option.argtypes = [P, ctypes.c_char_p, ctypes.c_int64]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise VspeError(rc, lib().vspe_last_error().decode(errors="replace"))
