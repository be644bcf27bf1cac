"""CPU-side tests: C ABI surface, host-side mirror logic, dense writer (no GPU needed)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, parse_info
from oracle import pe_oracle
from vstrains_b200 import _lib, pe_inference


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    syms = _lib.declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert getattr(L, s) is not None
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(syms) <= exported


def test_library_is_built_for_sm100a_only():
    out = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "-lelf", _lib.LIB_PATH]).decode()
    archs = {l.split(".")[-2] for l in out.splitlines() if "sm_" in l}
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_a_gpu():
    p = ctypes.c_void_p()
    rc = _lib.lib().vspe_create(0, ctypes.byref(p))
    assert rc == -6
    assert b"no CPU fallback" in _lib.lib().vspe_last_error()
    with pytest.raises(_lib.VspeError):
        pe_inference.pe_inference(b"S\t1\tACGTACGT\n", b"", b"", 3)


def test_parse_gfa_nodes_matches_oracle(golden):
    ids, seqs = pe_inference.parse_gfa_nodes(golden.gfa)
    oids, oseqs = pe_oracle.parse_gfa(golden.gfa)
    assert ids == oids
    assert [s.decode() for s in seqs] == oseqs


def test_parse_gfa_unterminated_last_line_loses_a_char():
    gfa = b"S\ta\tACGTAC\nS\tb\tGGGTTT"
    ids, seqs = pe_inference.parse_gfa_nodes(gfa)
    assert (ids, seqs) == pe_oracle.parse_gfa(gfa)[0:1] + ([b"ACGTAC", b"GGGTT"],)


def test_dense_writer_reproduces_reference_files(golden, tmp_path):
    if golden.status != 0:
        return
    ids, _ = pe_oracle.parse_gfa(golden.gfa)
    for name, data in (("pe_info", golden.pe_info), ("st_info", golden.st_info)):
        mat = parse_info(data, ids) if ids else np.zeros((0, 0), np.int64)
        path = str(tmp_path / name)
        pe_inference.write_info(path, ids, mat.astype(np.uint64))
        with open(path, "rb") as f:
            assert f.read() == data


def test_dense_writer_large_values_and_threads(tmp_path):
    n = 150
    rng = np.random.default_rng(3)
    mat = rng.integers(0, 2**40, size=(n, n), dtype=np.uint64)
    mat[0, 0] = 0
    mat[1, 1] = 2**63 + 5
    ids = ["n%d" % (i * 7919) for i in range(n)]
    path = str(tmp_path / "pe_info")
    pe_inference.write_info(path, ids, mat)
    exp = "".join("%s:%s:%d\n" % (ids[i], ids[j], int(mat[i, j])) for i in range(n) for j in range(n))
    with open(path) as f:
        assert f.read() == exp


def test_shims_keep_the_reference_names():
    assert pe_inference.reverse_seq("AACG") == "CGTT"
    with pytest.raises(KeyError):
        pe_inference.reverse_seq("ACNG")
    with pytest.raises(TypeError):
        pe_inference.single_end_read_mapping("ACGT", {}, [4], 3, 1)
    sys.path.insert(0, os.path.join(ROOT, "utils"))
    import importlib
    mod = importlib.import_module("VStrains_PE_Inference")
    for name in ("main", "reverse_seq", "single_end_read_mapping"):
        assert hasattr(mod, name)


def test_sparse_writer_emits_only_nonzero_lines(tmp_path):
    ids = ["a", "-7", "x9"]
    n = len(ids)
    keys = np.array([1, 4, 8, n * n + 0, n * n + 5], dtype=np.uint64)
    counts = np.array([3, 0, 12, 7, 2**40], dtype=np.uint64)
    p0, p1 = str(tmp_path / "pe_info"), str(tmp_path / "st_info")
    pe_inference.write_info_sparse(p0, ids, keys, counts, 0)
    pe_inference.write_info_sparse(p1, ids, keys, counts, 1)
    assert open(p0).read() == "a:-7:3\nx9:x9:12\n"
    assert open(p1).read() == "a:a:7\n-7:x9:%d\n" % 2**40


def test_read_input_plain_and_gzip(tmp_path):
    """vspe_read_input (what vspe_run feeds the GPU path): plain bytes as is, gzip -- one member,
    concatenated members, empty payload -- inflated; truncated / corrupt streams are loud errors."""
    import gzip
    import zlib
    from vstrains_b200 import pe_inference
    from vstrains_b200._lib import VspeError
    rng = np.random.default_rng(5)
    reads = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(list(b"ACGT"), 150).astype(np.uint8)), b"I" * 150)
                     for i in range(3000))
    plain = tmp_path / "r.fq"
    plain.write_bytes(reads)
    assert pe_inference.read_input(str(plain)) == reads
    one = tmp_path / "r.fq.gz"
    one.write_bytes(gzip.compress(reads))
    assert pe_inference.read_input(str(one)) == reads
    cut = len(reads) // 3
    multi = tmp_path / "m.fq.gz"
    multi.write_bytes(gzip.compress(reads[:cut]) + gzip.compress(reads[cut:2 * cut], 1) + gzip.compress(reads[2 * cut:], 9))
    assert pe_inference.read_input(str(multi)) == reads
    empty = tmp_path / "e.gz"
    empty.write_bytes(gzip.compress(b""))
    assert pe_inference.read_input(str(empty)) == b""
    zero = tmp_path / "zero"
    zero.write_bytes(b"")
    assert pe_inference.read_input(str(zero)) == b""
    # highly compressible payload: the output buffer has to grow well beyond 4x the input
    big = tmp_path / "big.gz"
    big.write_bytes(gzip.compress(b"A" * (8 << 20)))
    assert pe_inference.read_input(str(big)) == b"A" * (8 << 20)
    trunc = tmp_path / "t.gz"
    trunc.write_bytes(gzip.compress(reads)[:-40])
    with pytest.raises(VspeError):
        pe_inference.read_input(str(trunc))
    bad = bytearray(gzip.compress(reads))
    bad[len(bad) // 2] ^= 0xFF
    corrupt = tmp_path / "c.gz"
    corrupt.write_bytes(bytes(bad))
    with pytest.raises(VspeError):
        pe_inference.read_input(str(corrupt))
    with pytest.raises(VspeError):
        pe_inference.read_input(str(tmp_path / "missing.fq"))
    # a zlib (not gzip) stream is not inflated: the bytes pass through and the scan rejects them later
    z = tmp_path / "z.bin"
    z.write_bytes(zlib.compress(reads))
    assert pe_inference.read_input(str(z)) == zlib.compress(reads)


def _consumer_parse(node_ids, pe_bytes, st_bytes):
    """Test-side restatement of the pipeline's consumer of pe_info / st_info (reference
    utils/VStrains_IO.py:598-627): zero-initialise every unordered id pair, then add the third
    field of every line (up to a blank line) to the pair's entry if it exists."""
    table = {}
    for u in node_ids:
        for v in node_ids:
            table[(min(u, v), max(u, v))] = 0
    for blob in (pe_bytes, st_bytes):
        for line in blob.decode().splitlines(keepends=True):
            if line == "\n":
                break
            u, v, mark = line[:-1].split(":")[:3]
            key = (min(u, v), max(u, v))
            if key in table:
                table[key] += int(mark)
    return table


def test_consumer_sees_the_same_dict_from_files_and_from_matrices(golden, tmp_path):
    """What VStrains does next with the two files (process_pe_info) gives the same dict as
    pe_info_dict() on the matrices, and as the sparse (non-zero lines only) files."""
    if golden.status != 0:
        return
    ids, _ = pe_inference.parse_gfa_nodes(golden.gfa)
    if len(set(ids)) != len(ids) or any(":" in i for i in ids):
        return                                   # the reference's own text format is ambiguous for such ids
    node = parse_info(golden.pe_info, ids)
    short = parse_info(golden.st_info, ids)
    want = _consumer_parse(ids, golden.pe_info, golden.st_info)
    assert pe_inference.pe_info_dict(ids, node, short) == want
    # sparse files: only the non-zero lines
    flat = np.concatenate([node.reshape(-1), short.reshape(-1)]).astype(np.uint64)
    keys = np.nonzero(flat)[0].astype(np.uint64)
    counts = flat[keys.astype(np.int64)]
    pe_inference.write_info_sparse(str(tmp_path / "pe"), ids, keys, counts, 0)
    pe_inference.write_info_sparse(str(tmp_path / "st"), ids, keys, counts, 1)
    assert _consumer_parse(ids, (tmp_path / "pe").read_bytes(), (tmp_path / "st").read_bytes()) == want


def test_every_library_option_is_documented_in_the_header():
    """vspe_set_option's names (api.cu) and the list in include/vspe.h must not drift apart."""
    import re
    with open(os.path.join(ROOT, "vstrains_b200", "csrc", "api.cu")) as f:
        src = f.read()
    body = src[src.index("int vspe_set_option("):]
    names = set(re.findall(r'!strcmp\(name, "([a-z_0-9]+)"\)', body))
    assert len(names) >= 7
    with open(os.path.join(ROOT, "include", "vspe.h")) as f:
        hdr = f.read()
    doc = hdr[hdr.index("/* Tunables."):hdr.index("int vspe_set_option(")]
    documented = set(re.findall(r'"([a-z_0-9]+)"', doc))
    assert names <= documented, sorted(names - documented)
    assert documented <= names, sorted(documented - names)


def test_stats_struct_mirror_matches_the_header():
    """vstrains_b200._lib.Stats (ctypes) must declare the fields of vspe_stats (include/vspe.h) in the same order and
    with the same types: a drift would silently shift every counter the tests and the bench read."""
    import ctypes
    import re
    from vstrains_b200 import _lib
    with open(os.path.join(ROOT, "include", "vspe.h")) as f:
        hdr = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    body = hdr[hdr.index("typedef struct vspe_stats {") + len("typedef struct vspe_stats {"):hdr.index("} vspe_stats;")]
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, names = decl.split(None, 1)
        for name in names.split(","):
            fields.append((name.strip(), ctype))
    cmap = {"uint64_t": ctypes.c_uint64, "uint32_t": ctypes.c_uint32, "float": ctypes.c_float}
    assert [(n, cmap[t]) for n, t in fields] == list(_lib.Stats._fields_)
    assert ctypes.sizeof(_lib.Stats) % 8 == 0
