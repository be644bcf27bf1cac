"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the committed
outputs of the reference script.  Bit-exact: everything on this path is integer / byte work."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import c_oracle, pe_oracle
import synthgen as synth
from vstrains_b200 import pe_inference
from vstrains_b200._lib import VspeError

pytestmark = pytest.mark.gpu


# kernel-path variants every parity case runs through (library options, see vspe_set_option)
VARIANTS = [
    {"scan_mode": 0},                                   # default: k_scan_rows + k_walk + list-driven tiers
    {"scan_mode": 0, "subst": 0},                       # ... without the substitution-hit bitmap (no error tolerance in the walk)
    {"scan_mode": 0, "memo": 0},                        # ... without the read memo (every read walked)
    {"scan_mode": 1},                                   # look-back record scan + raw-byte seed-and-extend kernel
    {"scan_mode": 1, "subst": 0},
    {"scan_mode": 1, "force_generic": 1},               # exhaustive ASCII tier only
    {"scan_mode": 2},                                   # two-pass record scan
]


def _variant_id(o):
    return ",".join("%s%s" % kv for kv in o.items())


def _info(ids, mat, tmp_path, name):
    path = str(tmp_path / name)
    pe_inference.write_info(path, ids, mat)
    with open(path, "rb") as f:
        return f.read()


@pytest.mark.parametrize("options", VARIANTS, ids=_variant_id)
def test_golden_fixtures_bit_exact(golden, tmp_path, options):
    if golden.status != 0:
        with pytest.raises(VspeError) as ei:
            pe_inference.pe_inference(golden.gfa, golden.fwd, golden.rve, golden.k)
        assert ei.value.code == -2
        return
    ids, node, short, stats = pe_inference.pe_inference(golden.gfa, golden.fwd, golden.rve, golden.k,
                                                        options=options)
    assert _info(ids, node, tmp_path, "pe_info") == golden.pe_info
    assert _info(ids, short, tmp_path, "st_info") == golden.st_info
    _, _, ostats, _ = pe_oracle.run_bytes(golden.gfa, golden.fwd, golden.rve, golden.k)
    for k, v in ostats.items():
        assert stats[k] == v


def test_cli_drop_in_writes_identical_files(golden, tmp_path):
    for name, data in (("g.gfa", golden.gfa), ("f.fq", golden.fwd), ("r.fq", golden.rve)):
        (tmp_path / name).write_bytes(data)
    out = tmp_path / "aln"
    out.mkdir()
    (out / "stale").write_text("x")            # the script owns DIR: rm -rf then mkdir
    p = subprocess.run([sys.executable, os.path.join(ROOT, "utils", "VStrains_PE_Inference.py"),
                        "-g", str(tmp_path / "g.gfa"), "-o", str(out) + "/", "-f", str(tmp_path / "f.fq"),
                        "-r", str(tmp_path / "r.fq"), "-k", str(golden.k)], capture_output=True)
    if golden.status != 0:
        assert p.returncode != 0
        return
    assert p.returncode == 0, p.stderr.decode()
    assert sorted(os.listdir(out)) == ["pe_info", "st_info"]
    assert (out / "pe_info").read_bytes() == golden.pe_info
    assert (out / "st_info").read_bytes() == golden.st_info
    assert b"Paired-End Information Alignment" in p.stdout


def test_cli_reads_gzip_inputs(tmp_path):
    """.fastq.gz (and a gzipped GFA) give the files the plain inputs give."""
    import gzip
    from conftest import golden_cases
    g = [c for c in golden_cases() if c.name == "lowcx_00"][0]
    (tmp_path / "g.gfa.gz").write_bytes(gzip.compress(g.gfa))
    cut = len(g.fwd) // 2
    (tmp_path / "f.fq.gz").write_bytes(gzip.compress(g.fwd[:cut]) + gzip.compress(g.fwd[cut:]))   # two members
    (tmp_path / "r.fq.gz").write_bytes(gzip.compress(g.rve))
    out = tmp_path / "aln"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "utils", "VStrains_PE_Inference.py"),
                        "-g", str(tmp_path / "g.gfa.gz"), "-o", str(out), "-f", str(tmp_path / "f.fq.gz"),
                        "-r", str(tmp_path / "r.fq.gz"), "-k", str(g.k)], capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    assert (out / "pe_info").read_bytes() == g.pe_info
    assert (out / "st_info").read_bytes() == g.st_info


@pytest.mark.parametrize("two_pass", [0, 1])
def test_record_split_matches_universal_newlines(golden, two_pass):
    with pe_inference.PEIndex([b"ACGTACGTAC"], 3) as ix:
        ix.set_option("scan_two_pass", two_pass)
        for fq in (golden.fwd, golden.rve):
            n_lines, start, length = ix.split_records(fq)
            lines = pe_oracle.split_lines(fq)
            assert n_lines == len(lines)
            assert len(start) == len(lines) // 4
            text = fq.decode()
            for r in range(len(start)):
                seq = lines[4 * r + 1][:-1]
                assert text[int(start[r]):int(start[r]) + int(length[r])] == seq


@pytest.mark.parametrize("two_pass", [0, 1])
def test_record_split_edge_cases(two_pass):
    cases = [b"\n" * 70000, b"a\n" * 40000, (b"@r\nAC\n+\nII\n" * 9000)[:-1],
             b"", b"\n", b"\r", b"\r\n", b"a", b"a\nb\nc\nd", b"a\nb\nc\nd\n", b"\n\n\n\n\n\n\n\n",
             b"@h\r\nAC\r\n+\r\nII\r\n@h\rGT\r+\rII\r", b"@h\nACGT\r\r\n+\nIIII\n", b"x" * 100000 + b"\n" + b"ACGT\n+\nIIII\n" * 3,
             b"@a\nAC" + b"G" * 70000 + b"\n+\n" + b"I" * 70002 + b"\n"]
    with pe_inference.PEIndex([b"ACGTACGTAC"], 3) as ix:
        ix.set_option("scan_two_pass", two_pass)
        for fq in cases:
            n_lines, start, length = ix.split_records(fq)
            lines = pe_oracle.split_lines(fq)
            assert n_lines == len(lines), fq[:40]
            assert len(start) == len(lines) // 4
            for r in range(len(start)):
                assert fq[int(start[r]):int(start[r]) + int(length[r])].decode() == lines[4 * r + 1][:-1]


@pytest.mark.parametrize("force_generic,scan_mode", [(0, 0), (0, 1), (1, 1)])
def test_per_read_mapping_matches_oracle(golden, force_generic, scan_mode):
    if golden.status != 0:
        return
    ids, seqs = pe_inference.parse_gfa_nodes(golden.gfa)
    with pe_inference.PEIndex(seqs, golden.k) as ix:
        ix.set_option("force_generic", force_generic)
        ix.set_option("scan_mode", scan_mode)
        for fq in (golden.fwd, golden.rve):
            off, nodes, status = ix.map_reads(fq)
            ooff, onodes, ostatus = c_oracle.map_reads(golden.gfa, fq, golden.k)
            assert np.array_equal(status, ostatus)
            assert np.array_equal(off.astype(np.int64), ooff)
            assert np.array_equal(nodes.astype(np.int64), onodes.astype(np.int64))


@pytest.mark.parametrize("name,pairs", [("C1", 6000), ("C2", 6000), ("C3", 4000), ("C4", 3000)])
@pytest.mark.parametrize("options", VARIANTS, ids=_variant_id)
def test_synthetic_configs_match_c_oracle(name, pairs, options):
    cfg = synth.CONFIGS[name]
    g, f, r = synth.generate(cfg, pairs=pairs)
    gfa = g.to_gfa()
    ids, node, short, stats = pe_inference.pe_inference(gfa, f, r, cfg.k, options=options)
    onode, oshort, ostats = c_oracle.run(gfa, f, r, cfg.k)
    assert np.array_equal(node.astype(np.int64), onode)
    assert np.array_equal(short.astype(np.int64), oshort)
    for k, v in ostats.items():
        assert stats[k] == v
    assert stats["kernel_launches"] > 0


def test_chunked_streaming_equals_single_chunk():
    cfg = synth.CONFIGS["C1"]
    g, f, r = synth.generate(cfg, pairs=9000)
    ids, seqs = pe_inference.parse_gfa_nodes(g.to_gfa())
    res = []
    for chunk_mb, two_pass, scan_mode in ((256, 0, 0), (1, 0, 0), (2, 0, 0), (1, 0, 1), (1, 1, 0), (256, 0, 1)):
        with pe_inference.PEIndex(seqs, cfg.k) as ix:
            ix.set_option("chunk_mb", chunk_mb)
            ix.set_option("scan_two_pass", two_pass)
            ix.set_option("scan_mode", scan_mode)
            ix.count_host(f, r)
            res.append(ix.matrices() + (ix.stats(),))
    for other in res[1:]:
        assert np.array_equal(res[0][0], other[0]) and np.array_equal(res[0][1], other[1])
        for k in ("total_pairs", "n_pairs", "short_pairs", "used_pairs", "n_keys"):
            assert res[0][2][k] == other[2][k]


def test_counts_accumulate_over_calls_and_shards_sum_exactly():
    """Multi-GPU contract on one device: disjoint record ranges summed == whole (integer adds)."""
    cfg = synth.CONFIGS["C3"]
    g, f, r = synth.generate(cfg, pairs=4000)
    ids, seqs = pe_inference.parse_gfa_nodes(g.to_gfa())
    rec_f, rec_r = (2 * 150 + 18), (2 * 150 + 18)
    from vstrains_b200 import shard
    with pe_inference.PEIndex(seqs, cfg.k) as ix:
        ix.count_host(f, r)
        whole = ix.matrices()
        wstats = ix.stats()
        ix.reset()
        for lo_f, hi_f, lo_r, hi_r in shard.shard_ranges(f, r, 3):
            ix.count_host(f[lo_f:hi_f], r[lo_r:hi_r])
        parts = ix.matrices()
        pstats = ix.stats()
    assert np.array_equal(whole[0], parts[0]) and np.array_equal(whole[1], parts[1])
    for k in ("total_pairs", "n_pairs", "short_pairs", "used_pairs"):
        assert wstats[k] == pstats[k]


def test_device_resident_entry_point_matches_host_entry_point():
    import torch
    cfg = synth.CONFIGS["C2"]
    g, f, r = synth.generate(cfg, pairs=5000)
    ids, seqs = pe_inference.parse_gfa_nodes(g.to_gfa())
    with pe_inference.PEIndex(seqs, cfg.k) as ix:
        ix.count_host(f, r)
        host = ix.matrices()
        ix.reset()
        # deliberately misaligned device buffers (shards start at arbitrary byte offsets)
        df = torch.zeros(f.size + 7, dtype=torch.uint8, device="cuda")
        dr = torch.zeros(r.size + 3, dtype=torch.uint8, device="cuda")
        df[7:] = torch.from_numpy(f).cuda()
        dr[3:] = torch.from_numpy(r).cuda()
        torch.cuda.synchronize()
        ix.count_device(df.data_ptr() + 7, f.size, dr.data_ptr() + 3, r.size)
        dev = ix.matrices()
        p, n = ix.matrices_device()
        assert p != 0 and n == 2 * len(seqs) ** 2
    assert np.array_equal(host[0], dev[0]) and np.array_equal(host[1], dev[1])


def test_non_ascii_input_is_rejected():
    with pytest.raises(VspeError) as ei:
        pe_inference.pe_inference(b"S\t1\tACGTACGTAC\n", b"@r\nACGT\xc3\xa9\n+\nIIIII\n", b"@r\nACGTA\n+\nIIIII\n", 3)
    assert ei.value.code == -3


def test_run_rank_single_process_matches_oracle():
    from vstrains_b200 import dist as vdist
    cfg = synth.CONFIGS["C2"]
    g, f, r = synth.generate(cfg, pairs=3000)
    gfa = g.to_gfa()
    ids, node, short, counters = vdist.run_rank(gfa, f, r, cfg.k, 0, 1, device=0)
    onode, oshort, ostats = c_oracle.run(gfa, f, r, cfg.k)
    assert np.array_equal(node.astype(np.int64), onode) and np.array_equal(short.astype(np.int64), oshort)
    assert counters == ostats


def test_cli_multi_gpu_is_byte_identical(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cfg = synth.CONFIGS["C3"]
    g, f, r = synth.generate(cfg, pairs=5000)
    (tmp_path / "g.gfa").write_bytes(g.to_gfa())
    f.tofile(tmp_path / "f.fq")
    r.tofile(tmp_path / "r.fq")
    outs = []
    for gpus in (1, 2):
        out = tmp_path / ("aln%d" % gpus)
        env = dict(os.environ, VSPE_GPUS=str(gpus))
        p = subprocess.run([sys.executable, os.path.join(ROOT, "utils", "VStrains_PE_Inference.py"), "-g", str(tmp_path / "g.gfa"),
                            "-o", str(out), "-f", str(tmp_path / "f.fq"), "-r", str(tmp_path / "r.fq"), "-k", str(cfg.k)],
                           capture_output=True, env=env)
        assert p.returncode == 0, p.stderr.decode()
        outs.append(((out / "pe_info").read_bytes(), (out / "st_info").read_bytes()))
    assert outs[0] == outs[1]


def test_cli_multi_gpu_sparse_merge_is_identical(tmp_path):
    """vspe_run with n_gpus = 2 in sparse mode: the runs are exchanged with NCCL all-gathers and
    merged on device 0; the files must equal the single-GPU sparse files line for line."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cfg = synth.CONFIGS["C3"]
    g, f, r = synth.generate(cfg, pairs=5000)
    (tmp_path / "g.gfa").write_bytes(g.to_gfa())
    f.tofile(tmp_path / "f.fq")
    r.tofile(tmp_path / "r.fq")
    outs = []
    for gpus in (1, 2):
        out = tmp_path / ("aln%d" % gpus)
        env = dict(os.environ, VSPE_GPUS=str(gpus), VSPE_SPARSE="1")
        p = subprocess.run([sys.executable, os.path.join(ROOT, "utils", "VStrains_PE_Inference.py"), "-g", str(tmp_path / "g.gfa"),
                            "-o", str(out), "-f", str(tmp_path / "f.fq"), "-r", str(tmp_path / "r.fq"), "-k", str(cfg.k)],
                           capture_output=True, env=env)
        assert p.returncode == 0, p.stderr.decode()
        outs.append(((out / "pe_info").read_bytes(), (out / "st_info").read_bytes()))
    assert outs[0] == outs[1]
    assert len(outs[0][0]) > 0


@pytest.mark.parametrize("world", [2, 8])
def test_torchrun_nccl_ranks_match_oracle(world):
    """One process per GPU (torchrun, NCCL): dist.run_rank on record-aligned shards + one allreduce
    must reproduce the C oracle's matrices of the whole input on every rank."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs >= %d GPUs" % world)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(29513 + world),
                        os.path.join(ROOT, "tests", "nccl_rank_check.py"), "C3", "20000"], capture_output=True, timeout=600)
    assert p.returncode == 0, (p.stdout.decode()[-2000:], p.stderr.decode()[-2000:])
    assert p.stdout.decode().count("bit-exact") == world


def _mk_fastq(seqs, nl=b"\n"):
    return b"".join(b"@r" + nl + s + nl + b"+" + nl + b"I" * len(s) + nl for s in seqs)


def _decoy_fastq(seqs, seq_prefix):
    out = []
    for i, q in enumerate(seqs):
        if seq_prefix:
            out.append(b"@r%d\n" % i + seq_prefix + q + b"\n+\n@" + b"I" * len(q) + b"\n")
        else:                                   # no '@' / '+' anywhere
            out.append(b"r%d\n" % i + q + b"\n-\n" + b"I" * len(q) + b"\n")
    return b"".join(out)


@pytest.mark.parametrize("scan_mode", [0, 1, 2])
def test_whole_path_edge_shapes(scan_mode):
    """Shapes that stress the tiled scan: thousands of tiny records per tile (fallback path),
    reads longer than the packed rows / the scan margin, CRLF and lone-CR files, reads that
    straddle tile boundaries, misaligned shards."""
    rng = np.random.default_rng(17)
    cfg = synth.Config("edge", 1500, 3, 0.02, 150, 10, 23)
    st = synth.make_strains(cfg.genome_len, cfg.strains, cfg.divergence, rng)
    seqs, cov, links = synth.build_dbg(st, 31, None, seed=3)
    gfa = b"".join(b"S\t%d\t%s\n" % (i, s) for i, s in enumerate(seqs))
    genome = synth._ACGT[st[0]].tobytes()

    def reads(n, lo, hi):
        out = []
        for _ in range(n):
            ln = int(rng.integers(lo, hi))
            a = int(rng.integers(0, len(genome) - ln))
            out.append(genome[a:a + ln])
        return out

    cases = {
        "tiny": (_mk_fastq(reads(30000, 1, 12)), _mk_fastq(reads(30000, 33, 40))),
        "long": (_mk_fastq(reads(300, 300, 1400)), _mk_fastq(reads(300, 100, 700))),
        "crlf": (_mk_fastq(reads(4000, 40, 160), b"\r\n"), _mk_fastq(reads(4000, 40, 160), b"\r")),
        "mixed": (_mk_fastq(reads(3000, 1, 330)), _mk_fastq(reads(2990, 1, 330))[:-1]),
        # text that would fool any guess of the line phase: quality lines that start with '@' followed,
        # two lines on, by "sequence" lines that start with '+'; and headers without '@'
        "decoy": (_decoy_fastq(reads(25000, 35, 90), b"+"), _decoy_fastq(reads(25000, 35, 90), b"")),
    }
    for name, (f, r) in cases.items():
        ids, node, short, stats = pe_inference.pe_inference(gfa, f, r, 31, options={"scan_mode": scan_mode, "chunk_mb": 1})
        onode, oshort, ostats = c_oracle.run(gfa, f, r, 31)
        assert np.array_equal(node.astype(np.int64), onode), name
        assert np.array_equal(short.astype(np.int64), oshort), name
        for k, v in ostats.items():
            assert stats[k] == v, (name, k)
        if scan_mode == 0:
            # tiles guess their line phase and a wrong guess is packed again: only the decoy text may (and must) cause that
            assert (stats["scan_redo_tiles"] > 0) == (name == "decoy"), (name, stats["scan_redo_tiles"])


@pytest.mark.parametrize("subst,scan_mode", [(1, 0), (0, 0), (1, 1), (0, 1)])
@pytest.mark.parametrize("sub_rate", [0.002, 0.01, 0.04])
def test_noisy_reads_match_c_oracle(subst, scan_mode, sub_rate):
    """1 % and 4 % substitution rates: several errors per read, errors next to node ends and to
    each other, reads that follow another strain's bubble arm after an error."""
    for name, pairs in (("C2", 4000), ("C3", 3000)):
        cfg = synth.CONFIGS[name]
        rng = np.random.default_rng(cfg.seed + 5)
        g, genomes, ab = synth.make_graph(cfg, rng, 10.0)
        f, r = synth.make_reads(genomes, ab, cfg.read_len, pairs, cfg.k, rng, sub_rate=sub_rate)
        gfa = g.to_gfa()
        ids, node, short, stats = pe_inference.pe_inference(gfa, f, r, cfg.k, options={"subst": subst, "scan_mode": scan_mode})
        onode, oshort, ostats = c_oracle.run(gfa, f, r, cfg.k)
        assert np.array_equal(node.astype(np.int64), onode), name
        assert np.array_equal(short.astype(np.int64), oshort), name
        for k, v in ostats.items():
            assert stats[k] == v


# ------------------------------------------------------------------------------------------------
# sparse (COO) counting: 64-bit keys, LSD radix sort + run-length reduce
# ------------------------------------------------------------------------------------------------
def _coo_from_dense(node, short):
    n = node.shape[0]
    flat = np.concatenate([node.reshape(-1), short.reshape(-1)]).astype(np.uint64)
    keys = np.nonzero(flat)[0].astype(np.uint64)
    return keys, flat[keys.astype(np.int64)]


def test_sparse_mode_equals_dense_on_golden(golden):
    if golden.status != 0:
        return
    ids, seqs = pe_inference.parse_gfa_nodes(golden.gfa)
    with pe_inference.PEIndex(seqs, golden.k) as ix:
        ix.count_host(golden.fwd, golden.rve)
        dn, ds = ix.matrices()
        ix.set_option("sparse", 1)
        assert ix.is_sparse
        ix.reset()
        ix.count_host(golden.fwd, golden.rve)
        keys, counts = ix.sparse()
        ek, ec = _coo_from_dense(dn, ds)
        assert np.array_equal(keys, ek) and np.array_equal(counts, ec)
        assert np.all(np.diff(keys.astype(np.int64)) > 0)


@pytest.mark.parametrize("name,pairs", [("C2", 30000), ("C3", 12000)])
def test_sparse_mode_batches_and_merge_match_oracle(name, pairs):
    cfg = synth.CONFIGS[name]
    g, f, r = synth.generate(cfg, pairs=pairs)
    gfa = g.to_gfa()
    ids, seqs = pe_inference.parse_gfa_nodes(gfa)
    onode, oshort, ostats = c_oracle.run(gfa, f, r, cfg.k)
    ek, ec = _coo_from_dense(onode.astype(np.uint64), oshort.astype(np.uint64))
    from vstrains_b200 import shard
    with pe_inference.PEIndex(seqs, cfg.k) as ix, pe_inference.PEIndex(seqs, cfg.k) as iy:
        for i in (ix, iy):
            i.set_option("sparse", 1)
        ix.count_host(f, r)
        keys, counts = ix.sparse()
        assert np.array_equal(keys, ek) and np.array_equal(counts, ec)
        st = ix.stats()
        for k, v in ostats.items():
            assert st[k] == v
        # three shards: two accumulated on one context, one on another, then merged
        ix.reset()
        rng = shard.shard_ranges(f, r, 3)
        for a, b, c_, d in rng[:2]:
            ix.count_host(f[a:b], r[c_:d])
        a, b, c_, d = rng[2]
        iy.count_host(f[a:b], r[c_:d])
        ix.sparse_merge(*iy.sparse())
        keys, counts = ix.sparse()
        assert np.array_equal(keys, ek) and np.array_equal(counts, ec)


def test_large_graph_switches_to_sparse_automatically(tmp_path):
    """N above the dense limit (11 585): no N*N matrices anywhere; checked against the Python
    oracle's sparse dicts; the CLI writes the non-zero lines only."""
    cfg = synth.Config("big", 10_000, 4, 0.01, 150, 2000, 77, n_genomes=28)
    g, f, r = synth.generate(cfg)
    gfa = g.to_gfa()
    ids, seqs = pe_inference.parse_gfa_nodes(gfa)
    n = len(ids)
    assert n > 11585
    with pe_inference.PEIndex(seqs, cfg.k) as ix:
        assert ix.is_sparse
        ix.count_host(f, r)
        keys, counts = ix.sparse()
        stats = ix.stats()
    oids, oseqs = pe_oracle.parse_gfa(gfa)
    table = pe_oracle.build_index(oseqs, cfg.k + 1)
    node, short, ostats = pe_oracle.count_pairs(pe_oracle.split_lines(f.tobytes()), pe_oracle.split_lines(r.tobytes()),
                                                table, [len(s) for s in oseqs], cfg.k + 1)
    exp = {i * n + j: c for (i, j), c in node.items()}
    exp.update({n * n + i * n + j: c for (i, j), c in short.items()})
    assert dict(zip(keys.tolist(), counts.tolist())) == exp
    for k, v in ostats.items():
        assert stats[k] == v
    # CLI: sparse files
    (tmp_path / "g.gfa").write_bytes(gfa)
    f.tofile(tmp_path / "f.fq")
    r.tofile(tmp_path / "r.fq")
    out = tmp_path / "aln"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "utils", "VStrains_PE_Inference.py"), "-g", str(tmp_path / "g.gfa"),
                        "-o", str(out), "-f", str(tmp_path / "f.fq"), "-r", str(tmp_path / "r.fq"), "-k", str(cfg.k)], capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    pe_lines = (out / "pe_info").read_text().splitlines()
    st_lines = (out / "st_info").read_text().splitlines()
    assert sorted(pe_lines) == sorted("%s:%s:%d" % (ids[i], ids[j], c) for (i, j), c in node.items())
    assert sorted(st_lines) == sorted("%s:%s:%d" % (ids[i], ids[j], c) for (i, j), c in short.items())


def test_full_size_c2_matches_c_oracle_and_invariants():
    """BASELINE.json configs[1] at full size (1 M pairs 2x250): bit-exact against the C oracle, plus
    size-independent properties: every key is accounted for, short_mat is upper triangular, the
    counters partition the pairs, two runs give identical matrices (determinism)."""
    import bench
    cfg, g, f, r = bench.make_workload("C2", 1_000_000, 0)
    gfa = g.to_gfa()
    ids, seqs = pe_inference.parse_gfa_nodes(gfa)
    with pe_inference.PEIndex(seqs, cfg.k) as ix:
        ix.count_host(f, r)
        node, short = ix.matrices()
        st = ix.stats()
        ix.reset()
        ix.count_host(f, r)
        node2, short2 = ix.matrices()
    assert np.array_equal(node, node2) and np.array_equal(short, short2)
    assert st["total_pairs"] == 1_000_000 == st["n_pairs"] + st["short_pairs"] + st["used_pairs"]
    assert int(node.sum()) + int(short.sum()) == st["n_keys"]
    assert int(np.tril(short, -1).sum()) == 0
    onode, oshort, ostats = c_oracle.run(gfa, f, r, cfg.k)
    assert np.array_equal(node.astype(np.int64), onode)
    assert np.array_equal(short.astype(np.int64), oshort)
    for k, v in ostats.items():
        assert st[k] == v


@pytest.mark.parametrize("name,block,replay", [("C3", 10_000_000, 1), ("C4", 10_000_000, 2)])
def test_full_size_blocks_invariants_and_oracle_subsample(name, block, replay):
    """BASELINE.json configs[2] (10 M pairs 2x150) and the 10 M-pair unique block of configs[3] (replayed:
    the counts must scale exactly by the replay factor) at full size on one GPU: the counters
    partition the pairs, every link key is in the matrices, short_mat is upper triangular, two runs
    are identical; the first 100 000 pairs of the same stream are bit-exact against the C oracle."""
    import bench
    cfg, g, genomes, ab = bench.make_graph(name, block * replay)
    f, r = bench.make_reads(cfg, genomes, ab, block, 0)
    gfa = g.to_gfa()
    ids, seqs = pe_inference.parse_gfa_nodes(gfa)
    with pe_inference.PEIndex(seqs, cfg.k) as ix:
        ix.count_host(f, r)
        node1, short1 = ix.matrices()
        st1 = ix.stats()
        for _ in range(replay - 1):
            ix.count_host(f, r)
        node, short = ix.matrices()
        st = ix.stats()
        ix.reset()
        for _ in range(replay):
            ix.count_host(f, r)
        node2, short2 = ix.matrices()
        assert np.array_equal(node, node2) and np.array_equal(short, short2)          # determinism
        assert np.array_equal(node, node1 * np.uint64(replay)) and np.array_equal(short, short1 * np.uint64(replay))
        assert st1["total_pairs"] == block == st1["n_pairs"] + st1["short_pairs"] + st1["used_pairs"]
        assert st["total_pairs"] == block * replay
        assert int(node.sum()) + int(short.sum()) == st["n_keys"] == replay * st1["n_keys"]
        assert int(np.tril(short, -1).sum()) == 0
        # oracle on a sub-sample of the same stream
        sub = 100_000
        fs, rs = bench.prefix_pairs(f, r, 0, sub, cfg.read_len)
        ix.reset()
        ix.count_host(fs, rs)
        sn, ss = ix.matrices()
        sst = ix.stats()
    onode, oshort, ostats = c_oracle.run(gfa, fs, rs, cfg.k)
    assert np.array_equal(sn.astype(np.int64), onode)
    assert np.array_equal(ss.astype(np.int64), oshort)
    for k, v in ostats.items():
        assert sst[k] == v


def test_many_long_node_lists_grow_the_pools():
    """reads that map to more nodes than a list record / slot holds (a graph of many nodes with a
    one-base overhang) need private list records and spill words by the million; the pools grow and the
    launch is repeated instead of failing with VSPE_ERR_LIMIT."""
    rng = np.random.default_rng(77)
    k = 31
    genome = rng.integers(0, 4, size=4000, dtype=np.uint8)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    # nodes = every (k+2)-window of the genome: consecutive nodes overlap by k+1, each holds two (k+1)-mers
    seqs = [acgt[genome[i:i + k + 2]].tobytes() for i in range(0, genome.size - k - 2)]
    gfa = b"".join(b"S\t%d\t%s\n" % (i, s) for i, s in enumerate(seqs))
    n_reads, rl = 40000, 100
    starts = rng.integers(0, genome.size - rl, size=n_reads)
    reads = [acgt[genome[s:s + rl]].tobytes() for s in starts]
    fq = _mk_fastq(reads)
    ids, node, short, stats = pe_inference.pe_inference(gfa, fq, fq, k)
    onode, oshort, ostats = c_oracle.run(gfa, fq, fq, k)
    assert np.array_equal(node.astype(np.int64), onode)
    assert np.array_equal(short.astype(np.int64), oshort)
    for kk, v in ostats.items():
        assert stats[kk] == v


@pytest.mark.skipif(not os.environ.get("VSPE_TEST_C5"), reason="opt-in (VSPE_TEST_C5=1): builds the 200 000-node graph and a C-oracle index over it")
def test_c5_stress_graph_sparse_counts_match_oracle_on_a_subsample():
    """BASELINE.json configs[4]: the 200 000-node stress graph (HBM-resident index, N*N matrices impossible: sparse
    runs).  A 20 000-pair sub-sample is bit-exact against link counts built from the C oracle's per-read node
    lists; 1 M pairs check the invariants (every key in the runs, the counters partition the pairs, two runs equal)."""
    import bench
    cfg, g, genomes, ab = bench.make_graph("C5", 1_000_000)
    assert len(g.ids) == 200_000
    f, r = bench.make_reads(cfg, genomes, ab, 1_000_000, 0)
    gfa = g.to_gfa()
    ids, seqs = pe_inference.parse_gfa_nodes(gfa)
    n = len(ids)
    sub = 20_000
    fs, rs = bench.prefix_pairs(f, r, 0, sub, cfg.read_len)
    with pe_inference.PEIndex(seqs, cfg.k) as ix:
        assert ix.is_sparse
        ix.count_host(fs, rs)
        keys, counts = ix.sparse()
        st = ix.stats()
        ix.reset()
        ix.count_host(f, r)
        k1, c1 = ix.sparse()
        st1 = ix.stats()
        ix.reset()
        ix.count_host(f, r)
        k2, c2 = ix.sparse()
    assert np.array_equal(k1, k2) and np.array_equal(c1, c2)
    assert int(c1.sum()) == st1["n_keys"] and st1["total_pairs"] == 1_000_000 == st1["n_pairs"] + st1["short_pairs"] + st1["used_pairs"]
    assert np.all(np.diff(k1.astype(np.int64)) > 0)
    # oracle: per-read node lists of both mates -> the reference's accumulation loops (PE_Inference.py:160-188)
    of, nf, sf = c_oracle.map_reads(gfa, fs, cfg.k)
    orr, nr, sr = c_oracle.map_reads(gfa, rs, cfg.k)
    exp = {}
    used = n_skip = short_skip = 0
    for p in range(sub):
        if sf[p] == 1 or sr[p] == 1:
            n_skip += 1
            continue
        if sf[p] == 2 or sr[p] == 2:
            short_skip += 1
            continue
        used += 1
        L = nf[of[p]:of[p + 1]].tolist()
        R = nr[orr[p]:orr[p + 1]].tolist()
        for lst in (L, R):
            for a in range(len(lst)):
                for b in range(a, len(lst)):
                    kk = n * n + lst[a] * n + lst[b]
                    exp[kk] = exp.get(kk, 0) + 1
        for i in L:
            for j in R:
                kk = i * n + j
                exp[kk] = exp.get(kk, 0) + 1
    assert (st["used_pairs"], st["n_pairs"], st["short_pairs"]) == (used, n_skip, short_skip)
    assert dict(zip(keys.tolist(), counts.tolist())) == exp


def test_short_first_records_do_not_push_longer_reads_to_the_exhaustive_tier():
    """the packed-row capacity is sized from samples across the input, not from its first lines only:
    a file that starts with trimmed reads and continues with full-length ones still runs the packed tiers."""
    cfg = synth.CONFIGS["C2"]
    g, f, r = synth.generate(cfg, pairs=3000)
    gfa = g.to_gfa()
    short = _mk_fastq([b"ACGTACGTACGTACGTACGTAC"] * 40)
    f2, r2 = short + f.tobytes(), short + r.tobytes()
    ids, node, short_m, stats = pe_inference.pe_inference(gfa, f2, r2, cfg.k)
    onode, oshort, ostats = c_oracle.run(gfa, f2, r2, cfg.k)
    assert np.array_equal(node.astype(np.int64), onode) and np.array_equal(short_m.astype(np.int64), oshort)
    assert stats["reads_generic"] == 0
    for k, v in ostats.items():
        assert stats[k] == v


def _records(fq):
    lines = bytes(fq).split(b"\n")
    assert lines[-1] == b""
    return [b"\n".join(lines[i:i + 4]) + b"\n" for i in range(0, len(lines) - 1, 4)]


def test_read_memo_is_exact_and_used():
    """Repeated reads (deep coverage) take their node list from the read memo; the counts are those of walking every
    read, within one call, over several calls on the same context, and after a reset (which forgets the memo)."""
    for name in ("C2", "C4"):
        cfg = synth.CONFIGS[name]
        g, f, r = synth.generate(cfg, pairs=1500)
        rf, rr = _records(f), _records(r)
        order = np.random.default_rng(5).permutation(np.repeat(np.arange(len(rf)), 40))
        f2, r2 = b"".join(rf[i] for i in order), b"".join(rr[i] for i in order)
        gfa = g.to_gfa()
        onode, oshort, ostats = c_oracle.run(gfa, f2, r2, cfg.k)
        ids, seqs = pe_inference.parse_gfa_nodes(gfa)
        for memo in (1, 0):
            with pe_inference.PEIndex(seqs, cfg.k) as ix:
                ix.set_option("memo", memo)
                ix.count_host(f2, r2)
                node, short = ix.matrices()
                st = ix.stats()
                assert np.array_equal(node.astype(np.int64), onode) and np.array_equal(short.astype(np.int64), oshort), (name, memo)
                for k, v in ostats.items():
                    assert st[k] == v, (name, memo, k)
                # a second call adds the same counts again, now with a warm memo (a round of reads asks the memo before
                # that round's walks fill it, and this input is a single round); a reset starts over
                ix.count_host(f2, r2)
                node2, short2 = ix.matrices()
                assert np.array_equal(node2.astype(np.int64), 2 * onode) and np.array_equal(short2.astype(np.int64), 2 * oshort), (name, memo)
                st2 = ix.stats()
                if memo:
                    assert st2["reads_memo"] > len(order), (name, st2["reads_memo"])     # more than half of the second call's 2 x len(order) reads
                else:
                    assert st2["reads_memo"] == 0
                ix.reset()
                ix.count_host(f2, r2)
                node3, short3 = ix.matrices()
                assert np.array_equal(node3.astype(np.int64), onode) and np.array_equal(short3.astype(np.int64), oshort), (name, memo)


def test_matrices_into_caller_buffers():
    """PEIndex.matrices(out=...) fills the caller's (e.g. pinned) arrays with the same counts."""
    cfg = synth.CONFIGS["C1"]
    g, f, r = synth.generate(cfg, pairs=2000)
    ids, seqs = pe_inference.parse_gfa_nodes(g.to_gfa())
    with pe_inference.PEIndex(seqs, cfg.k) as ix:
        ix.count_host(f, r)
        node, short = ix.matrices()
        n = len(ids)
        out = (np.full((n, n), 7, dtype=np.uint64), np.full((n, n), 7, dtype=np.uint64))
        node2, short2 = ix.matrices(out=out)
        assert node2 is out[0] and short2 is out[1]
        assert np.array_equal(node, node2) and np.array_equal(short, short2)
        with pytest.raises(VspeError):
            ix.matrices(out=(np.zeros((n, n), dtype=np.int32), out[1]))
