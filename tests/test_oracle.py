"""Pins the oracle (Python and C restatements) to the reference's own outputs.

tests/golden/*.npz were produced by oracle/make_golden.py running the unmodified reference
script; every fixture must be reproduced byte for byte."""
import os
import random
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from oracle import c_oracle, pe_oracle

REF = "/root/reference/utils/VStrains_PE_Inference.py"


def test_python_oracle_matches_reference_bytes(golden):
    if golden.status != 0:
        with pytest.raises(KeyError):
            pe_oracle.run_bytes(golden.gfa, golden.fwd, golden.rve, golden.k)
        return
    pe, st, stats, ids = pe_oracle.run_bytes(golden.gfa, golden.fwd, golden.rve, golden.k)
    assert pe == golden.pe_info
    assert st == golden.st_info
    assert stats["n_pairs"] + stats["short_pairs"] + stats["used_pairs"] == stats["total_pairs"]


def test_c_oracle_matches_reference_bytes(golden):
    if golden.status != 0:
        with pytest.raises(ValueError):
            c_oracle.run(golden.gfa, golden.fwd, golden.rve, golden.k)
        return
    ids, _ = pe_oracle.parse_gfa(golden.gfa)
    node, short, stats = c_oracle.run(golden.gfa, golden.fwd, golden.rve, golden.k)
    assert c_oracle.info_bytes(ids, node) == golden.pe_info
    assert c_oracle.info_bytes(ids, short) == golden.st_info
    _, _, pstats, _ = pe_oracle.run_bytes(golden.gfa, golden.fwd, golden.rve, golden.k)
    assert stats == pstats


def test_micro_case_is_the_hand_checked_answer():
    from conftest import Golden, GOLDEN_DIR
    g = Golden(os.path.join(GOLDEN_DIR, "micro.npz"))
    nz = lambda b: sorted(l for l in b.decode().split("\n") if l and not l.endswith(":0"))
    assert nz(g.pe_info) == ["7:-9:4"]
    assert nz(g.st_info) == ["-9:-9:5", "7:7:4"]


def test_c_map_reads_matches_python(golden):
    if golden.status != 0:
        return
    ids, seqs = pe_oracle.parse_gfa(golden.gfa)
    sl = golden.k + 1
    table = pe_oracle.build_index(seqs, sl)
    lens = [len(s) for s in seqs]
    for fq in (golden.fwd, golden.rve):
        off, nodes, status = c_oracle.map_reads(golden.gfa, fq, golden.k)
        lines = pe_oracle.split_lines(fq)
        assert len(status) == len(lines) // 4
        for r in range(len(status)):
            seq = lines[4 * r + 1][:-1]
            if "N" in seq:
                assert status[r] == 1
            elif len(seq) < sl:
                assert status[r] == 2
            else:
                assert status[r] == 0
                assert list(nodes[off[r]:off[r + 1]]) == pe_oracle.map_read(seq, table, lens, sl)


def test_integer_predicate_equals_reference_float_rule():
    """The float rule of utils/VStrains_PE_Inference.py:36-47, evaluated literally, against the
    integer restatement on an exhaustive small grid."""
    for rlen in range(4, 40):
        for sl in range(2, rlen + 1):
            for ln in range(sl, 60, 3):
                for kidx in range(0, rlen - sl + 1):
                    for v in range(1, 2 * (rlen - sl + 1) + 2, 1):
                        sat_f = min(0 + ln - 1, 0 - kidx + rlen - 1) - 0 - (sl - 1) + 1
                        exp_f = (min(rlen, ln) - sl + 1) * (rlen - sl) / rlen
                        keep_f = v >= max(min(sat_f, exp_f), 1)
                        sat = min(ln, rlen - kidx) - sl + 1
                        ab = (min(rlen, ln) - sl + 1) * (rlen - sl)
                        assert keep_f == (v >= sat or v * rlen >= ab)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
def test_live_reference_on_fresh_random_cases():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import make_golden
    for seed in (9001, 9002, 9003):
        gfa, fwd, rve, k = make_golden.lowcomplexity(seed, nl=random.Random(seed).choice(["\n", "\r\n"]))
        status, pe, st = make_golden.run_reference(gfa, fwd, rve, k)
        assert status == 0
        ope, ost, _, ids = pe_oracle.run_bytes(gfa, fwd, rve, k)
        assert (ope, ost) == (pe, st)
        node, short, _ = c_oracle.run(gfa, fwd, rve, k)
        assert c_oracle.info_bytes(ids, node) == pe and c_oracle.info_bytes(ids, short) == st
