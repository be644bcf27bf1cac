#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference script.

Run in the build container only (needs /root/reference); the fixtures are committed so the
GPU box never needs the reference.  Each fixture holds the exact input bytes
(gfa, fwd, rve, k) and the reference's exact output bytes (pe_info, st_info) or, for inputs
the reference rejects, ``status != 0``.

    python oracle/make_golden.py            # regenerate everything
"""
from __future__ import annotations

import os
import random
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthgen as synth  # noqa: E402

REF = "/root/reference/utils/VStrains_PE_Inference.py"
OUT = os.path.join(ROOT, "tests", "golden")
_RC = {"A": "T", "C": "G", "G": "C", "T": "A"}


def rc(s: str) -> str:
    return "".join(_RC[c] for c in reversed(s))


def run_reference(gfa: bytes, fwd: bytes, rve: bytes, k: int):
    with tempfile.TemporaryDirectory() as d:
        for name, data in (("g.gfa", gfa), ("f.fq", fwd), ("r.fq", rve)):
            with open(os.path.join(d, name), "wb") as f:
                f.write(data)
        p = subprocess.run([sys.executable, REF, "-g", d + "/g.gfa", "-o", d + "/out", "-f", d + "/f.fq",
                            "-r", d + "/r.fq", "-k", str(k)], capture_output=True)
        if p.returncode != 0:
            return p.returncode, b"", b""
        with open(d + "/out/pe_info", "rb") as f:
            pe = f.read()
        with open(d + "/out/st_info", "rb") as f:
            st = f.read()
        assert sorted(os.listdir(d + "/out")) == ["pe_info", "st_info"]
        return 0, pe, st


def save(name: str, gfa: bytes, fwd: bytes, rve: bytes, k: int, note: str):
    status, pe, st = run_reference(gfa, fwd, rve, k)
    u8 = lambda b: np.frombuffer(b, dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), gfa=u8(gfa), fwd=u8(fwd), rve=u8(rve),
                        k=np.int64(k), status=np.int64(status), pe_info=u8(pe), st_info=u8(st),
                        note=np.array(note))
    print("%-28s k=%-3d status=%d gfa=%dB fq=%d+%dB out=%d+%dB" %
          (name, k, status, len(gfa), len(fwd), len(rve), len(pe), len(st)))


def fastq(seqs, mate, nl="\n", tail=""):
    out = []
    for i, s in enumerate(seqs):
        out.append("@r%d/%d%s%s%s+%s%s%s" % (i, mate, nl, s, nl, nl, "I" * len(s), nl))
    return ("".join(out) + tail).encode()


def micro():
    """SURVEY.md §8c hand-checkable case: expect pe_info 7:-9:4, st_info 7:7:4 and -9:-9:5."""
    gfa = ("H\tVN:Z:1.0\nS\t7\tACGTTGCAAGGCTTAACGGATC\tDP:f:3.5\nS\t-9\tTTTTGGGGCCCCAAAATTTT\tDP:f:1\n"
           "S\tx\tACG\nL\t7\t+\t-9\t+\t5M\n").encode()
    F = ["ACGTTGCAAGGC", "ACGTTGcAAGGC", "ACGTTG", "ACGTTGCAAGGCN", "GATCCGTTAAGC", "ACGTTGCAAGGCTTAACGGATCAAAA"]
    R = ["TTTTGGGGCCCC"] * 4 + ["AAAATTTTGGGG", "TTTTGG"]
    return gfa, fastq(F, 1), fastq(R, 2, tail="@partial/2\nACGT\n"), 5


def lowcomplexity(seed: int, nl: str = "\n", unterminated: bool = False):
    rng = random.Random(seed)
    sl = rng.choice([4, 5, 6, 8])
    alpha = rng.choice(["AC", "AT", "ACG", "ACGT", "ACGT"])
    n_nodes = rng.randint(2, 9)
    nodes = ["".join(rng.choice(alpha) for _ in range(rng.randint(1, 40))) for _ in range(n_nodes)]
    if rng.random() < 0.5:                      # force an even-length reverse palindrome
        h = "".join(rng.choice("ACGT") for _ in range(sl // 2 + 2))
        nodes.append(h + rc(h))
    ids = [str(i) for i in range(len(nodes))]
    rng.shuffle(ids)
    lines = ["S\t%s\t%s\tDP:f:%.2f" % (i, s, rng.random() * 50) for i, s in zip(ids, nodes)]
    lines += ["L\t%s\t+\t%s\t-\t%dM" % (ids[0], ids[-1], sl - 1), "# comment", ""]
    gfa = (nl.join(lines) + nl).encode()

    def read():
        cat = "".join(rng.choice(nodes) for _ in range(3))
        L = rng.randint(1, min(34, len(cat)))
        st = rng.randint(0, len(cat) - L)
        r = cat[st:st + L]
        if rng.random() < 0.5:
            r = rc(r)
        r = "".join(c if rng.random() > 0.05 else rng.choice("ACGTNacgt") for c in r)
        return r

    n = rng.randint(20, 60)
    F = [read() for _ in range(n)]
    R = [read() for _ in range(n + rng.randint(-3, 3))]
    fwd, rve = fastq(F, 1, nl), fastq(R, 2, nl)
    if unterminated:
        fwd = fwd[:-len(nl)]
    return gfa, fwd, rve, sl - 1


def synth_case(read_len: int, k: int, genome: int, strains: int, pairs: int, seed: int):
    rng = np.random.default_rng(seed)
    st = synth.make_strains(genome, strains, 0.012, rng)
    seqs, cov, links = synth.build_dbg(st, k, synth.abundances(strains), seed=seed)
    lens = np.array([len(s) for s in seqs])
    order = np.argsort(-lens, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    g = synth.Graph(k, [str(i) for i in range(len(seqs))], [seqs[i] for i in order], cov[order], rank[links])
    f, r = synth.make_reads([st], synth.abundances(strains), read_len, pairs, k, rng,
                            sub_rate=0.004, n_rate=0.02, short_rate=0.02)
    return g.to_gfa(), f.tobytes(), r.tobytes(), k


def main():
    os.makedirs(OUT, exist_ok=True)
    save("micro", *micro(), note="SURVEY 8c hand-checkable micro case (ids 7,-9,x; k=5)")
    for s in range(12):
        save("lowcx_%02d" % s, *lowcomplexity(100 + s), note="low-complexity random graph, repeats/palindromes/rc/N/lower-case")
    save("lowcx_crlf", *lowcomplexity(201, nl="\r\n"), note="CRLF line endings everywhere")
    save("lowcx_cr", *lowcomplexity(202, nl="\r"), note="lone-CR line endings everywhere")
    save("lowcx_unterminated", *lowcomplexity(203, unterminated=True), note="fwd file lacks the final newline")
    save("synth_2x250_k127", *synth_case(250, 127, 2600, 4, 250, 11), note="dBG, 4 strains x 2.6 kb, 250 pairs 2x250, k=127")
    save("synth_2x150_k77", *synth_case(150, 77, 2400, 6, 350, 12), note="dBG, 6 strains x 2.4 kb, 350 pairs 2x150, k=77")
    # inputs the reference rejects (exit status != 0)
    g, f, r, k = micro()
    save("err_lowercase_node", g.replace(b"ACGTTGCAAGGCTTAACGGATC", b"ACGTTGCAAGGCTTAACGGATc"), f, r, k,
         note="node >= split_len with a lower-case base: KeyError in reverse_seq")
    save("err_n_in_node", g.replace(b"TTTTGGGGCCCCAAAATTTT", b"TTTTGGGGNCCCAAAATTTT"), f, r, k,
         note="node >= split_len containing N: KeyError in reverse_seq")
    save("ok_bad_short_node", g.replace(b"S\tx\tACG\n", b"S\tx\tnNn\n"), f, r, k,
         note="non-ACGT in a node shorter than split_len is accepted")
    save("empty_reads", g, b"", b"", k, note="empty FASTQ files: all-zero matrices")
    save("empty_graph", b"H\tVN:Z:1.0\n", f, r, k, note="no S lines: empty outputs")


if __name__ == "__main__":
    main()
