"""Deterministic synthetic quasispecies graphs + paired-end reads.

Test / bench tooling only (not part of the product package): this module produces the *inputs* the
hot path consumes,
in the formats the reference reads:

* a ``s_graph_L1.gfa``-dialect assembly graph (``S\\t<id>\\t<SEQ>\\tDP:f:<cov>`` then
  ``L\\t<u>\\t+\\t<v>\\t+\\t<k>M``) as written by the reference's ``graph_to_gfa``
  (reference ``utils/VStrains_IO.py:337-372``) and parsed by the hot path at
  ``utils/VStrains_PE_Inference.py:101-112``;
* two FASTQ files (4-line records, ``@p%09d/1`` headers, ``+``, ``I`` qualities) as read
  by ``utils/VStrains_PE_Inference.py:147-159``.

The graph is the edge-centric compacted de Bruijn graph of the strains' forward strands:
every distinct (k+1)-mer is an edge, segments are maximal non-branching edge paths,
adjacent segments overlap by exactly k bases and every segment is >= k+1 long -- the same
shape SPAdes hands to VStrains.

Everything is vectorised numpy so that the 1M-pair config builds in seconds on the
GPU box's host.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, List, Optional, Tuple

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)  # A<->T, C<->G on codes 0..3


@dataclasses.dataclass(frozen=True)
class Config:
    name: str
    genome_len: int
    strains: int
    divergence: float
    read_len: int
    pairs: int
    seed: int
    n_genomes: int = 1  # C5: many independent genomes
    target_nodes: Optional[int] = None  # C5: trim/pad to exactly this many segments

    @property
    def k(self) -> int:
        return 127 if self.read_len >= 250 else 77


# BASELINE.json configs: (G, S, d, rl, pairs, seed)
CONFIGS: Dict[str, Config] = {
    "C1": Config("C1-hiv5-2x250-100k", 9_700, 5, 0.01, 250, 100_000, 1001),
    "C2": Config("C2-polio6-2x250-1M", 7_500, 6, 0.01, 250, 1_000_000, 1002),
    "C3": Config("C3-zikv15-2x150-10M", 10_800, 15, 0.01, 150, 10_000_000, 1003),
    "C4": Config("C4-sars10-2x150-50M", 30_000, 10, 0.005, 150, 50_000_000, 1004),
    "C5": Config("C5-stress200k-2x150-100M", 10_000, 4, 0.01, 150, 100_000_000, 1005,
                 n_genomes=2_000, target_nodes=200_000),
}


@dataclasses.dataclass
class Graph:
    k: int
    ids: List[str]          # segment ids in file order (== matrix index order)
    seqs: List[bytes]       # upper-case ACGT
    cov: np.ndarray         # DP:f: value per segment
    links: np.ndarray       # [n_links, 2] (u, v) segment indices, '+' '+' orientation

    def to_gfa(self) -> bytes:
        out = []
        for n, (i, s) in enumerate(zip(self.ids, self.seqs)):
            out.append(b"S\t%s\t%s\tDP:f:%.6f\n" % (i.encode(), s, self.cov[n]))
        for u, v in self.links:
            out.append(b"L\t%s\t+\t%s\t+\t%dM\n" % (self.ids[u].encode(), self.ids[v].encode(), self.k))
        return b"".join(out)

    def to_paths(self, min_len: int = 0) -> bytes:
        """SPAdes-style ``contigs.paths`` (format parsed at reference
        ``utils/VStrains_IO.py:447-471``): one single-segment contig per segment."""
        out = []
        for n, (i, s) in enumerate(zip(self.ids, self.seqs), start=1):
            if len(s) < min_len:
                continue
            name = "NODE_%d_length_%d_cov_%.6f" % (n, len(s), self.cov[n - 1])
            out.append("%s\n%s+\n%s'\n%s-\n" % (name, i, name, i))
        return "".join(out).encode()


# ----------------------------------------------------------------------------------------
# strains
# ----------------------------------------------------------------------------------------

def make_strains(genome_len: int, n_strains: int, divergence: float,
                 rng: np.random.Generator) -> np.ndarray:
    """[S, G] uint8 codes 0..3: base genome i.i.d. uniform, each strain = base + independent
    SNPs at rate ``divergence`` (no indels)."""
    base = rng.integers(0, 4, size=genome_len, dtype=np.uint8)
    strains = np.repeat(base[None, :], n_strains, axis=0)
    for s in range(n_strains):
        sites = np.nonzero(rng.random(genome_len) < divergence)[0]
        shift = rng.integers(1, 4, size=sites.size, dtype=np.uint8)
        strains[s, sites] = (strains[s, sites] + shift) & 3
    return strains


def abundances(n_strains: int) -> np.ndarray:
    a = 0.5 ** np.arange(n_strains, dtype=np.float64)
    return a / a.sum()


# ----------------------------------------------------------------------------------------
# compacted de Bruijn graph
# ----------------------------------------------------------------------------------------

_HASH_B = np.uint64(0x9E3779B97F4A7C15)  # odd => invertible mod 2^64


def _inv64(b: int) -> int:
    x = b  # Newton iteration for the inverse of an odd number mod 2^64
    for _ in range(6):
        x = (x * (2 - b * x)) & 0xFFFFFFFFFFFFFFFF
    return x


def _window_hashes(codes: np.ndarray, w: int, sym: np.ndarray) -> np.ndarray:
    """Polynomial hash (mod 2^64) of every length-``w`` window of ``codes``."""
    n = codes.size
    if n < w:
        return np.zeros(0, dtype=np.uint64)
    with np.errstate(over="ignore"):
        binv = np.uint64(_inv64(int(_HASH_B)))
        pw_inv = np.cumprod(np.concatenate([[np.uint64(1)], np.full(n - 1, binv, dtype=np.uint64)]))
        pw = np.cumprod(np.concatenate([[np.uint64(1)], np.full(n - 1, _HASH_B, dtype=np.uint64)]))
        q = np.cumsum(sym[codes] * pw_inv)
        q = np.concatenate([[np.uint64(0)], q])
        # sum_{j in [p, p+w)} sym[c_j] * B^{-j}, scaled by B^{p+w-1}
        return (q[w:] - q[:-w]) * pw[w - 1:]


def build_dbg(strains: np.ndarray, k: int, weights: Optional[np.ndarray] = None,
              seed: int = 0) -> Tuple[List[bytes], np.ndarray, np.ndarray]:
    """Compacted dBG over the rows of ``strains`` (forward strands only).

    Returns (segment sequences, coverage per segment, links[n,2]).  Segment order is the
    discovery order; the caller renumbers."""
    S, G = strains.shape
    if weights is None:
        weights = np.ones(S)
    sym = np.random.default_rng(seed ^ 0x5EED).integers(1, 2**63, size=4, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    n_e = G - k          # (k+1)-mers per strain
    n_v = G - k + 1      # k-mers per strain
    if n_e <= 0:
        return [], np.zeros(0), np.zeros((0, 2), dtype=np.int64)
    he = np.stack([_window_hashes(strains[s], k + 1, sym) for s in range(S)])  # [S, n_e]
    hv = np.stack([_window_hashes(strains[s], k, sym) for s in range(S)])      # [S, n_v]
    flat_e = he.ravel()
    ue, first, inv = np.unique(flat_e, return_index=True, return_inverse=True)
    uid = inv.reshape(S, n_e)
    pre_u = hv[:, :-1].ravel()[first]
    suf_u = hv[:, 1:].ravel()[first]
    # degrees of k-mer vertices counted over distinct edges
    vo, co = np.unique(pre_u, return_counts=True)
    vi, ci = np.unique(suf_u, return_counts=True)

    def deg(tab_v, tab_c, x):
        j = np.searchsorted(tab_v, x)
        j = np.minimum(j, tab_v.size - 1)
        return np.where(tab_v[j] == x, tab_c[j], 0)

    # link between edge p and p+1 of a strain passes through vertex hv[:, p+1]
    mid = hv[:, 1:-1]                                    # [S, n_e-1]
    ok = (deg(vo, co, mid.ravel()) == 1) & (deg(vi, ci, mid.ravel()) == 1)
    ok = ok.reshape(S, n_e - 1)
    start = np.ones((S, n_e), dtype=bool)
    start[:, 1:] = ~ok
    # run lengths along each strain
    s_idx, p_idx = np.nonzero(start)
    nxt = np.concatenate([p_idx[1:], [0]])
    last_of_strain = np.concatenate([s_idx[1:] != s_idx[:-1], [True]])
    run_len = np.where(last_of_strain, n_e - p_idx, nxt - p_idx)
    run_uid = uid[s_idx, p_idx]
    _, keep = np.unique(run_uid, return_index=True)
    keep.sort()
    s_idx, p_idx, run_len = s_idx[keep], p_idx[keep], run_len[keep]
    seqs = [_ACGT[strains[s, p:p + n + k]].tobytes() for s, p, n in zip(s_idx, p_idx, run_len)]
    # coverage: abundance-weighted multiplicity of the first edge
    wsum = np.bincount(inv, weights=np.repeat(weights, n_e), minlength=ue.size)
    cov = wsum[uid[s_idx, p_idx]]
    # links: suffix vertex of last edge == prefix vertex of first edge
    first_pre = hv[s_idx, p_idx]
    last_suf = hv[s_idx, p_idx + run_len]
    order = np.argsort(first_pre, kind="stable")
    fp_sorted = first_pre[order]
    lo = np.searchsorted(fp_sorted, last_suf, side="left")
    hi = np.searchsorted(fp_sorted, last_suf, side="right")
    cnt = hi - lo
    u = np.repeat(np.arange(len(seqs)), cnt)
    off = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    v = order[np.repeat(lo, cnt) + off]
    return seqs, cov, np.stack([u, v], axis=1).astype(np.int64) if u.size else np.zeros((0, 2), dtype=np.int64)


def _one_genome(args):
    cfg, g, depth = args
    grng = np.random.default_rng([cfg.seed, 104729, g])
    st = make_strains(cfg.genome_len, cfg.strains, cfg.divergence, grng)
    seqs, cov, links = build_dbg(st, cfg.k, abundances(cfg.strains) * depth, seed=cfg.seed + g)
    return st, seqs, cov, links


def make_graph(cfg: Config, rng: np.random.Generator, depth: float = 1.0):
    """Returns (Graph, list of strain arrays per genome, abundance per strain).  Multi-genome configs (the
    200 000-node stress graph: 2 000 independent genomes) draw every genome from its own random stream, so
    they can be built by a pool of worker processes."""
    all_seqs: List[bytes] = []
    all_cov: List[np.ndarray] = []
    all_links: List[np.ndarray] = []
    genomes = []
    ab = abundances(cfg.strains)
    if cfg.n_genomes > 1:
        jobs = [(cfg, g, depth) for g in range(cfg.n_genomes)]
        n_proc = min(len(os.sched_getaffinity(0)), 32, cfg.n_genomes)
        if n_proc > 1 and cfg.n_genomes >= 64:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(n_proc) as pool:
                parts = pool.map(_one_genome, jobs, chunksize=max(1, cfg.n_genomes // (4 * n_proc)))
        else:
            parts = [_one_genome(j) for j in jobs]
        for st, seqs, cov, links in parts:
            genomes.append(st)
            all_links.append(links + len(all_seqs))
            all_seqs.extend(seqs)
            all_cov.append(cov)
    else:
        st = make_strains(cfg.genome_len, cfg.strains, cfg.divergence, rng)
        genomes.append(st)
        seqs, cov, links = build_dbg(st, cfg.k, ab * depth, seed=cfg.seed)
        all_links.append(links)
        all_seqs.extend(seqs)
        all_cov.append(cov)
    cov = np.concatenate(all_cov) if all_cov else np.zeros(0)
    links = np.concatenate(all_links) if all_links else np.zeros((0, 2), dtype=np.int64)
    n = len(all_seqs)
    if cfg.target_nodes is not None and n != cfg.target_nodes:
        if n > cfg.target_nodes:                      # trim: drop the shortest segments
            lens = np.array([len(s) for s in all_seqs])
            keep = np.sort(np.argsort(-lens, kind="stable")[:cfg.target_nodes])
            remap = -np.ones(n, dtype=np.int64)
            remap[keep] = np.arange(keep.size)
            all_seqs = [all_seqs[i] for i in keep]
            cov = cov[keep]
            links = remap[links]
            links = links[(links >= 0).all(axis=1)]
        else:                                         # pad: unrelated random segments
            extra = cfg.target_nodes - n
            for _ in range(extra):
                all_seqs.append(_ACGT[rng.integers(0, 4, size=cfg.k + 1 + 40, dtype=np.uint8)].tobytes())
            cov = np.concatenate([cov, np.full(extra, 0.01)])
    lens = np.array([len(s) for s in all_seqs], dtype=np.int64)
    order = np.argsort(-lens, kind="stable")          # ids 0..N-1 in descending-length order
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    g = Graph(cfg.k, [str(i) for i in range(order.size)], [all_seqs[i] for i in order],
              cov[order], rank[links] if links.size else links)
    return g, genomes, ab


# ----------------------------------------------------------------------------------------
# reads
# ----------------------------------------------------------------------------------------

def _fastq_bytes(bases: np.ndarray, lens: np.ndarray, first_idx: int, mate: int) -> np.ndarray:
    """bases [n, rl] ASCII, lens [n] (<= rl) -> flat uint8 FASTQ with records
    ``@p%09d/<mate>\\n<seq>\\n+\\n<I*len>\\n`` (2*len + 18 bytes each)."""
    n, rl = bases.shape
    rec = 14 + (rl + 1) + 2 + (rl + 1)
    buf = np.empty((n, rec), dtype=np.uint8)
    keep = np.ones((n, rec), dtype=bool)
    idx = np.arange(first_idx, first_idx + n, dtype=np.int64)
    buf[:, 0] = ord("@")
    buf[:, 1] = ord("p")
    for d in range(9):
        buf[:, 2 + d] = ord("0") + (idx // 10 ** (8 - d)) % 10
    buf[:, 11] = ord("/")
    buf[:, 12] = ord("0") + mate
    buf[:, 13] = 10
    buf[:, 14:14 + rl] = bases
    buf[:, 14 + rl] = 10
    buf[:, 15 + rl] = ord("+")
    buf[:, 16 + rl] = 10
    buf[:, 17 + rl:17 + 2 * rl] = ord("I")
    buf[:, 17 + 2 * rl] = 10
    trunc = np.nonzero(lens < rl)[0]
    if trunc.size:
        col = np.arange(rl)[None, :] >= lens[trunc, None]
        keep[trunc, 14:14 + rl] = ~col
        keep[trunc, 17 + rl:17 + 2 * rl] = ~col
        return buf[keep]
    return buf.ravel()


def make_reads(genomes: List[np.ndarray], ab: np.ndarray, read_len: int, pairs: int, k: int,
               rng: np.random.Generator, first_idx: int = 0, sub_rate: float = 0.001,
               n_rate: float = 0.005, short_rate: float = 0.005) -> Tuple[np.ndarray, np.ndarray]:
    """SURVEY.md §8d read model.  Returns (fwd.fastq bytes, rve.fastq bytes) as uint8 arrays."""
    rl = read_len
    S, G = genomes[0].shape
    stack = np.stack(genomes)                                   # [n_genomes, S, G]
    flat = stack.reshape(-1)
    gsel = rng.integers(0, len(genomes), size=pairs)
    ssel = rng.choice(S, size=pairs, p=ab)
    ins = np.clip(np.rint(rng.normal(2 * rl + 100, 30, size=pairs)).astype(np.int64), rl, G)
    start = (rng.random(pairs) * (G - ins + 1)).astype(np.int64)
    base_off = (gsel * S + ssel) * G
    ar = np.arange(rl, dtype=np.int64)
    m1 = flat[(base_off + start)[:, None] + ar[None, :]]                      # frag[:rl]
    m2 = _COMP[flat[(base_off + start + ins - 1)[:, None] - ar[None, :]]]     # revcomp(frag)[:rl]
    swap = rng.random(pairs) < 0.5
    f = np.where(swap[:, None], m2, m1)
    r = np.where(swap[:, None], m1, m2)
    del m1, m2
    out = []
    for mate, codes in ((1, f), (2, r)):
        n_sub = rng.binomial(codes.size, sub_rate)
        pos = rng.integers(0, codes.size, size=n_sub)
        cf = codes.reshape(-1)
        cf[pos] = (cf[pos] + rng.integers(1, 4, size=n_sub, dtype=np.uint8)) & 3
        out.append(_ACGT[codes])
    f, r = out
    # 0.5 % of pairs get one N (in a random mate)
    npair = np.nonzero(rng.random(pairs) < n_rate)[0]
    which = rng.random(npair.size) < 0.5
    col = rng.integers(0, rl, size=npair.size)
    f[npair[which], col[which]] = ord("N")
    r[npair[~which], col[~which]] = ord("N")
    # 0.5 % of mates truncated to < k+1
    lens = []
    for _ in range(2):
        ln = np.full(pairs, rl, dtype=np.int64)
        t = np.nonzero(rng.random(pairs) < short_rate)[0]
        ln[t] = rng.integers(1, k + 1, size=t.size)
        lens.append(ln)
    return _fastq_bytes(f, lens[0], first_idx, 1), _fastq_bytes(r, lens[1], first_idx, 2)


# ----------------------------------------------------------------------------------------
# whole configs
# ----------------------------------------------------------------------------------------

def generate(cfg: Config, pairs: Optional[int] = None, block: int = 250_000):
    """Build (Graph, fwd_bytes, rve_bytes) for a config; ``pairs`` overrides the config's
    pair count (parity tests use small prefixes of the same stream)."""
    rng = np.random.default_rng(cfg.seed)
    n_pairs = cfg.pairs if pairs is None else pairs
    depth = n_pairs * 2.0 * cfg.read_len / cfg.genome_len / max(cfg.n_genomes, 1)
    g, genomes, ab = make_graph(cfg, rng, depth)
    fs, rs = [], []
    done = 0
    while done < n_pairs:
        n = min(block, n_pairs - done)
        f, r = make_reads(genomes, ab, cfg.read_len, n, cfg.k, rng, first_idx=done)
        fs.append(f)
        rs.append(r)
        done += n
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.uint8)
    return g, cat(fs), cat(rs)


def write_dataset(cfg: Config, out_dir: str, pairs: Optional[int] = None) -> Dict[str, str]:
    os.makedirs(out_dir, exist_ok=True)
    g, f, r = generate(cfg, pairs)
    paths = {k: os.path.join(out_dir, v) for k, v in
             (("gfa", "s_graph_L1.gfa"), ("fwd", "fwd.fastq"), ("rve", "rve.fastq"),
              ("paths", "contigs.paths"))}
    with open(paths["gfa"], "wb") as fh:
        fh.write(g.to_gfa())
    f.tofile(paths["fwd"])
    r.tofile(paths["rve"])
    with open(paths["paths"], "wb") as fh:
        fh.write(g.to_paths())
    return paths


# ----------------------------------------------------------------------------------------
# fast generator (C, OpenMP): bench-size inputs in seconds
# ----------------------------------------------------------------------------------------
_FQ = None


def _fq_lib():
    """libfastqgen.so (synthgen/fastq_gen.c), built on first use."""
    global _FQ
    if _FQ is None:
        import ctypes
        import subprocess
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "libfastqgen.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", here])
        _FQ = ctypes.CDLL(path)
    return _FQ


def make_reads_fast(genomes: List[np.ndarray], ab: np.ndarray, read_len: int, pairs: int, k: int, seed: int,
                    first_idx: int = 0, sub_rate: float = 0.001, n_rate: float = 0.005, short_rate: float = 0.005,
                    out: Optional[Tuple[np.ndarray, np.ndarray]] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Same read model as :func:`make_reads` from a counter-based random stream keyed by
    (seed, pair index): pairs [first_idx, first_idx + pairs) are the same bytes whatever range they
    are generated in.  ``out``: optional (fwd, rve) uint8 buffers of at least
    ``pairs * (2 * read_len + 18)`` bytes (e.g. pinned host memory); views of the filled parts are
    returned."""
    import ctypes

    class P(ctypes.Structure):
        _fields_ = [("genomes", ctypes.c_void_p), ("n_genomes", ctypes.c_uint64), ("strains", ctypes.c_uint64),
                    ("G", ctypes.c_uint64), ("ab_cdf", ctypes.c_void_p), ("read_len", ctypes.c_uint32), ("k", ctypes.c_uint32),
                    ("sub_rate", ctypes.c_double), ("n_rate", ctypes.c_double), ("short_rate", ctypes.c_double),
                    ("seed", ctypes.c_uint64)]
    stack = np.ascontiguousarray(np.stack(genomes).astype(np.uint8))            # [n_genomes, S, G]
    cdf = np.ascontiguousarray(np.cumsum(np.asarray(ab, dtype=np.float64)))
    cdf[-1] = 1.0
    cap = pairs * (2 * read_len + 18)
    if out is None:
        out = (np.empty(cap, dtype=np.uint8), np.empty(cap, dtype=np.uint8))
    f, r = out
    assert f.size >= cap and r.size >= cap and f.dtype == np.uint8 and r.dtype == np.uint8
    prm = P(stack.ctypes.data, stack.shape[0], stack.shape[1], stack.shape[2], cdf.ctypes.data, read_len, k,
            sub_rate, n_rate, short_rate, seed)
    nf, nr = ctypes.c_uint64(), ctypes.c_uint64()
    lib = _fq_lib()
    lib.fq_generate.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.c_void_p]
    rc = lib.fq_generate(ctypes.byref(prm), first_idx, pairs, f.ctypes.data, r.ctypes.data, ctypes.byref(nf), ctypes.byref(nr))
    if rc != 0:
        raise ValueError("fq_generate: bad parameters")
    return f[: nf.value], r[: nr.value]
