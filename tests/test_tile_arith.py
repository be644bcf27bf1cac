"""The arithmetic k_scan_rows / k_tile_fix / k_memo share (vstrains_b200/csrc/scan_map.cu), restated in Python and
checked against a brute-force line count: which reads a tile owns, the chunk-local index of its first read, the
line-phase guess and the block -> tile table.  CPU only: it pins the formulas the kernels implement (the kernels
themselves are compared with the oracle by the GPU parity tests)."""
import numpy as np
import pytest

WK = 128


def _fastq(rng, n_rec, well_formed=True):
    out = []
    for i in range(n_rec):
        ln = int(rng.integers(1, 60))
        seq = bytes(rng.choice(list(b"ACGT"), ln).tolist())
        if well_formed:
            out.append(b"@r%d\n" % i + seq + b"\n+\n" + b"I" * ln + b"\n")
        else:                                   # decoy: quality starts with '@', sequence with '+', header without '@'
            out.append(b"r%d\n+" % i + seq + b"\n-\n@" + b"I" * ln + b"\n")
    return b"".join(out)


def _tiles(buf, tile):
    pos = np.flatnonzero(np.frombuffer(buf, dtype=np.uint8) == 10)
    n_tiles = (len(buf) + tile - 1) // tile
    return [pos[(pos >= t * tile) & (pos < (t + 1) * tile)] for t in range(n_tiles)]


def _guess(buf, terms):
    """smallest phase b without a contradiction among the lines that follow the tile's first 32 terminators"""
    viol = 0
    for i, p in enumerate(terms[:32]):
        j = int(p) + 1
        if j < len(buf):
            ch = buf[j]
            if ch != ord("@"):
                viol |= 1 << ((3 - i) & 3)
            if ch != ord("+"):
                viol |= 1 << ((1 - i) & 3)
    return 0 if viol == 0xF else [b for b in range(4) if not (viol >> b) & 1][0]


@pytest.mark.parametrize("seed", range(6))
def test_tile_ownership_prefix_and_block_table(seed):
    rng = np.random.default_rng(seed)
    whole = _fastq(rng, 400)
    lines_before = int(rng.integers(0, 4))                      # the chunk may start at any line of a record
    nl = np.flatnonzero(np.frombuffer(whole, dtype=np.uint8) == 10)
    start = 0 if lines_before == 0 else int(nl[lines_before - 1]) + 1
    buf = whole[start:]
    line_base = 4 * 7 + lines_before                            # lines of earlier chunks (any multiple of 4 plus the offset)
    rec_first = (line_base + 2) // 4                            # seq_lines_before
    tile = int(rng.choice([64, 100, 256, 1000]))
    tiles = _tiles(buf, tile)
    # brute force: sequence lines (global line index % 4 == 1) and where they start
    term = np.flatnonzero(np.frombuffer(buf, dtype=np.uint8) == 10)
    line_start = np.concatenate([[0], term + 1])[: len(term) + 1]
    seq_starts = [int(line_start[i]) for i in range(len(line_start)) if (line_base + i) % 4 == 1 and line_start[i] < len(buf)]
    run, r_first, owned = 0, [], []
    for t, tp in enumerate(tiles):
        tot = len(tp)
        base = line_base + run
        b = base & 3
        h0 = (4 - b) & 3
        n_own = (tot - h0 + 3) // 4 if tot > h0 else 0
        shift = 1 if (t == 0 and (line_base & 3) == 1) else 0
        rf = ((base + 3) >> 2) - shift - rec_first
        r_first.append(rf)
        starts = []
        for li in range(n_own + shift):
            if li < shift:
                starts.append(0)
            else:
                starts.append(int(tp[h0 + 4 * (li - shift)]) + 1)
        owned.append(starts)
        # a well-formed stream never contradicts its true phase, and the guess finds it whenever the tile shows a header or a separator
        if tot >= 4:
            assert _guess(buf, tp) == b
        run += tot
    r_first.append(((line_base + run + 3) >> 2) - rec_first)
    flat = [s for starts in owned for s in starts]
    # every sequence line that starts inside the buffer is owned exactly once, in order, and r_first is their prefix count
    assert flat == seq_starts[: len(flat)] and len(flat) >= len(seq_starts) - 0
    assert len(flat) == len(seq_starts)
    acc = 0
    for t, starts in enumerate(owned):
        assert r_first[t] == acc
        acc += len(starts)
    assert r_first[-1] == acc
    # block table: blk_tile[b] is the tile that holds read 128 b; k_memo walks forward from it
    n_reads = acc
    n_blk = (n_reads + WK - 1) // WK
    blk_tile = [None] * n_blk
    for t in range(len(tiles)):
        rf, rn = r_first[t], r_first[t + 1]
        bk = (rf + WK - 1) // WK
        while bk < n_blk and bk * WK < rn:
            blk_tile[bk] = t
            bk += 1
    for r in range(n_reads):
        t = blk_tile[r // WK]
        while r >= r_first[t + 1]:
            t += 1
        assert r_first[t] <= r < r_first[t + 1]


def test_decoy_text_defeats_the_guess_somewhere():
    """... which is why every guess is verified: on text without the FASTQ markers (or with misleading ones) some tiles
    guess wrong and must be listed for the exact second launch."""
    rng = np.random.default_rng(3)
    buf = _fastq(rng, 300, well_formed=False)
    wrong, run = 0, 0
    for t, tp in enumerate(_tiles(buf, 256)):
        if t and len(tp) and _guess(buf, tp) != (run & 3):
            wrong += 1
        run += len(tp)
    assert wrong > 0
