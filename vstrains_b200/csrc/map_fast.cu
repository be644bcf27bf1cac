// map_fast.cu -- K2 + K4, seed-and-extend tier.
//
// Replaces single_end_read_mapping (reference utils/VStrains_PE_Inference.py:16-48) for reads
// whose result it can PROVE; every other read goes to the exhaustive tier (map_generic.cu).
//
// The reference looks up all rlen-split_len+1 windows of a read.  Here a window is looked up
// only when nothing is known about it:
//   * seed:    hash + probe one window; accept only a verified posting whose uniq bit is set
//              (=> that posting is the window's entire postings multiset);
//   * extend:  compare the read against the packed node text 32 bases per step; every further
//              base that matches proves the next window equals the next text window, and its
//              uniq bit proves it has no other posting => one more hit for the same node,
//              no table access;
//   * walk:    at the end of a node strand the successor table (built from the index itself)
//              gives the unique window that continues with the read's next base;
//   * bail:    a verified posting without the uniq bit (repeat / palindrome), a non-ACGT
//              character, more than MAXN distinct nodes or a read longer than the packed
//              capacity sends the read to the exhaustive tier.
// A sequencing error breaks the chain.  Pass A walks the read left to right until its first
// unknown window uA; pass B walks the REVERSE COMPLEMENT of the read (same code, same tables:
// both strands of every node are indexed) from the other end down to its first unknown
// window.  What is left in between -- normally exactly the windows covering the erroneous
// base -- is probed by the whole warp cooperatively (phase 3); all of them must miss, as they
// do for the reference, otherwise the read goes to the exhaustive tier.
//
// Phase 1 (K2): half-warps pack reads (and their reverse complements) to 2 bits/base in
//               shared memory with coalesced loads, flagging 'N' (upper case:
//               PE_Inference.py:160) and other non-ACGT bytes.
// Phase 2 (K4): one thread per read, passes A and B.
// Phase 3:      warp-cooperative confirmation probes of the unknown windows (one range at a
//               time, 32 windows per step).
// Phase 4:      one thread per read: sort by node index, saturation predicate, write the slot.
#include <algorithm>

#include "map_common.cuh"

namespace vspe {

static constexpr int MF_THREADS = 128;
static constexpr int MAXN = 16;

// ---------------------------------------------------------------------------------------------
// Per-read node list of the walk kernels, in REGISTERS: up to FL_MAX entries (node << 32 | v | kmin << 16),
// empty = all ones.  Appends are predicated register writes; one fixed 12-exchange network sorts
// the list by node index at the end, so duplicates become neighbours and the output rank of a kept
// node is a popcount -- no data-dependent loops, every thread of the warp runs the same code.
// ---------------------------------------------------------------------------------------------
static constexpr int FL_MAX = 6;
static constexpr uint64_t FL_EMPTY = ~0ull;

struct FlatList {
    uint64_t e[FL_MAX];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < FL_MAX; i++) e[i] = FL_EMPTY;
    }
    __device__ __forceinline__ void set(uint32_t at, uint32_t node, uint32_t vk) {
        const uint64_t x = ((uint64_t)node << 32) | vk;
#pragma unroll
        for (int i = 0; i < FL_MAX; i++) if (at == (uint32_t)i) e[i] = x;
    }
    __device__ __forceinline__ void cex(int i, int j) {
        const uint64_t a = e[i], b = e[j];
        e[i] = a < b ? a : b;
        e[j] = a < b ? b : a;
    }
    __device__ __forceinline__ void sort() {
        cex(0, 5); cex(1, 3); cex(2, 4);
        cex(1, 2); cex(3, 4);
        cex(0, 3); cex(2, 5);
        cex(0, 1); cex(2, 3); cex(4, 5);
        cex(1, 2); cex(3, 4);
    }
};

// saturation predicate over a sorted list; writes the kept node indices in ascending order
__device__ __forceinline__ uint32_t flat_finalize(const FlatList& fl, const IndexView& ix, uint32_t rlen, uint32_t L, ReadSlot* out) {
    uint32_t keepmask = 0;
#pragma unroll
    for (int i = 0; i < FL_MAX; i++) {
        if (fl.e[i] != FL_EMPTY) {
            const uint32_t node = (uint32_t)(fl.e[i] >> 32), vk = (uint32_t)fl.e[i];
            if (keep_node_f(vk & 0xFFFF, vk >> 16, __ldg(ix.node_len + node), rlen, L)) keepmask |= 1u << i;
        }
    }
#pragma unroll
    for (int i = 0; i < FL_MAX; i++)
        if ((keepmask >> i) & 1) out->ids[__popc(keepmask & ((1u << i) - 1))] = (uint32_t)(fl.e[i] >> 32);
    return (uint32_t)__popc(keepmask);
}

// number of equal bases of read[rb..] and text[tb..], at most max_ext
__device__ __forceinline__ uint32_t match_len(const uint32_t* row, uint32_t rb, const uint64_t* __restrict__ text,
                                              uint32_t tb, uint32_t max_ext) {
    uint32_t done = 0;
    while (done < max_ext) {
        uint64_t x = read64(row, rb + done) ^ extract64(text, (uint64_t)tb + done);
        if (x) {
            done += (uint32_t)(__ffsll((long long)x) - 1) >> 1;
            break;
        }
        done += 32;
    }
    return min(done, max_ext);
}

// length of the run of set uniq bits starting at text position p, at most n
__device__ __forceinline__ uint32_t uniq_run(const uint32_t* __restrict__ uniq, uint32_t p, uint32_t n) {
    uint32_t done = 0;
    while (done < n) {
        uint32_t q = p + done;
        uint32_t w = ~(__ldg(uniq + (q >> 5)) >> (q & 31));      // zero bits become ones
        uint32_t avail = 32 - (q & 31);
        if (avail < 32) w |= ~0u << avail;                        // bits beyond this word stop the run
        uint32_t run = w ? (uint32_t)(__ffs((int)w) - 1) : 32u;
        done += run;
        if (run < avail) break;
    }
    return min(done, n);
}

// add `hits` windows of `node` (smallest read position kminc) to the node list of read t
__device__ __forceinline__ bool list_add(uint32_t (*s_node)[MF_THREADS], uint32_t (*s_vk)[MF_THREADS], uint32_t t,
                                         uint32_t& nn, uint32_t node, uint32_t hits, uint32_t kminc) {
    uint32_t a = 0;
    for (; a < nn; a++) if (s_node[a][t] == node) break;
    if (a == nn) {
        if (nn == MAXN) return false;
        s_node[a][t] = node;
        s_vk[a][t] = hits | (kminc << 16);
        nn++;
    } else {
        const uint32_t old = s_vk[a][t];
        s_vk[a][t] = ((old & 0xFFFF) + hits) | (min(old >> 16, kminc) << 16);
    }
    return true;
}

// Load the packed row of read r (row_words 16-base words, 16-byte aligned) with independent
// 128-bit loads and store it to this thread's shared-memory row; optionally also its reverse
// complement (reverse the 2-bit groups of every word, complement, realign by the padding).
template <int STRIDE, bool WITH_RC>
__device__ __forceinline__ void load_row(const uint32_t* __restrict__ rows, uint64_t r, uint32_t row_words, uint32_t rlen,
                                         uint32_t* row, uint32_t* rrow) {
    constexpr int NW = STRIDE - 3;                             // data words a row can hold
    constexpr int XW = (NW + 3) / 4 * 4;
    uint32_t x[XW];
    const uint4* src = reinterpret_cast<const uint4*>(rows + r * row_words);
#pragma unroll
    for (int q = 0; q < XW / 4; q++) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if ((uint32_t)(4 * q) < row_words) v = __ldg(src + q);
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int w = 0; w < NW; w++) row[w] = x[w];
    row[NW] = 0; row[NW + 1] = 0; row[NW + 2] = 0;
    if (WITH_RC) {
        const uint32_t nwords = (rlen + 15) >> 4, pad = 16 * nwords - rlen;
        // rc word j = funnel(RV[nwords-1-j], RV[nwords-2-j], 2*pad) with RV[k] = revcomp16(x[k])
        uint32_t prev = 0;                                     // RV[nwords-1-j] of the previous (lower) step
#pragma unroll
        for (int k = 0; k < NW; k++) {
            // walk k downwards from NW-1: produce rc words in increasing j only for k < nwords
            const int kk = NW - 1 - k;
            uint32_t rv = __brev(x[kk]);
            rv = (((rv & 0x55555555u) << 1) | ((rv >> 1) & 0x55555555u)) ^ 0xAAAAAAAAu;
            if ((uint32_t)kk >= nwords) { continue; }
            // RV[kk] is "a0" of rc word j = nwords-1-kk and "a1" of rc word j-1
            const int j = (int)nwords - 1 - kk;
            if (j > 0) rrow[j - 1] = __funnelshift_r(prev, rv, 2 * pad);
            prev = rv;
        }
        if (nwords) rrow[nwords - 1] = __funnelshift_r(prev, 0u, 2 * pad);
        for (uint32_t w = nwords; w < (uint32_t)STRIDE; w++) rrow[w] = 0;
    }
}

// Unknown window ranges of one read (forward-read coordinates), at most two; a third one sends
// the read to the exhaustive tier.
struct Unk {
    uint32_t r0 = 0, r1 = 0;         // from | count << 16 (count 0 = unused)
    __device__ __forceinline__ bool add(uint32_t from, uint32_t cnt) {
        if (cnt == 0) return true;
        const uint32_t v = from | (cnt << 16);
        if (!(r0 >> 16)) { r0 = v; return true; }
        if (!(r1 >> 16)) { r1 = v; return true; }
        return false;
    }
};

// One directional pass over windows [0, limit) of a packed row.  Returns the first window whose
// status is unknown (== limit when everything was resolved or recorded in `unk`).  mirror: the
// row is the reverse complement, so window i of the row is window npos-1-i of the read.
//
// A mismatch at read base e inside a node strand is a sequencing error (a variant present in the
// graph ends the strand instead).  The windows covering e equal the text windows on the same
// diagonal except for that base, so
//   * if the substitution-hit bit of (text base, read base) is clear, they all miss -- no probes;
//   * otherwise they are recorded for the cooperative probes of phase 3;
// and the pass RESUMES at window e+1 on the same diagonal after one direct comparison.
__device__ __forceinline__ uint32_t run_pass(const IndexView& ix, const uint32_t* row, uint32_t rlen, uint32_t limit,
                                             bool mirror, uint32_t npos, uint32_t (*s_node)[MF_THREADS],
                                             uint32_t (*s_vk)[MF_THREADS], uint32_t t, uint32_t& nn, Unk& unk, bool& bail) {
    const uint32_t L = ix.split_len;
    uint32_t i = 0;
    uint32_t tp = NONE32, node = 0;
    while (i < limit) {
        if (tp == NONE32) {
            const int res = probe_window(ix, row, i, tp, node);
            if (res == PROBE_MULTI) { bail = true; return limit; }
            if (res == PROBE_MISS) return i + 1;              // window i is a confirmed miss
        }
        // ---- extend along the node strand ----
        const uint32_t s1 = __ldg(ix.strand_start + 2 * node + 1);
        const bool rcs = tp >= s1;
        const uint32_t q = 2 * node + (rcs ? 1u : 0u);
        const uint32_t send = rcs ? __ldg(ix.strand_start + 2 * node + 2) : s1;
        const uint32_t room_t = send - (tp + L), room_r = rlen - (i + L);
        const uint32_t max_ext = min(room_t, room_r);
        const uint32_t ext = match_len(row, i + L, ix.text, tp + L, max_ext);
        if (ext) {
            if (uniq_run(ix.uniq, tp + 1, ext) < ext) { bail = true; return limit; }   // repeat inside the match
        }
        uint32_t hits = 1 + ext;
        if (i + hits > limit) hits = limit - i;               // pass B must not re-count pass A's windows
        const uint32_t kminc = mirror ? npos - i - hits : i;
        if (!list_add(s_node, s_vk, t, nn, node, hits, kminc)) { bail = true; return limit; }
        const uint32_t i_next = i + hits;                     // first window not proven yet
        if (i_next >= limit) return limit;
        if (ext == max_ext && ext == room_t) {
            // the strand ended exactly here: successor for the read's next base, if unique
            const uint32_t nb = i_next + L - 1;
            const uint32_t b = (row[nb >> 4] >> ((nb & 15) * 2)) & 3u;
            const uint2 sc = __ldg(reinterpret_cast<const uint2*>(ix.succ) + 4 * q + b);
            tp = sc.x;
            node = sc.y;
            i = i_next;
            continue;
        }
        // ---- mismatch at read base e = i + L + ext (text base tp + L + ext), inside the strand ----
        const uint32_t e = i + L + ext, te = tp + L + ext;
        const uint32_t rb = (row[e >> 4] >> ((e & 15) * 2)) & 3u;
        const bool clear = ix.subst && !((__ldg(ix.subst + (te >> 3)) >> (4 * (te & 7) + rb)) & 1u);
        const uint32_t i_res = e + 1, t_res = te + 1;         // window right after the error, same diagonal
        if (i_res + L <= rlen) {
            // a whole window fits after the error: resume there if it is the unique text window
            const bool ok = (uint64_t)t_res + L <= send && read_equals_text(row, i_res, ix.text, t_res, L) &&
                            ((__ldg(ix.uniq + (t_res >> 5)) >> (t_res & 31)) & 1u);
            if (!ok) return i_next;                           // second error / strand end: other pass, then phase 3
            // unknown windows [i_next, e]: all cover e and all have an in-strand text window
            const uint32_t hi = min(i_res, limit);
            if (!clear) {
                const uint32_t from = mirror ? npos - hi : i_next, cnt = hi - i_next;
                if (!unk.add(from, cnt)) { bail = true; return limit; }
            }
            if (i_res >= limit) return limit;
            i = i_res;
            tp = t_res;
            continue;
        }
        // no window starts after the error: the rest [i_next, npos) all cover e
        {
            const bool in_strand = room_t >= room_r;          // every remaining window has an in-strand text window
            // ... and equals it except for base e only if the read's tail matches the text too
            const uint32_t tail = rlen - e - 1;
            const bool tail_ok = in_strand && match_len(row, e + 1, ix.text, te + 1, tail) == tail;
            if (!(clear && tail_ok)) {
                const uint32_t from = mirror ? npos - limit : i_next, cnt = limit - i_next;
                if (!unk.add(from, cnt)) { bail = true; return limit; }
            }
            return limit;
        }
    }
    return limit;
}

// PACKED: phase 1 loads the 2-bit rows written by k_scan_rows (scan_map.cu) instead of packing
// the raw bytes itself.
template <int STRIDE, int LPR, bool PACKED>
__device__ __forceinline__ void
map_fast_block(const IndexView& ix, const uint8_t* __restrict__ buf, const uint64_t* __restrict__ seq_start,
               const uint64_t* __restrict__ seq_end, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ hdr,
               uint32_t row_words, uint64_t n_reads, const bool listed, uint32_t spread,
               ReadSlot* __restrict__ slots, uint32_t* __restrict__ worklist, unsigned long long* __restrict__ counters,
               const uint32_t block_id) {
    // One block's worth of reads: MF_THREADS / spread of them, read block_id * that onwards.
    // listed: they are the reads the first tiers deferred, compact [0, n_reads).  Those are
    // few and each is a long serial chain (passes, range probes): only every spread-th thread
    // takes a read, which spreads them over `spread` times more warps.
    if ((uint64_t)block_id * (MF_THREADS / spread) >= n_reads) return;
    constexpr uint32_t CAP = (STRIDE - 3) * 16;                   // bases per packed row
    constexpr uint32_t GROUPS = 32 / LPR;                          // reads packed per warp step
    __shared__ uint32_t s_fwd[MF_THREADS * STRIDE];
    __shared__ uint32_t s_rc[MF_THREADS * STRIDE];
    __shared__ uint32_t s_node[MAXN][MF_THREADS];
    __shared__ uint32_t s_vk[MAXN][MF_THREADS];                    // v | kmin << 16
    __shared__ uint32_t s_len[MF_THREADS];                         // rlen | flags << 24
    constexpr uint32_t F_N = 1u << 24, F_BAD = 2u << 24, F_LONG = 4u << 24, F_NONE = 8u << 24;
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t r0 = (uint64_t)block_id * (MF_THREADS / spread);
    const uint32_t L = ix.split_len;

    if (PACKED) {
        // ---- phase 1 (packed rows): each thread loads its own row + builds the reverse complement
        const uint32_t t = threadIdx.x;
        const bool live = (t % spread) == 0 && r0 + t / spread < n_reads;
        uint32_t lfv = F_NONE;
        if (live) {
            const uint64_t r = listed ? r0 + t / spread : r0 + t;
            const uint32_t h = __ldg(hdr + r);
            if (h & PH_LONG) lfv = F_LONG;
            else {
                const uint32_t rlen = h & 0xFFFFFF;
                load_row<STRIDE, true>(rows, r, row_words, rlen, s_fwd + t * STRIDE, s_rc + t * STRIDE);
                lfv = rlen | ((h & PH_N) ? F_N : 0) | ((h & PH_BAD) ? F_BAD : 0);
            }
        }
        s_len[t] = lfv;
    } else
    // ---- phase 1: cooperative pack (LPR lanes per read) -------------------------------------
    {
        const uint32_t grp = lane / LPR, gl = lane % LPR;
        const uint32_t gmask = LPR == 32 ? 0xFFFFFFFFu : (((1u << LPR) - 1) << (grp * LPR));
        for (uint32_t k = 0; k < 32 / GROUPS; k++) {
            const uint32_t t = wib * 32 + k * GROUPS + grp;
            const bool live = r0 + t < n_reads;
            const uint64_t r = r0 + t;                       // (raw-byte mode is never list driven)
            uint32_t* row = s_fwd + t * STRIDE;
            uint32_t* rrow = s_rc + t * STRIDE;
            uint64_t s = 0, len64 = 0;
            uint32_t h = 0;
            if (PACKED) {
                if (live) h = __ldg(hdr + r);
                len64 = (h & PH_LONG) ? (uint64_t)CAP + 1 : (h & 0xFFFFFF);
            } else if (live) {
                s = seq_start[r];
                len64 = seq_end[r] - s;
            }
            const bool fits = live && len64 <= CAP;
            const uint32_t rlen = fits ? (uint32_t)len64 : 0;
            const uint32_t nwords = (rlen + 15) >> 4;
            uint32_t packed = 0;
            bool hasN = false, bad = false;
            if (PACKED) {
                if (gl < nwords) packed = __ldg(rows + r * row_words + gl);
                hasN = (h & PH_N) != 0;
                bad = (h & PH_BAD) != 0;
            } else if (gl < nwords) {
                // first byte of this lane's 16 bases, as an ABSOLUTE address (shards may be misaligned)
                const uintptr_t a = reinterpret_cast<uintptr_t>(buf) + s + 16ull * gl;
                const uint32_t nb = min(16u, rlen - 16 * gl);
                const uint32_t sh = (uint32_t)(a & 3) * 8;
                const uint32_t* p = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
                uint32_t x[5];
                const uint32_t need = ((uint32_t)(a & 3) + nb + 3) >> 2;
#pragma unroll
                for (int j = 0; j < 5; j++) x[j] = (uint32_t)j < need ? __ldg(p + j) : 0u;
                uint32_t diff = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t c = __funnelshift_r(x[j], x[j + 1], sh);
                    const int left = (int)nb - 4 * j;               // valid bytes in this word
                    if (left <= 0) break;
                    const uint32_t vm = left >= 4 ? 0xFFFFFFFFu : ((1u << (8 * left)) - 1);
                    const uint32_t c2 = ((c & vm) >> 1) & 0x03030303u;
                    // the only byte with code k is "ACTG"[k] = 0x41 + 2k (+15 when k == 2)
                    const uint32_t is2 = (c2 >> 1) & ~c2 & 0x01010101u;
                    const uint32_t expect = 0x41414141u + 2 * c2 + 15 * is2;
                    diff |= (expect ^ c) & vm;
                    packed |= ((c2 * 0x01041040u) >> 24) << (8 * j);
                }
                if (diff) {                                         // rare: some byte is not ACGT
                    for (uint32_t j = 0; j < nb; j++) {
                        const uint32_t c = (__funnelshift_r(x[j >> 2], x[(j >> 2) + 1], sh) >> (8 * (j & 3))) & 0xFF;
                        if (c == 'N') hasN = true;
                        else if (!is_acgt(c)) bad = true;
                    }
                }
            }
            // reverse complement: reverse the 2-bit groups of every word, complement, then
            // realign by the (16*nwords - rlen) padding bases that now sit in front
            uint32_t rv = __brev(packed);
            rv = (((rv & 0x55555555u) << 1) | ((rv >> 1) & 0x55555555u)) ^ 0xAAAAAAAAu;
            const int src0 = (int)nwords - 1 - (int)gl, src1 = src0 - 1;
            uint32_t a0 = __shfl_sync(0xFFFFFFFFu, rv, (src0 >= 0 ? src0 : 0) + grp * LPR);
            uint32_t a1 = __shfl_sync(0xFFFFFFFFu, rv, (src1 >= 0 ? src1 : 0) + grp * LPR);
            if (src0 < 0) a0 = 0;
            if (src1 < 0) a1 = 0;
            const uint32_t pad = 16 * nwords - rlen;
            const uint32_t rcw = __funnelshift_r(a0, a1, 2 * pad);
            const uint32_t bN = __ballot_sync(0xFFFFFFFFu, hasN) & gmask;
            const uint32_t bB = __ballot_sync(0xFFFFFFFFu, bad) & gmask;
            if (fits) {
                for (uint32_t w = gl; w < (uint32_t)STRIDE; w += LPR) {
                    row[w] = w < nwords ? packed : 0u;
                    rrow[w] = w < nwords ? rcw : 0u;
                }
            }
            if (gl == 0)
                s_len[t] = !live ? F_NONE : !fits ? F_LONG : (rlen | (bN ? F_N : 0) | (bB ? F_BAD : 0));
        }
    }
    __syncwarp();        // every warp packs, maps and confirms only its own 32 reads

    // ---- phase 2: one thread per read, passes A and B ---------------------------------------
    const uint32_t t = threadIdx.x;
    uint32_t lf = s_len[t];
    const uint64_t r = (lf & F_NONE) ? 0 : listed ? r0 + t / spread : r0 + t;
    const uint32_t rlen = lf & 0xFFFFFF;
    const uint32_t* row = s_fwd + t * STRIDE;
    uint32_t nn = 0;
    Unk unk;
    bool active = !(lf & (F_NONE | F_LONG | F_BAD | F_N)) && rlen >= L;
    bool bail = (lf & (F_LONG | F_BAD)) != 0 && !(lf & F_NONE);
    if ((lf & F_BAD) && ((lf & F_N) || rlen < L)) bail = false;    // N / short win over the bail
    uint32_t npos = 0;
    if (active) {
        npos = rlen - L + 1;
        const uint32_t uA = run_pass(ix, row, rlen, npos, false, npos, s_node, s_vk, t, nn, unk, bail);
        if (!bail && uA < npos) {
            const uint32_t lim = npos - uA;
            const uint32_t uB = run_pass(ix, s_rc + t * STRIDE, rlen, lim, true, npos, s_node, s_vk, t, nn, unk, bail);
            if (!bail && uB < lim && !unk.add(uA, lim - uB)) bail = true;
        }
        if (bail) { unk.r0 = 0; unk.r1 = 0; }
    }

    // ---- phase 3: the warp probes the unknown windows of its reads, one range at a time --------
    // Normally every one of them misses (they cover a sequencing error).  A window that does hit
    // a unique posting is one more hit for that node; a window with several postings sends the
    // read to the exhaustive tier.
#pragma unroll 1
    for (int slot = 0; slot < 2; slot++) {
        const uint32_t mine = slot == 0 ? unk.r0 : unk.r1;
        uint32_t m = __ballot_sync(0xFFFFFFFFu, (mine >> 16) != 0);
        while (m) {
            const int src = __ffs((int)m) - 1;
            m &= m - 1;
            const uint32_t rg = __shfl_sync(0xFFFFFFFFu, mine, src);
            const uint32_t from = rg & 0xFFFF, cnt = rg >> 16;
            uint32_t tnn = __shfl_sync(0xFFFFFFFFu, nn, src);          // list length of read src (uniform copy)
            bool tbail = __shfl_sync(0xFFFFFFFFu, (uint32_t)bail, src) != 0;
            const uint32_t tt = wib * 32 + src;
            for (uint32_t w0 = 0; w0 < cnt && !tbail; w0 += 32) {
                const uint32_t w = w0 + lane;
                uint32_t tp = 0, node = 0;
                int res = PROBE_MISS;
                if (w < cnt) res = probe_window(ix, s_fwd + tt * STRIDE, from + w, tp, node);
                if (__any_sync(0xFFFFFFFFu, res == PROBE_MULTI)) { tbail = true; break; }
                uint32_t hm = __ballot_sync(0xFFFFFFFFu, res == PROBE_UNIQUE);
                while (hm) {
                    const int leader = __ffs((int)hm) - 1;
                    const uint32_t lnode = __shfl_sync(0xFFFFFFFFu, node, leader);
                    const uint32_t grp = __ballot_sync(0xFFFFFFFFu, res == PROBE_UNIQUE && node == lnode);
                    // every lane runs the (uniform) list update on shared memory; one lane stores
                    uint32_t a = 0;
                    for (; a < tnn; a++) if (s_node[a][tt] == lnode) break;
                    const uint32_t hits = __popc(grp), kminc = from + w0 + (uint32_t)(__ffs((int)grp) - 1);
                    if (a == tnn) {
                        if (tnn == MAXN) { tbail = true; break; }
                        if (lane == 0) { s_node[a][tt] = lnode; s_vk[a][tt] = hits | (kminc << 16); }
                        tnn++;
                    } else if (lane == 0) {
                        const uint32_t old = s_vk[a][tt];
                        s_vk[a][tt] = ((old & 0xFFFF) + hits) | (min(old >> 16, kminc) << 16);
                    }
                    __syncwarp();
                    hm &= ~grp;
                }
            }
            if (lane == (uint32_t)src) { nn = tnn; if (tbail) bail = true; }
        }
        __syncwarp();
    }

    // ---- phase 4: finalize -----------------------------------------------------------------
    if (lf & F_NONE) return;
    if (t == 0 && block_id == 0 && !listed) atomicAdd(&counters[CNT_FAST], (unsigned long long)n_reads);
    ReadSlot* out = slots + r;
    if (!(lf & F_LONG)) {
        if (lf & F_N) { out->hdr = ST_N; return; }
        if (rlen < L) { out->hdr = ST_SHORT; return; }
    }
    uint32_t n_out = 0;
    if (!bail && nn <= (uint32_t)FL_MAX) {
        // the common case: up to FL_MAX distinct nodes (list_add merged repeats) -- sort network,
        // saturation predicate, ranks by popcount
        FlatList fl;
#pragma unroll
        for (int i = 0; i < FL_MAX; i++)
            fl.e[i] = (uint32_t)i < nn ? (((uint64_t)s_node[i][t] << 32) | s_vk[i][t]) : FL_EMPTY;
        fl.sort();
        n_out = flat_finalize(fl, ix, rlen, L, out);
    } else if (!bail) {
        // saturation predicate per node, then ascending node order (as enumerate(nodes) gives) from
        // ranks instead of a data-dependent sort
        uint32_t keepmask = 0;
        for (uint32_t a = 0; a < nn; a++) {
            const uint32_t vk = s_vk[a][t];
            if (keep_node_f(vk & 0xFFFF, vk >> 16, __ldg(ix.node_len + s_node[a][t]), rlen, L)) keepmask |= 1u << a;
        }
        n_out = __popc(keepmask);
        if (n_out > (uint32_t)SLOT_IDS) bail = true;           // 16 kept nodes do not fit the slot
        else {
            for (uint32_t a = 0; a < nn; a++) {
                if (!((keepmask >> a) & 1)) continue;
                const uint32_t node = s_node[a][t];
                uint32_t rank = 0;
                for (uint32_t b = 0; b < nn; b++) rank += ((keepmask >> b) & 1) && s_node[b][t] < node;
                out->ids[rank] = node;
            }
        }
    }
    if (bail) {
        const unsigned long long idx = atomicAdd(&counters[CNT_WORK], 1ull);
        worklist[idx] = (uint32_t)(r);
        return;
    }
    out->hdr = ST_OK | (n_out << 8);
}

// Direct mode (in_count == nullptr): one block per MF_THREADS reads.  List mode: a fixed grid walks the compact
// arrays of the deferred reads (their number is only known on the device), so no empty blocks are launched.
template <int STRIDE, int LPR, bool PACKED>
#ifdef VSPE_MF_MINB
__global__ void __launch_bounds__(MF_THREADS, VSPE_MF_MINB)
#else
__global__ void __launch_bounds__(MF_THREADS)
#endif
k_map_fast(IndexView ix, const uint8_t* __restrict__ buf, const uint64_t* __restrict__ seq_start,
           const uint64_t* __restrict__ seq_end, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ hdr,
           uint32_t row_words, uint64_t n_reads_arg,
           const unsigned long long* __restrict__ in_count, uint32_t list_spread, ReadSlot* __restrict__ slots,
           uint32_t* __restrict__ worklist, unsigned long long* __restrict__ counters) {
    if (!in_count) {
        map_fast_block<STRIDE, LPR, PACKED>(ix, buf, seq_start, seq_end, rows, hdr, row_words, n_reads_arg, false, 1u, slots, worklist,
                                            counters, blockIdx.x);
        return;
    }
    const uint64_t n_items = *in_count;
    const uint64_t per_block = MF_THREADS / list_spread;
    const uint64_t n_blocks = (n_items + per_block - 1) / per_block;
    for (uint64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        map_fast_block<STRIDE, LPR, PACKED>(ix, buf, seq_start, seq_end, rows, hdr, row_words, n_items, true, list_spread, slots,
                                            worklist, counters, (uint32_t)b);
        __syncthreads();                                   // the block's shared-memory rows are reused by the next round
    }
}

// ---------------------------------------------------------------------------------------------
// k_map_windows: the reads k_map_fast could not prove (repeats, palindromes, more than 16 nodes).
// One warp per read, one lane per window, every window looked up -- the reference's algorithm
// (PE_Inference.py:24-31) with no shortcuts, so every postings multiplicity is exact.  All lanes
// stay busy; hit counts and first positions live in lane registers (lane k owns the k-th distinct
// node of the read, up to 32), so there is no per-read scratch in memory.
// ---------------------------------------------------------------------------------------------
static constexpr int MW_WARPS = 4;

template <int STRIDE>
__global__ void __launch_bounds__(MW_WARPS * 32)
k_map_windows(IndexView ix, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ hdr, uint32_t row_words,
              const uint32_t* __restrict__ in_list, const unsigned long long* __restrict__ in_count,
              ReadSlot* __restrict__ slots, uint32_t* __restrict__ out_list, unsigned long long* __restrict__ out_count,
              uint32_t* __restrict__ spill, uint64_t spill_cap, unsigned long long* __restrict__ counters) {
    __shared__ uint32_t s_row[MW_WARPS][STRIDE];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t n_items = *in_count;
    const uint64_t gw = (uint64_t)blockIdx.x * MW_WARPS + wib, nw = (uint64_t)gridDim.x * MW_WARPS;
    const uint32_t L = ix.split_len;
    uint32_t* row = s_row[wib];
    for (uint64_t item = gw; item < n_items; item += nw) {
        const uint32_t r = in_list[item];
        const uint32_t h = __ldg(hdr + r);
        if (h & (PH_LONG | PH_BAD)) {                       // not representable in 2 bits: exhaustive ASCII tier
            if (lane == 0) out_list[atomicAdd(out_count, 1ull)] = r;
            continue;
        }
        const uint32_t rlen = h & 0xFFFFFF, npos = rlen - L + 1, nwords = (rlen + 15) >> 4;
        __syncwarp();
        for (uint32_t w = lane; w < (uint32_t)STRIDE; w += 32) row[w] = w < nwords ? __ldg(rows + (uint64_t)r * row_words + w) : 0u;
        __syncwarp();
        // lane k: k-th distinct node of this read
        uint32_t acc_node = NONE32, acc_v = 0, acc_kmin = NONE32, nn = 0;
        bool overflow = false;
        for (uint32_t w0 = 0; w0 < npos; w0 += 32) {
            const uint32_t w = w0 + lane;
            bool done = w >= npos;
            uint64_t hsh = 0;
            uint32_t j = 0;
            if (!done) {
                hsh = hash_read(row, w, L);
                j = slot_of(hsh, ix.slot_mask);
                if (ix.bloom != nullptr && !bloom_maybe(ix.bloom, ix.bloom_mask, hsh)) done = true;   // proven miss
            }
            while (true) {
                // advance every lane to its next verified posting (or to the end of its cluster)
                uint32_t node = NONE32;
                while (!done) {
                    const uint2 ent = __ldg(ix.slots + j);
                    j = (j + 1) & ix.slot_mask;
                    if (ent.x == EMPTY_TP) { done = true; break; }
                    if (fp_match(ent.y, hsh, ix.node_mask) && read_equals_text(row, w, ix.text, ent.x, L)) {
                        node = ent.y & ix.node_mask;
                        break;
                    }
                }
                uint32_t pending = __ballot_sync(0xFFFFFFFFu, node != NONE32);
                if (!pending) break;
                while (pending) {
                    const int leader = __ffs((int)pending) - 1;
                    const uint32_t lnode = __shfl_sync(0xFFFFFFFFu, node, leader);
                    const uint32_t grp = __ballot_sync(0xFFFFFFFFu, node == lnode);
                    const uint32_t cnt = __popc(grp), first = w0 + (uint32_t)(__ffs((int)grp) - 1);
                    const uint32_t owner = __ballot_sync(0xFFFFFFFFu, acc_node == lnode);
                    if (owner) {
                        if (acc_node == lnode) { acc_v += cnt; acc_kmin = min(acc_kmin, first); }
                    } else if (nn < 32) {
                        if (lane == nn) { acc_node = lnode; acc_v = cnt; acc_kmin = first; }
                        nn++;
                    } else {
                        overflow = true;
                    }
                    pending &= ~grp;
                }
            }
        }
        if (overflow) {                                       // more than 32 distinct nodes: exhaustive tier
            if (lane == 0) out_list[atomicAdd(out_count, 1ull)] = r;
            continue;
        }
        // saturation predicate per node, then ascending node order
        const bool keep = lane < nn && keep_node_f(acc_v, acc_kmin, __ldg(ix.node_len + acc_node), rlen, L);
        const uint32_t key = keep ? acc_node : NONE32;
        uint32_t rank = 0;
#pragma unroll 8
        for (int k = 0; k < 32; k++) rank += __shfl_sync(0xFFFFFFFFu, key, k) < key;
        const uint32_t n_out = __popc(__ballot_sync(0xFFFFFFFFu, keep));
        ReadSlot* out = slots + r;
        if (n_out <= (uint32_t)SLOT_IDS) {
            if (keep) out->ids[rank] = acc_node;
            if (lane == 0) out->hdr = ST_OK | (n_out << 8);
        } else {
            unsigned long long off = 0;
            if (lane == 0) off = atomicAdd(&counters[CNT_SPILL_CURSOR], (unsigned long long)n_out);
            off = __shfl_sync(0xFFFFFFFFu, off, 0);
            if (off + n_out > spill_cap) {
                if (lane == 0) { atomicOr(&counters[CNT_ERR], (unsigned long long)ERRF_SPILL_FULL); out->hdr = ST_OK; }
            } else {
                if (keep) spill[off + rank] = acc_node;
                if (lane == 0) { out->ids[0] = (uint32_t)off; out->hdr = ST_OK | (n_out << 8); }
            }
        }
    }
}

int map_reads_generic_dev(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                          const uint32_t* d_worklist, const unsigned long long* d_n_items, ReadSlot* d_slots);

// Work lists of the list-driven tiers for a chunk of up to n_reads reads (reserved and emptied before
// the first kernel that appends to them).
int map_prepare_lists(Ctx* c, uint64_t n_reads) {
    if (n_reads > 0xFFFFFFFFull) { set_error("more than 2^32 reads in one chunk"); return VSPE_ERR_LIMIT; }
    VSPE_TRY(c->worklist.reserve(n_reads + 1));
    VSPE_TRY(c->defer_list.reserve(n_reads + 2));
    VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_WORK, 0, 8, c->stream));
    VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_WORK2, 0, 8, c->stream));
    return VSPE_OK;
}

static constexpr uint32_t LIST_SPREAD = 2;      // list-driven k_map_fast: every 2nd thread takes a read (long serial chains)

// The tiers behind the walk.  pre_count != nullptr: the reads to map are the *pre_count deferred reads k_walk copied
// to compact arrays (rows, headers, byte ranges; slots are compact too); otherwise every read [0, n_reads) goes
// through the full seed-and-extend kernel, which packs the raw bytes itself.
//   stage 1  k_map_fast     seed-and-extend in both directions + cooperative confirmation probes
//   stage 2  k_map_windows  what stage 1 could not prove (repeats, > 16 nodes): one warp per read, every window looked up
//                           (packed reads only)
//   stage 3  k_map_generic  the exhaustive ASCII tier: any read length / alphabet
static int launch_map_fast(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                           const uint32_t* d_rows, const uint32_t* d_hdr, uint32_t row_words, uint32_t cap,
                           uint64_t n_reads, ReadSlot* d_slots, const unsigned long long* pre_count) {
    const bool pre_listed = pre_count != nullptr;
    if (n_reads == 0) return VSPE_OK;
    if (n_reads > 0xFFFFFFFFull) { set_error("more than 2^32 reads in one chunk"); return VSPE_ERR_LIMIT; }
    if (!pre_listed) {
        VSPE_TRY(c->worklist.reserve(n_reads));
        VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_WORK, 0, 8, c->stream));
    }
    IndexView v = c->index.view();
    const unsigned long long* in_count = pre_count;
    const uint32_t* to_generic = c->worklist.p;                 // reads the ASCII tier must map
    const unsigned long long* to_generic_n = c->counters.p + CNT_WORK;
    // list mode: one resident wave (the blocks loop over the device-side list; more blocks than fit would run as a
    // second, partly filled wave)
    if (pre_listed && c->mf_blocks_per_sm == 0) {
        int nb = 0;
        cudaError_t e = cap <= 160 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_map_fast<13, 16, true>, MF_THREADS, 0)
                      : cap <= 256 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_map_fast<19, 16, true>, MF_THREADS, 0)
                                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_map_fast<23, 32, true>, MF_THREADS, 0);
        c->mf_blocks_per_sm = (e == cudaSuccess && nb > 0) ? nb : 4;
        c->mf_blocks_cap = cap;
    }
    if (pre_listed && c->mf_blocks_cap != cap) { c->mf_blocks_per_sm = 0; return launch_map_fast(c, d_buf, d_seq_start, d_seq_end, d_rows, d_hdr, row_words, cap, n_reads, d_slots, pre_count); }
    const uint32_t grid = pre_listed ? (uint32_t)std::min<uint64_t>((n_reads * LIST_SPREAD + MF_THREADS - 1) / MF_THREADS, (uint64_t)c->sm_count * c->mf_blocks_per_sm)
                                     : (uint32_t)((n_reads + MF_THREADS - 1) / MF_THREADS);
#define VSPE_MF(S, LP, PK) k_map_fast<S, LP, PK><<<grid, MF_THREADS, 0, c->stream>>>(v, d_buf, d_seq_start, d_seq_end, d_rows, d_hdr, \
                                                                              row_words, n_reads, in_count, LIST_SPREAD, d_slots, \
                                                                              c->worklist.p, c->counters.p)
    if (pre_listed) { if (cap <= 160) VSPE_MF(13, 16, true); else if (cap <= 256) VSPE_MF(19, 16, true); else VSPE_MF(23, 32, true); }
    else { if (cap <= 160) VSPE_MF(13, 16, false); else if (cap <= 256) VSPE_MF(19, 16, false); else VSPE_MF(23, 32, false); }
#undef VSPE_MF
    VSPE_LAUNCH_CHECK(c);
    if (pre_listed) {
        uint32_t* list2 = c->defer_list.p;
        const uint32_t wgrid = (uint32_t)std::min<uint64_t>((n_reads + MW_WARPS - 1) / MW_WARPS, (uint64_t)c->sm_count * 16);
#define VSPE_MW(S) k_map_windows<S><<<wgrid, MW_WARPS * 32, 0, c->stream>>>(v, d_rows, d_hdr, row_words, c->worklist.p, c->counters.p + CNT_WORK, \
                                                                      d_slots, list2, c->counters.p + CNT_WORK2, c->spill.p, c->spill.cap, c->counters.p)
        if (cap <= 160) VSPE_MW(13); else if (cap <= 256) VSPE_MW(19); else VSPE_MW(23);
#undef VSPE_MW
        VSPE_LAUNCH_CHECK(c);
        to_generic = list2;
        to_generic_n = c->counters.p + CNT_WORK2;
    }
    // last stage: the exhaustive ASCII tier consumes what is left; list lengths stay on the device
    VSPE_TRY(map_reads_generic_dev(c, d_buf, d_seq_start, d_seq_end, to_generic, to_generic_n, d_slots));
    return VSPE_OK;
}

// packed-row capacity (bases) by the read length seen in the first records; longer reads bail
uint32_t map_fast_cap(uint32_t hint) { return hint <= 160 ? 160 : hint <= 256 ? 256 : 320; }

int map_reads_fast(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                   uint64_t n_reads, ReadSlot* d_slots) {
    if (c->index.split_len > 320) return map_reads_generic(c, d_buf, d_seq_start, d_seq_end, n_reads, d_slots);
    return launch_map_fast(c, d_buf, d_seq_start, d_seq_end, nullptr, nullptr, 0, map_fast_cap(c->read_len_hint), n_reads, d_slots, nullptr);
}

// the *d_count reads k_walk left unresolved (compact arrays): full seed-and-extend kernel, then the
// all-windows kernel, then the ASCII tier; results go to d_slots[i], i < *d_count
int map_reads_deferred(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                       const uint32_t* d_rows, const uint32_t* d_hdr, uint32_t row_words, uint32_t cap,
                       uint64_t n_reads_cap, ReadSlot* d_slots, const unsigned long long* d_count) {
    return launch_map_fast(c, d_buf, d_seq_start, d_seq_end, d_rows, d_hdr, row_words, cap, n_reads_cap, d_slots, d_count);
}

}  // namespace vspe
