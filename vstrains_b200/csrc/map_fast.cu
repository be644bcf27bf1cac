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
static constexpr uint32_t BIG_SAMPLE_BLOCKS = 64;          // k_map_first blocks that report reads with > FL_MAX stretches
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
    // sorted list: fold every run of equal nodes into its last entry (hits add up, smallest position wins)
    __device__ __forceinline__ void merge_repeats() {
#pragma unroll
        for (int i = 0; i + 1 < FL_MAX; i++) {
            if (e[i + 1] != FL_EMPTY && (uint32_t)(e[i] >> 32) == (uint32_t)(e[i + 1] >> 32)) {
                const uint32_t x = (uint32_t)e[i], y = (uint32_t)e[i + 1];
                const uint32_t vk = ((x & 0xFFFF) + (y & 0xFFFF)) | (min(x >> 16, y >> 16) << 16);
                e[i + 1] = (e[i + 1] & 0xFFFFFFFF00000000ull) | vk;
                e[i] = FL_EMPTY;
            }
        }
    }
    // copy the register entries to the thread-local arrays of the general path
    __device__ __forceinline__ void spill_to(uint32_t* l_node, uint32_t* l_vk) const {
#pragma unroll
        for (int i = 0; i < FL_MAX; i++) { l_node[i] = (uint32_t)(e[i] >> 32); l_vk[i] = (uint32_t)e[i]; }
    }
    // two stretches of the same node (cyclic graph)?
    __device__ __forceinline__ bool has_repeat() const {
        bool r = false;
#pragma unroll
        for (int i = 0; i + 1 < FL_MAX; i++) r |= e[i + 1] != FL_EMPTY && (uint32_t)(e[i] >> 32) == (uint32_t)(e[i + 1] >> 32);
        return r;
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

// The general case (more than FL_MAX stretches: graphs with many short nodes per read): the list
// lives in thread-local arrays; repeats are merged (or reported), then the same predicate and an
// O(n^2) rank.  Returns false if the read must go to the next tier.
__device__ __noinline__ bool list_finalize_slow(uint32_t* l_node, uint32_t* l_vk, uint32_t nn, bool merge, const IndexView& ix,
                                                uint32_t rlen, uint32_t L, ReadSlot* out, uint32_t& n_out) {
    for (uint32_t a = 1; a < nn; a++) {
        for (uint32_t b = 0; b < a; b++) {
            if (l_node[b] == l_node[a] && l_vk[b]) {
                if (!merge) return false;
                const uint32_t x = l_vk[b], y = l_vk[a];
                l_vk[b] = ((x & 0xFFFF) + (y & 0xFFFF)) | (min(x >> 16, y >> 16) << 16);
                l_vk[a] = 0;
                break;
            }
        }
    }
    uint32_t keepmask = 0;
    for (uint32_t a = 0; a < nn; a++) {
        const uint32_t vk = l_vk[a];
        if (vk && keep_node_f(vk & 0xFFFF, vk >> 16, __ldg(ix.node_len + l_node[a]), rlen, L)) keepmask |= 1u << a;
    }
    n_out = __popc(keepmask);
    if (n_out > (uint32_t)SLOT_IDS) return false;
    for (uint32_t a = 0; a < nn; a++) {
        if (!((keepmask >> a) & 1)) continue;
        uint32_t rank = 0;
        for (uint32_t b = 0; b < nn; b++) rank += ((keepmask >> b) & 1) && l_node[b] < l_node[a];
        out->ids[rank] = l_node[a];
    }
    return true;
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

// PACKED: phase 1 loads the 2-bit rows written by k_scan_pack (scan_pack.cu) instead of packing
// the raw bytes itself.
template <int STRIDE, int LPR, bool PACKED>
__device__ __forceinline__ void
map_fast_block(const IndexView& ix, const uint8_t* __restrict__ buf, const uint64_t* __restrict__ seq_start,
               const uint64_t* __restrict__ seq_end, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ hdr,
               uint32_t row_words, uint64_t n_reads, const uint32_t* __restrict__ in_list, uint32_t spread,
               ReadSlot* __restrict__ slots, uint32_t* __restrict__ worklist, unsigned long long* __restrict__ counters,
               const uint32_t block_id) {
    // One block's worth of reads: MF_THREADS / spread of them, read block_id * that onwards.
    // in_list != nullptr: they are in_list[...] (the reads the first tiers deferred).  Those are
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
            const uint64_t r = in_list ? (uint64_t)in_list[r0 + t / spread] : r0 + t;
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
    const uint64_t r = (lf & F_NONE) ? 0 : in_list ? (uint64_t)in_list[r0 + t / spread] : r0 + t;
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
    if (t == 0 && block_id == 0 && !in_list) atomicAdd(&counters[CNT_FAST], (unsigned long long)n_reads);
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

// Direct mode (in_list == nullptr): one block per MF_THREADS reads.  List mode: a fixed grid walks
// the device-side worklist (its length is only known on the device), so no empty blocks are launched.
template <int STRIDE, int LPR, bool PACKED>
__global__ void __launch_bounds__(MF_THREADS)
k_map_fast(IndexView ix, const uint8_t* __restrict__ buf, const uint64_t* __restrict__ seq_start,
           const uint64_t* __restrict__ seq_end, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ hdr,
           uint32_t row_words, uint64_t n_reads_arg, const uint32_t* __restrict__ in_list,
           const unsigned long long* __restrict__ in_count, uint32_t list_spread, ReadSlot* __restrict__ slots,
           uint32_t* __restrict__ worklist, unsigned long long* __restrict__ counters) {
    if (!in_list) {
        map_fast_block<STRIDE, LPR, PACKED>(ix, buf, seq_start, seq_end, rows, hdr, row_words, n_reads_arg, nullptr, 1u, slots, worklist,
                                            counters, blockIdx.x);
        return;
    }
    const uint64_t n_items = *in_count;
    const uint64_t per_block = MF_THREADS / list_spread;
    const uint64_t n_blocks = (n_items + per_block - 1) / per_block;
    for (uint64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        map_fast_block<STRIDE, LPR, PACKED>(ix, buf, seq_start, seq_end, rows, hdr, row_words, n_items, in_list, list_spread, slots,
                                            worklist, counters, (uint32_t)b);
        __syncthreads();                                   // the block's shared-memory rows are reused by the next round
    }
}

// ---------------------------------------------------------------------------------------------
// k_map_first: the common case only.  Packs nothing (rows come from k_scan_pack), runs ONE clean
// forward pass -- seed window 0, extend, walk successors -- and finishes the read if that pass
// proves every window.  Anything else (a miss, a mismatch, a repeat, too many nodes) defers the
// read, untouched, to k_map_fast via a worklist, so the two populations never share a warp.
// ---------------------------------------------------------------------------------------------
// GENERAL: reads with more than FL_MAX stretches keep the extra entries in thread-local arrays and
// finish in the O(n^2) path; otherwise they are deferred.  The host picks the variant from the share
// of such reads in the previous launch (CNT_BIG, sampled in the first blocks): graphs with long nodes
// run the lean variant, graphs with many short nodes per read the general one.
template <int STRIDE, int LPR, bool FLAT, bool GENERAL>
__global__ void __launch_bounds__(MF_THREADS)
k_map_first(IndexView ix, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ hdr, uint32_t row_words,
            uint64_t n_reads, ReadSlot* __restrict__ slots, uint32_t* __restrict__ defer_list,
            unsigned long long* __restrict__ counters) {
    __shared__ uint32_t s_fwd[MF_THREADS * STRIDE];
    __shared__ uint32_t s_len[MF_THREADS];
    const uint64_t r0 = (uint64_t)blockIdx.x * MF_THREADS;
    const uint32_t L = ix.split_len;
    {
        const uint32_t t = threadIdx.x;
        const uint64_t r = r0 + t;
        uint32_t h = 0xFFFFFFFFu;
        if (r < n_reads) {
            h = __ldg(hdr + r);
            if (!(h & PH_LONG)) load_row<STRIDE, false>(rows, r, row_words, h & 0xFFFFFF, s_fwd + t * STRIDE, nullptr);
        }
        s_len[t] = h;
    }
    __syncwarp();
    const uint32_t t = threadIdx.x;
    const uint64_t r = r0 + t;
    const uint32_t h = s_len[t];
    if (h == 0xFFFFFFFFu) return;
    if (t == 0 && blockIdx.x == 0) atomicAdd(&counters[CNT_FAST], (unsigned long long)n_reads);
    ReadSlot* out = slots + r;
    const uint32_t rlen = h & 0xFFFFFF;
    bool defer = (h & (PH_LONG | PH_BAD)) != 0;
    if (!(h & PH_LONG)) {
        if (h & PH_N) { out->hdr = ST_N; return; }
        if (rlen < L) { out->hdr = ST_SHORT; return; }
    }
    const uint32_t* row = s_fwd + t * STRIDE;
    uint32_t nn = 0;
    FlatList fl;                                           // per-read node list: first FL_MAX entries in registers,
    uint32_t l_node[GENERAL ? MAXN : 1], l_vk[GENERAL ? MAXN : 1];   // the rest (GENERAL) thread-local
    constexpr uint32_t LIST_CAP = GENERAL ? (uint32_t)MAXN : (uint32_t)FL_MAX;
    bool big = false;                                      // more than FL_MAX stretches
    fl.clear();
    if (!defer) {
        uint32_t tp = NONE32, node = 0;
        if (probe_window(ix, row, 0, tp, node) != PROBE_UNIQUE) defer = true;
        // Walk the read along one diagonal per node strand: read base p sits at text position
        // p + delta.  Each step compares 32 bases and tests the uniq bits of the 32 windows that
        // END at those bases; the first window of a stretch is unique by construction (seed:
        // PROBE_UNIQUE, later ones: successor table).
        // FLAT: one loop whose every turn is "compare up to 32 bases, then -- if the stretch is complete --
        // append it and step to the successor strand", so the threads of a warp stay in step however
        // their reads are cut into stretches (the nested form runs max-stretches x max-chunks turns).
        uint32_t i0 = 0, p = L, q = 0, lim = 0;
        int delta = 0;
        auto enter = [&]() {                                   // strand of window i0 = text position tp
            const uint32_t s1 = __ldg(ix.strand_start + 2 * node + 1);
            const bool rcs = tp >= s1;
            q = 2 * node + (rcs ? 1u : 0u);
            const uint32_t send = rcs ? __ldg(ix.strand_start + 2 * node + 2) : s1;
            delta = (int)tp - (int)i0;
            lim = min(rlen, (uint32_t)((int)send - delta));    // read position where the strand ends
        };
        auto chunk = [&]() -> bool {                           // bases [p, p + n) and the windows ending there
            const uint32_t n = min(32u, lim - p);
            uint64_t x = read64(row, p) ^ extract64(ix.text, (uint64_t)((int)p + delta));
            const uint32_t u = (uint32_t)((int)p + delta) - L + 1;             // text position of the first window ending here
            const uint32_t ub = __funnelshift_r(__ldg(ix.uniq + (u >> 5)), __ldg(ix.uniq + (u >> 5) + 1), u & 31);
            const uint32_t m32 = n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1);
            if (n < 32) x &= (1ull << (2 * n)) - 1;
            if (x != 0 || (ub & m32) != m32) return false;
            p += n;
            return true;
        };
        if (FLAT) {
            bool running = !defer;
            if (running) enter();
            while (running) {
                if (p < lim && !chunk()) { defer = true; running = false; }
                if (running && p >= lim) {
                    // append the stretch; a node met twice (cyclic graph) is detected at the end and deferred
                    big |= nn == (uint32_t)FL_MAX;
                    if (nn == LIST_CAP) { defer = true; running = false; }
                    else {
                        const uint32_t vk = (lim - L + 1 - i0) | (i0 << 16);
                        if (nn < (uint32_t)FL_MAX) fl.set(nn, node, vk);
                        else if (GENERAL) { l_node[nn] = node; l_vk[nn] = vk; }
                        nn++;
                        if (lim >= rlen) running = false;
                        else {
                            // the strand ended before the read: successor window for the read's next base
                            const uint32_t b = (row[lim >> 4] >> ((lim & 15) * 2)) & 3u;
                            const uint2 sc = __ldg(reinterpret_cast<const uint2*>(ix.succ) + 4 * q + b);
                            if (sc.x == NONE32) { defer = true; running = false; }
                            else {
                                i0 = lim - L + 1;
                                tp = sc.x;
                                node = sc.y;
                                p = lim + 1;
                                enter();
                            }
                        }
                    }
                }
            }
        } else {
            while (!defer) {
                enter();
                while (p < lim) {
                    if (!chunk()) { defer = true; break; }
                }
                if (defer) break;
                big |= nn == (uint32_t)FL_MAX;
                if (nn == LIST_CAP) { defer = true; break; }
                if (nn < (uint32_t)FL_MAX) fl.set(nn, node, (lim - L + 1 - i0) | (i0 << 16));
                else if (GENERAL) { l_node[nn] = node; l_vk[nn] = (lim - L + 1 - i0) | (i0 << 16); }
                nn++;
                if (lim >= rlen) break;
                const uint32_t b = (row[lim >> 4] >> ((lim & 15) * 2)) & 3u;
                const uint2 sc = __ldg(reinterpret_cast<const uint2*>(ix.succ) + 4 * q + b);
                if (sc.x == NONE32) { defer = true; break; }
                i0 = lim - L + 1;
                tp = sc.x;
                node = sc.y;
                p = lim + 1;
            }
        }
    }
    uint32_t n_out = 0;
    if (!defer && nn <= (uint32_t)FL_MAX) {
        fl.sort();
        if (fl.has_repeat()) defer = true;                 // needs merging: full kernel
        else n_out = flat_finalize(fl, ix, rlen, L, out);
    } else if (!defer && GENERAL) {
        fl.spill_to(l_node, l_vk);
        if (!list_finalize_slow(l_node, l_vk, nn, false, ix, rlen, L, out, n_out)) defer = true;
    }
    if (big && blockIdx.x < BIG_SAMPLE_BLOCKS) atomicAdd(&counters[CNT_BIG], 1ull);
    if (defer) {
        const unsigned long long idx = atomicAdd(&counters[CNT_DEFER], 1ull);
        defer_list[idx] = (uint32_t)r;
        return;
    }
    out->hdr = ST_OK | (n_out << 8);
}

// ---------------------------------------------------------------------------------------------
// k_map_second: the reads k_map_first deferred, one thread per read, still one walk -- but ONE
// mismatching base is tolerated.  The windows covering it equal text windows except for that base,
// so a clear substitution-hit bit in every strand that holds such windows proves that they all miss
// (exactly what the reference's look-ups would find); they are simply not counted.  If window 0
// does not seed (error in the first split_len bases) the same walk runs on the reverse complement
// from the other end.  A second mismatch, a set bit, a repeat or a missing successor sends the read
// on to k_map_fast.
// ---------------------------------------------------------------------------------------------
template <int STRIDE, bool GENERAL>
__device__ __forceinline__ void
map_second_read(const IndexView& ix, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ hdr, uint32_t row_words,
                const uint32_t r, uint32_t* row, ReadSlot* __restrict__ slots, uint32_t* __restrict__ out_list,
                unsigned long long* __restrict__ out_count) {
    const uint32_t L = ix.split_len;
    const uint32_t h = __ldg(hdr + r);
    bool defer = (h & (PH_LONG | PH_BAD)) != 0 || ix.subst == nullptr;
    const uint32_t rlen = h & 0xFFFFFF;
    uint32_t nn = 0;
    FlatList fl;                                           // per-read node list: first FL_MAX entries in registers,
    uint32_t l_node[GENERAL ? MAXN : 1], l_vk[GENERAL ? MAXN : 1];   // the rest (GENERAL) thread-local
    constexpr uint32_t LIST_CAP = GENERAL ? (uint32_t)MAXN : (uint32_t)FL_MAX;
    if (!defer) {
        load_row<STRIDE, false>(rows, r, row_words, rlen, row, nullptr);
        const int npos = (int)(rlen - L + 1);
        // Which end seeds?  Window 0 of the read, else (error in the first split_len bases) window 0 of
        // its reverse complement.  The walk itself then runs ONCE, for all threads of the warp together.
        bool resolved = false, mirror = false;
        uint32_t tp = NONE32, node = 0;
        int pr = probe_window(ix, row, 0, tp, node);
        if (pr == PROBE_MISS) {
            // reverse-complement the packed row in place
            constexpr int NW = STRIDE - 3;
            const uint32_t nwords = (rlen + 15) >> 4, pad = 16 * nwords - rlen;
            uint32_t y[NW];
            uint32_t prev = 0;
#pragma unroll
            for (int k = 0; k < NW; k++) {
                y[k] = 0;
                const int kk = NW - 1 - k;
                if ((uint32_t)kk >= nwords) continue;
                uint32_t rv = __brev(row[kk]);
                rv = (((rv & 0x55555555u) << 1) | ((rv >> 1) & 0x55555555u)) ^ 0xAAAAAAAAu;
                const int j = (int)nwords - 1 - kk;
                if (j > 0) y[j - 1] = __funnelshift_r(prev, rv, 2 * pad);
                prev = rv;
            }
            if (nwords) y[nwords - 1] = __funnelshift_r(prev, 0u, 2 * pad);
#pragma unroll
            for (int k = 0; k < NW; k++) row[k] = y[k];
            mirror = true;
            pr = probe_window(ix, row, 0, tp, node);
        }
        fl.clear();
        // (PROBE_MULTI, or both ends miss: a real complication, the full kernel decides)
        bool running = pr == PROBE_UNIQUE;
        bool err = false;
        int e = 0;
        uint32_t rb = 0;
        uint32_t i0 = 0, p = L, q = 0, lim = 0;
        int delta = 0;
        // strand of window i0 = text position tp; false if the error's windows cannot be proven to miss there
        auto enter = [&]() -> bool {
            const uint32_t s1 = __ldg(ix.strand_start + 2 * node + 1);
            const bool rcs = tp >= s1;
            q = 2 * node + (rcs ? 1u : 0u);
            const uint32_t send = rcs ? __ldg(ix.strand_start + 2 * node + 2) : s1;
            delta = (int)tp - (int)i0;
            lim = min(rlen, (uint32_t)((int)send - delta));      // read position where the strand ends
            // a strand entered after the error still holds windows covering it if it starts at or before e
            if (err && (int)i0 <= e) {
                const uint32_t te = (uint32_t)(e + delta);
                if ((__ldg(ix.subst + (te >> 3)) >> (4 * (te & 7) + rb)) & 1u) return false;
            }
            return true;
        };
        if (running && !enter()) running = false;
        // flat loop: every turn compares up to 32 bases and, when the stretch is complete, books it and
        // steps to the successor strand
        while (running) {
            if (p < lim) {
                const uint32_t n = min(32u, lim - p);
                uint64_t x = read64(row, p) ^ extract64(ix.text, (uint64_t)((int)p + delta));
                if (n < 32) x &= (1ull << (2 * n)) - 1;
                bool ok = true;
                if (x) {
                    const uint32_t off = (uint32_t)(__ffsll((long long)x) - 1) >> 1;
                    if (err || (x & ~(3ull << (2 * off)))) ok = false;          // second mismatch
                    else {
                        err = true;
                        e = (int)(p + off);
                        rb = (row[(uint32_t)e >> 4] >> (((uint32_t)e & 15) * 2)) & 3u;
                        const uint32_t te = (uint32_t)(e + delta);
                        if ((__ldg(ix.subst + (te >> 3)) >> (4 * (te & 7) + rb)) & 1u) ok = false;
                    }
                }
                const uint32_t u = (uint32_t)((int)p + delta) - L + 1;     // text position of the first window ending here
                const uint32_t ub = __funnelshift_r(__ldg(ix.uniq + (u >> 5)), __ldg(ix.uniq + (u >> 5) + 1), u & 31);
                const uint32_t m32 = n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1);
                if ((ub & m32) != m32) ok = false;
                if (!ok) { running = false; break; }
                p += n;
            }
            if (p >= lim) {
                // windows [a, bw] of this node are resolved; those covering e are proven misses
                const int a = (int)i0, bw = (int)lim - (int)L;
                int c1 = bw - a + 1, c2 = 0, last_hit = bw, first_hit = a;
                if (err) {
                    c1 = min(bw, e - (int)L) - a + 1;
                    if (c1 < 0) c1 = 0;
                    const int a2 = max(a, e + 1);
                    c2 = bw - a2 + 1;
                    if (c2 < 0) c2 = 0;
                    first_hit = c1 > 0 ? a : a2;
                    last_hit = c2 > 0 ? bw : min(bw, e - (int)L);
                }
                if (c1 + c2 > 0) {
                    if (nn == LIST_CAP) { running = false; break; }
                    const uint32_t vk = (uint32_t)(c1 + c2) | ((uint32_t)(mirror ? npos - 1 - last_hit : first_hit) << 16);
                    if (nn < (uint32_t)FL_MAX) fl.set(nn, node, vk);
                    else if (GENERAL) { l_node[nn] = node; l_vk[nn] = vk; }
                    nn++;
                }
                if (lim >= rlen) { resolved = true; running = false; break; }
                // the strand ended before the read: successor window for the read's next base
                const uint32_t b = (row[lim >> 4] >> ((lim & 15) * 2)) & 3u;
                const uint2 sc = __ldg(reinterpret_cast<const uint2*>(ix.succ) + 4 * q + b);
                if (sc.x == NONE32) { running = false; break; }
                i0 = lim - L + 1;
                tp = sc.x;
                node = sc.y;
                p = lim + 1;
                if (!enter()) { running = false; break; }
            }
        }
        if (!resolved) defer = true;
    }
    ReadSlot* out = slots + r;
    uint32_t n_out = 0;
    if (!defer && nn <= (uint32_t)FL_MAX) {
        // a node met in two stretches (cyclic graph): after the sort they are neighbours; the last
        // of a run takes the sum of the hits and the smallest position
        fl.sort();
        fl.merge_repeats();
        n_out = flat_finalize(fl, ix, rlen, L, out);
    } else if (!defer && GENERAL) {
        fl.spill_to(l_node, l_vk);
        if (!list_finalize_slow(l_node, l_vk, nn, true, ix, rlen, L, out, n_out)) defer = true;
    }
    if (defer) {
        out_list[atomicAdd(out_count, 1ull)] = r;
        return;
    }
    out->hdr = ST_OK | (n_out << 8);
}

// A fixed grid walks the device-side worklist (its length is only known on the device).  The
// deferred reads are few and each is a long serial chain: only every `spread`-th thread takes one.
template <int STRIDE, bool GENERAL>
__global__ void __launch_bounds__(MF_THREADS)
k_map_second(IndexView ix, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ hdr, uint32_t row_words,
             const uint32_t* __restrict__ in_list, const unsigned long long* __restrict__ in_count, uint32_t spread,
             ReadSlot* __restrict__ slots, uint32_t* __restrict__ out_list, unsigned long long* __restrict__ out_count) {
    __shared__ uint32_t s_fwd[MF_THREADS * STRIDE];
    const uint64_t n_items = *in_count;
    const uint32_t per_block = MF_THREADS / spread;
    if (threadIdx.x % spread != 0) return;
    for (uint64_t base = (uint64_t)blockIdx.x * per_block; base < n_items; base += (uint64_t)gridDim.x * per_block) {
        const uint64_t item = base + threadIdx.x / spread;
        if (item < n_items)
            map_second_read<STRIDE, GENERAL>(ix, rows, hdr, row_words, in_list[item], s_fwd + threadIdx.x * STRIDE, slots, out_list, out_count);
    }
}

// ---------------------------------------------------------------------------------------------
// k_map_windows: the reads k_map_first deferred (sequencing errors, repeats, many nodes).
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
            if (!done) { hsh = hash_read(row, w, L); j = slot_of(hsh, ix.slot_mask); }
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

// threads per deferred read in the list-driven kernels: a power of two in [1, 32]
static uint32_t pow2_spread(int64_t v) {
    uint32_t s = 1;
    while (s < 32 && (int64_t)s * 2 <= v) s *= 2;
    return s;
}

// Work lists of the list-driven tiers for a chunk of up to n_reads reads (reserved and emptied before
// the first kernel that appends to them).
int map_prepare_lists(Ctx* c, uint64_t n_reads) {
    if (n_reads > 0xFFFFFFFFull) { set_error("more than 2^32 reads in one chunk"); return VSPE_ERR_LIMIT; }
    VSPE_TRY(c->worklist.reserve(n_reads + 1));
    VSPE_TRY(c->defer_list.reserve(3 * n_reads + 3));
    VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_WORK, 0, 8, c->stream));
    VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_DEFER, 0, 8, c->stream));
    VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_DEFER2, 0, 8, c->stream));
    VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_WORK2, 0, 8, c->stream));
    return VSPE_OK;
}

// pre_listed: the reads to map are the counters[CNT_DEFER] entries of c->defer_list (written by k_scan_map
// after map_prepare_lists); the walk stages 1 / 1b already ran inside that kernel.
static int launch_map_fast(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                           const uint32_t* d_rows, const uint32_t* d_hdr, uint32_t row_words, uint32_t cap,
                           uint64_t n_reads, ReadSlot* d_slots, bool pre_listed = false) {
    if (n_reads == 0) return VSPE_OK;
    if (n_reads > 0xFFFFFFFFull) { set_error("more than 2^32 reads in one chunk"); return VSPE_ERR_LIMIT; }
    if (!pre_listed) {
        VSPE_TRY(c->worklist.reserve(n_reads));
        VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_WORK, 0, 8, c->stream));
    }
    const uint32_t grid = (uint32_t)((n_reads + MF_THREADS - 1) / MF_THREADS);
    IndexView v = c->index.view();
    const uint32_t* in_list = nullptr;
    const unsigned long long* in_count = nullptr;
    const uint32_t* to_generic = c->worklist.p;                 // reads the ASCII tier must map
    const unsigned long long* to_generic_n = c->counters.p + CNT_WORK;
    if (d_rows && !pre_listed) {
        VSPE_TRY(c->defer_list.reserve(3 * n_reads));
        VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_DEFER, 0, 8, c->stream));
        VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_DEFER2, 0, 8, c->stream));
        VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_WORK2, 0, 8, c->stream));
    }
    // lean or general walk kernels (forced by the map_general option, else from the last launch's CNT_BIG)
    const bool general = c->opt_map_general >= 0 ? c->opt_map_general != 0 : c->map_general;
    if (pre_listed) {
        in_list = c->defer_list.p;
        in_count = c->counters.p + CNT_DEFER;
    } else if (d_rows && !c->opt_single_map) {
        // stage 1: the clean-pass kernel; what it defers goes through the full kernel
#define VSPE_M1(S, LP, FL, GN) k_map_first<S, LP, FL, GN><<<grid, MF_THREADS, 0, c->stream>>>(v, d_rows, d_hdr, row_words, n_reads, d_slots, \
                                                                                        c->defer_list.p, c->counters.p)
#define VSPE_M1S(FL, GN) do { if (cap <= 160) VSPE_M1(13, 16, FL, GN); else if (cap <= 256) VSPE_M1(19, 16, FL, GN); else VSPE_M1(23, 32, FL, GN); } while (0)
        VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_BIG, 0, 8, c->stream));
        if (c->opt_flat_walk) { if (general) VSPE_M1S(true, true); else VSPE_M1S(true, false); }
        else { if (general) VSPE_M1S(false, true); else VSPE_M1S(false, false); }
        c->big_sampled = std::min<uint64_t>(n_reads, (uint64_t)BIG_SAMPLE_BLOCKS * MF_THREADS);
        c->big_pending = true;
#undef VSPE_M1S
#undef VSPE_M1
        VSPE_LAUNCH_CHECK(c);
        in_list = c->defer_list.p;
        in_count = c->counters.p + CNT_DEFER;
        if (c->index.has_subst && !c->opt_no_second) {
            // stage 1b: one-error-tolerant walk on the deferred reads; its leftovers go to stage 2
            uint32_t* list1b = c->defer_list.p + 2 * n_reads;
            const uint32_t spread2 = pow2_spread(c->opt_second_spread);
            const uint32_t grid_s = (uint32_t)std::min<uint64_t>((n_reads * spread2 + MF_THREADS - 1) / MF_THREADS, (uint64_t)c->sm_count * 16);
#define VSPE_M2(S, GN) k_map_second<S, GN><<<grid_s, MF_THREADS, 0, c->stream>>>(v, d_rows, d_hdr, row_words, c->defer_list.p, \
                                                                           c->counters.p + CNT_DEFER, spread2, d_slots, list1b, c->counters.p + CNT_DEFER2)
            if (general) { if (cap <= 160) VSPE_M2(13, true); else if (cap <= 256) VSPE_M2(19, true); else VSPE_M2(23, true); }
            else { if (cap <= 160) VSPE_M2(13, false); else if (cap <= 256) VSPE_M2(19, false); else VSPE_M2(23, false); }
#undef VSPE_M2
            VSPE_LAUNCH_CHECK(c);
            in_list = list1b;
            in_count = c->counters.p + CNT_DEFER2;
        }
    }
    // what stage 3 reads: the reads stage 2 gave up on -- or, without stage 2, everything stage 1 deferred
    const uint32_t* win_list = c->worklist.p;
    const unsigned long long* win_count = c->counters.p + CNT_WORK;
    const bool skip_fast = d_rows && in_list && !c->opt_fast_tier;
    if (skip_fast) {
        win_list = in_list;
        win_count = in_count;
    }
    // stage 2: the full seed-and-extend kernel (on the deferred reads, or on everything)
    const uint32_t list_spread = pow2_spread(c->opt_list_spread);
    const uint32_t grid2 = in_list ? (uint32_t)std::min<uint64_t>((n_reads * list_spread + MF_THREADS - 1) / MF_THREADS, (uint64_t)c->sm_count * 8) : grid;
#define VSPE_MF(S, LP, PK) k_map_fast<S, LP, PK><<<grid2, MF_THREADS, 0, c->stream>>>(v, d_buf, d_seq_start, d_seq_end, d_rows, d_hdr, \
                                                                               row_words, n_reads, in_list, in_count, list_spread, d_slots, \
                                                                               c->worklist.p, c->counters.p)
    if (skip_fast) {
        // (stage 2 skipped: the deferred reads go straight to the all-windows kernel)
    } else if (d_rows) {
        if (cap <= 160) VSPE_MF(13, 16, true); else if (cap <= 256) VSPE_MF(19, 16, true); else VSPE_MF(23, 32, true);
        VSPE_LAUNCH_CHECK(c);
    } else {
        if (cap <= 160) VSPE_MF(13, 16, false); else if (cap <= 256) VSPE_MF(19, 16, false); else VSPE_MF(23, 32, false);
        VSPE_LAUNCH_CHECK(c);
    }
#undef VSPE_MF
    if (d_rows) {
        // stage 3: what stage 2 could not prove (repeats, > 16 nodes): one warp per read, every
        // window looked up, exact for any postings multiplicity; only reads that do not fit the
        // 2-bit rows at all (non-ACGT, very long) are left for the ASCII tier
        if (!c->spill.p) VSPE_TRY(c->spill.reserve(4u << 20));
        uint32_t* list2 = c->defer_list.p + n_reads;
        const uint32_t wgrid = (uint32_t)std::min<uint64_t>((n_reads + MW_WARPS - 1) / MW_WARPS, (uint64_t)c->sm_count * 16);
#define VSPE_MW(S) k_map_windows<S><<<wgrid, MW_WARPS * 32, 0, c->stream>>>(v, d_rows, d_hdr, row_words, win_list, \
                                                                      win_count, d_slots, list2, \
                                                                      c->counters.p + CNT_WORK2, c->spill.p, c->spill.cap, c->counters.p)
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

// After a stream synchronisation: look at how many of the sampled reads of the last k_map_first
// launch had more than FL_MAX stretches and pick the walk-kernel variant for the next launch.
// Both variants are exact; this only moves reads between tiers.
int adapt_map_variant(Ctx* c) {
    if (!c->big_pending) return VSPE_OK;
    unsigned long long big = 0;
    VSPE_CUDA(cudaMemcpy(&big, c->counters.p + CNT_BIG, 8, cudaMemcpyDeviceToHost));
    c->big_pending = false;
    if (c->big_sampled) c->map_general = big * 50 > c->big_sampled;      // more than 2 %
    return VSPE_OK;
}

// packed-row capacity (bases) by the read length seen in the first records; longer reads bail
uint32_t map_fast_cap(uint32_t hint) { return hint <= 160 ? 160 : hint <= 256 ? 256 : 320; }

int map_reads_fast(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                   uint64_t n_reads, ReadSlot* d_slots) {
    if (c->index.split_len > 320) return map_reads_generic(c, d_buf, d_seq_start, d_seq_end, n_reads, d_slots);
    return launch_map_fast(c, d_buf, d_seq_start, d_seq_end, nullptr, nullptr, 0, map_fast_cap(c->read_len_hint), n_reads, d_slots);
}

int map_reads_packed(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                     const uint32_t* d_rows, const uint32_t* d_hdr, uint32_t row_words, uint32_t cap,
                     uint64_t n_reads, ReadSlot* d_slots) {
    return launch_map_fast(c, d_buf, d_seq_start, d_seq_end, d_rows, d_hdr, row_words, cap, n_reads, d_slots);
}

// the reads k_scan_map left unresolved (listed in c->defer_list): full seed-and-extend kernel, then the
// all-windows kernel, then the ASCII tier; results go to d_slots[r] for the listed r
int map_reads_deferred(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                       const uint32_t* d_rows, const uint32_t* d_hdr, uint32_t row_words, uint32_t cap,
                       uint64_t n_reads_cap, ReadSlot* d_slots) {
    return launch_map_fast(c, d_buf, d_seq_start, d_seq_end, d_rows, d_hdr, row_words, cap, n_reads_cap, d_slots, true);
}

}  // namespace vspe
