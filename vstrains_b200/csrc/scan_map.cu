// scan_map.cu -- K1 + K2 (k_scan_rows) and the first tier of K4 (k_walk): the default path of a
// FASTQ chunk from raw bytes to node-list handles.
//
// Replaces `readlines()` + `[s[:-1] ...]` (reference utils/VStrains_PE_Inference.py:149-159), the
// per-character work of `fseq.count("N")` / k-mer slicing (:160, :25) and single_end_read_mapping
// (:16-48) for every read whose result the walk can PROVE; the rest (about 3 % on the bench
// workloads: two or more sequencing errors, repeats, non-ACGT characters, very long reads) is
// listed for the list-driven tiers of map_fast.cu / map_generic.cu.
//
// k_scan_rows -- ONE pass over the bytes, per 40 KiB tile (one CTA of 10 warps, 4 CTAs per SM, tiles
// handed out by a ticket counter):
//   1. one elected thread issues TMA bulk copies (cp.async.bulk, mbarrier complete_tx) of the tile
//      + a 16-byte front margin + a 512-byte back margin into shared memory;
//   2. every lane tests its 16-byte vectors for bytes < 0x10 or >= 0x80 (two instructions per
//      32-bit word); candidate vectors go to a per-warp queue (warp ballots) and only they get exact
//      terminator masks (universal newlines: '\n', "\r\n" once, lone '\r'); a warp scan ranks them;
//   3. warp totals + one block exchange give the tile's terminator count; a decoupled look-back
//      over the tiles' status words gives the line number of the tile's first line;
//   4. terminators with line%4==0 start a sequence line, line%4==1 end it -> per-tile read table;
//   5. one thread per read packs the read the tile owns (its sequence line STARTS here) to 2 bits/base
//      straight from the tile, 16 bases per step (SIMD-in-word ACGT validity test, 'N' flag), into a
//      shared-memory row; each warp then copies its 32 rows to HBM with coalesced stores
//      (48 / 64 / 80 bytes per read + a header word + the byte range).
// k_walk -- one thread per read, 128-thread blocks, 8 blocks per SM (the walk is a chain of dependent
// L2 accesses: it wants many warps and a large L1, which is why it is NOT fused into the scan kernel --
// the fused variant was built, bit-exact, and 3x slower: 15 warps per SM, 31 % issue slots):
//   6. seed window 0 (hash + probe + verify + uniq bit; on a miss the reverse complement is seeded from
//      the other end), then a flat loop whose every turn compares 32 bases + 32 uniq bits on the current
//      diagonal and, when the stretch is complete, books it and steps to the successor strand, whose
//      table entry (text position, strand, strand end, node length: 16 bytes) was requested when the
//      stretch was entered.  ONE mismatching base is tolerated when the substitution-hit bit proves that
//      the windows covering it have no posting.  The saturation predicate (:36-47, integer form) is
//      applied as each stretch is booked;
//   7. the kept node list is interned (link.cuh) and its handle stored; an unresolved read is listed.
// Why this is exact: see map_fast.cu (a window is counted without a table access only if its text
// equality and the uniq bit of that text window were both checked; it is skipped only if the
// index build already looked that k-mer up and found nothing).
#include <algorithm>

#include "link.cuh"
#include "map_common.cuh"

namespace vspe {

#ifndef VSPE_SM_WARPS
#define VSPE_SM_WARPS 10
#endif
#ifndef VSPE_SM_ITERS
#define VSPE_SM_ITERS 8
#endif
#ifndef VSPE_SM_MINB
#define VSPE_SM_MINB 4
#endif
static constexpr int SM_WARPS = VSPE_SM_WARPS;
static constexpr int SM_THREADS = SM_WARPS * 32;
static constexpr int SM_ITERS = VSPE_SM_ITERS;                   // 512-byte warp rows per warp
static constexpr int SM_TILE = SM_WARPS * SM_ITERS * 32 * 16;    // 40 KiB
static constexpr int SM_FRONT = 16;                              // bytes kept before the tile
static constexpr int SM_BACK = 512;                              // bytes kept after the tile (>= longest packed read + 1)
static constexpr int SM_MAXREC = SM_TILE / 64 - 64;              // reads a tile may own (else the chunk takes the plain path): 576
static constexpr int SM_QCAP = 12 * SM_ITERS;                    // per warp: vectors that may hold a terminator
static constexpr int WK_THREADS = 128;                           // k_walk: reads per block
static constexpr int SM_MAXST = 16;                              // stretches (nodes with hits) per read in the walk

#define LB_AGG (1ull << 62)
#define LB_INC (2ull << 62)
#define LB_VAL ((1ull << 62) - 1)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t movemask4b(uint32_t cmp) { return ((cmp & 0x80808080u) * 0x00204081u) >> 28; }

// terminator / crlf masks of one 16-byte vector held in registers; `valid` = bitmask of the
// bytes that belong to the buffer; next/prev = the neighbouring bytes (0 if outside)
__device__ __forceinline__ void masks_from_vec(uint4 v, uint32_t valid, uint32_t next_byte, uint32_t prev_byte,
                                               bool& non_ascii, uint32_t& term, uint32_t& crlf) {
    term = 0;
    crlf = 0;
    if (valid == 0xFFFFu && ((v.x | v.y | v.z | v.w) & 0x80808080u)) non_ascii = true;
    uint32_t nl = movemask4b(__vcmpeq4(v.x, 0x0A0A0A0Au)) | (movemask4b(__vcmpeq4(v.y, 0x0A0A0A0Au)) << 4) |
                  (movemask4b(__vcmpeq4(v.z, 0x0A0A0A0Au)) << 8) | (movemask4b(__vcmpeq4(v.w, 0x0A0A0A0Au)) << 12);
    // '\r' is rare: one zero-byte test over the four words decides whether its mask is needed at all
    const uint32_t zx = v.x ^ 0x0D0D0D0Du, zy = v.y ^ 0x0D0D0D0Du, zz = v.z ^ 0x0D0D0D0Du, zw = v.w ^ 0x0D0D0D0Du;
    uint32_t cr = 0;
    if ((((zx - 0x01010101u) & ~zx) | ((zy - 0x01010101u) & ~zy) | ((zz - 0x01010101u) & ~zz) | ((zw - 0x01010101u) & ~zw)) & 0x80808080u)
        cr = movemask4b(__vcmpeq4(v.x, 0x0D0D0D0Du)) | (movemask4b(__vcmpeq4(v.y, 0x0D0D0D0Du)) << 4) |
             (movemask4b(__vcmpeq4(v.z, 0x0D0D0D0Du)) << 8) | (movemask4b(__vcmpeq4(v.w, 0x0D0D0D0Du)) << 12);
    if (valid != 0xFFFFu) {
        const uint32_t na = movemask4b(v.x) | (movemask4b(v.y) << 4) | (movemask4b(v.z) << 8) | (movemask4b(v.w) << 12);
        if (na & valid) non_ascii = true;
        nl &= valid;
        cr &= valid;
    }
    term = nl | (cr & ~((nl >> 1) | (next_byte == '\n' ? 0x8000u : 0u)));
    crlf = nl & ((cr << 1) | (prev_byte == '\r' ? 1u : 0u));
}

struct ScanMapArgs {
    const uint8_t* buf;              // chunk start (may be misaligned)
    uint64_t n;                      // chunk bytes
    uint32_t head;                   // address of buf mod 16
    uint32_t n_tiles;
    unsigned long long* status;      // look-back words, one per tile (zeroed)
    unsigned int* ticket;
    unsigned long long* total_out;   // terminators in the chunk
    uint64_t line_base;              // lines before this chunk
    uint64_t rec_first;              // record number of slot 0 of the outputs
    uint64_t n_slots;                // capacity of the per-read outputs
    uint64_t* seq_start;             // [n_slots] chunk-relative byte range of the sequence line
    uint64_t* seq_end;
    uint32_t* rows;                  // [n_slots][row_words] packed read
    uint32_t* hdr;                   // [n_slots] rlen | flags << 24
    uint32_t row_words;              // 12, 16 or 20
    uint32_t cap;                    // longest read (bases) a packed row holds
    unsigned long long* counters;
};

// The walk of step 6.  row: this thread's packed read (STRIDE words, zero padded); lst: its node list
// column (entry i at lst[i * LS]).  Returns true when every window of the read is accounted for;
// n_kept nodes that pass the saturation predicate are then at the front of the column.
template <int STRIDE, int LS>
__device__ __forceinline__ bool walk_read(const IndexView& ix, uint32_t* row, const uint32_t rlen, uint32_t* lst, uint32_t& n_kept) {
    constexpr int NW = STRIDE - 3;
    const uint32_t L = ix.split_len;
    const int npos = (int)(rlen - L + 1);
    bool mirror = false;
    uint32_t tp = NONE32, node = 0;
    int pr = probe_window(ix, row, 0, tp, node);
    if (pr == PROBE_MISS && ix.subst != nullptr) {
        // error in the first split_len bases: seed window 0 of the reverse complement (the other end)
        revcomp_row<NW>(row, rlen);
        mirror = true;
        pr = probe_window(ix, row, 0, tp, node);
    }
    // (PROBE_MULTI, or both ends miss: a real complication, the next tier decides)
    bool running = pr == PROBE_UNIQUE, resolved = false;
    bool err = false;
    int e = 0;
    uint32_t rb = 0;
    uint32_t i0 = 0, p = L, q = 0, lim = 0, nn = 0, n_front = 0, nlen = 0;
    unsigned long long seen = 0, seen2 = 0;                    // two 64-bit filters over the nodes booked so far
    int delta = 0;
    uint4 nxt = make_uint4(NONE32, 0, 0, 0);                   // successor entry of the current strand for the read's base at lim
    // enter the strand q that holds text position tp_ (its end: send_, its node length: nlen_) at window i0;
    // false if the error's windows cannot be proven to miss there
    auto enter = [&](uint32_t tp_, uint32_t q_, uint32_t send_, uint32_t nlen_) -> bool {
        q = q_;
        node = q_ >> 1;
        nlen = nlen_;
        delta = (int)tp_ - (int)i0;
        lim = min(rlen, (uint32_t)((int)send_ - delta));       // read position where the strand ends
        if (lim < rlen) {
            // the strand ends before the read: request the successor entry for the read's next base NOW, it is
            // consumed when the stretch is booked (the chunk compares in between hide the access)
            const uint32_t b = (row[lim >> 4] >> ((lim & 15) * 2)) & 3u;
            nxt = __ldg(ix.succ16 + 4 * (size_t)q + b);
        }
        // a strand entered after the error still holds windows covering it if it starts at or before e
        if (err && (int)i0 <= e) {
            const uint32_t te = (uint32_t)(e + delta);
            if ((__ldg(ix.subst + (te >> 3)) >> (4 * (te & 7) + rb)) & 1u) return false;
        }
        return true;
    };
    if (running) {
        const uint4 nr = __ldg(ix.node_rec + node);            // {forward start, rc start, end, node length}
        const bool rcs = tp >= nr.y;
        if (!enter(tp, 2 * node + (rcs ? 1u : 0u), rcs ? nr.z : nr.y, nr.w)) running = false;
    }
    while (running) {
        if (p < lim) {
            const uint32_t n = min(32u, lim - p);
            uint64_t x = read64(row, p) ^ extract64(ix.text, (uint64_t)((int)p + delta));
            if (n < 32) x &= (1ull << (2 * n)) - 1;
            bool ok = true;
            if (x) {
                const uint32_t off = (uint32_t)(__ffsll((long long)x) - 1) >> 1;
                if (err || ix.subst == nullptr || (x & ~(3ull << (2 * off)))) ok = false;      // second mismatch
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
                if (nn == (uint32_t)SM_MAXST) { running = false; break; }
                // a node met twice (cyclic graph) needs its hits merged by the next tier: two 64-bit filters gate
                // the exact comparison with the nodes booked so far
                const uint32_t hb = (node * 0x9E3779B1u) >> 26, hb2 = (node * 0x85EBCA77u) >> 26;
                if (((seen >> hb) & (seen2 >> hb2)) & 1ull) {
                    bool dup = false;
                    for (uint32_t i = 0; i < n_front; i++) dup |= lst[i * LS] == node;
                    for (uint32_t i = 0; i < nn - n_front; i++) dup |= lst[(SM_MAXST - 1 - i) * LS] == node;
                    if (dup) { running = false; break; }
                }
                seen |= 1ull << hb;
                seen2 |= 1ull << hb2;
                const uint32_t v = (uint32_t)(c1 + c2), kmin = (uint32_t)(mirror ? npos - 1 - last_hit : first_hit);
                // kept nodes fill the column from the front, the others from the back
                if (keep_node_f(v, kmin, nlen, rlen, L)) lst[(n_front++) * LS] = node;
                else lst[(SM_MAXST - 1 - (nn - n_front)) * LS] = node;
                nn++;
            }
            if (lim >= rlen) { resolved = true; running = false; break; }
            if (nxt.x == NONE32) { running = false; break; }       // no unique successor window for that base
            i0 = lim - L + 1;
            p = lim + 1;
            if (!enter(nxt.x, nxt.y, nxt.z, nxt.w)) { running = false; break; }
        }
    }
    if (!resolved) return false;
    n_kept = n_front;
    return true;
}

// One thread per read of the chunk: pair-skipping classes, else the walk; what it cannot prove is listed.
struct WalkArgs {
    const uint32_t* rows;
    const uint32_t* hdr;
    uint32_t row_words;
    uint64_t n_slots;                  // capacity of the per-read arrays
    const unsigned long long* total;   // terminators of the chunk (written by k_scan_rows)
    uint64_t line_base, rec_first;
    uint32_t* handles;
    uint32_t* defer_list;
    unsigned long long* defer_count;
    unsigned long long* counters;
};

#ifndef VSPE_WK_MINB
#define VSPE_WK_MINB 10
#endif
template <int STRIDE>
__global__ void __launch_bounds__(WK_THREADS, VSPE_WK_MINB)
k_walk(const WalkArgs a, const IndexView ix, const LinkView lv) {
    __shared__ uint32_t s_rows[WK_THREADS * STRIDE];
    __shared__ uint32_t s_lst[SM_MAXST * WK_THREADS];
    // sequence lines of the chunk = #{l in [line_base, line_base + total) : l % 4 == 1}
    const uint64_t n_lines1 = (uint64_t)(a.line_base + *a.total + 2) / 4 - a.rec_first;
    const uint64_t n_reads = n_lines1 < a.n_slots ? n_lines1 : a.n_slots;
    const uint64_t r = (uint64_t)blockIdx.x * WK_THREADS + threadIdx.x;
    if (r >= n_reads) return;
    if (r == 0) atomicAdd(&a.counters[CNT_FAST], (unsigned long long)n_reads);
    const uint32_t h = __ldg(a.hdr + r);
    const uint32_t rlen = h & 0xFFFFFF, L = ix.split_len;
    uint32_t handle = H_PENDING;
    bool defer = (h & (PH_LONG | PH_BAD)) != 0;
    if (!(h & PH_LONG)) {
        if (h & PH_N) handle = H_N;                                    // 'N' before the length (PE_Inference.py:160-163)
        else if (rlen < L) handle = H_SHORT;
    }
    if (handle == H_PENDING && !defer) {
        uint32_t* row = s_rows + threadIdx.x * STRIDE;
        constexpr int NW = STRIDE - 3, XW = (NW + 3) / 4 * 4;
        const uint4* src = reinterpret_cast<const uint4*>(a.rows + r * a.row_words);
        uint32_t x[XW];
#pragma unroll
        for (int q = 0; q < XW / 4; q++) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if ((uint32_t)(4 * q) < a.row_words) v = __ldg(src + q);
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int w = 0; w < NW; w++) row[w] = x[w];
        row[NW] = 0; row[NW + 1] = 0; row[NW + 2] = 0;
        uint32_t n_kept = 0;
        if (walk_read<STRIDE, WK_THREADS>(ix, row, rlen, s_lst + threadIdx.x, n_kept)) handle = intern_list(lv, n_kept, s_lst + threadIdx.x, WK_THREADS);
        else defer = true;
    }
    if (handle == H_PENDING && defer) a.defer_list[atomicAdd(a.defer_count, 1ull)] = (uint32_t)r;
    a.handles[r] = handle;
}

static constexpr uint32_t SM_QUEUE_BYTES = SM_WARPS * SM_QCAP * 8;
static constexpr uint32_t SM_SMEM = SM_FRONT + SM_TILE + SM_BACK + SM_MAXREC * 8 + SM_QUEUE_BYTES + 64;

template <int RW>                                                // row words in HBM: 12, 16 or 20 (a multiple of 4 >= cap / 16)
__global__ void __launch_bounds__(SM_THREADS, VSPE_SM_MINB)
k_scan_rows(const ScanMapArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_bytes = smem;                                           // [SM_FRONT + SM_TILE + SM_BACK]
    uint32_t* s_rs = reinterpret_cast<uint32_t*>(smem + SM_FRONT + SM_TILE + SM_BACK);   // read start (tile-relative)
    uint32_t* s_re = s_rs + SM_MAXREC;                                 // read end
    uint32_t* s_un = s_re + SM_MAXREC;                                 // candidate queues
    uint32_t* s_qmk = s_un;                                            // [SM_WARPS][SM_QCAP] term | crlf << 16
    uint16_t* s_qid = reinterpret_cast<uint16_t*>(s_qmk + SM_WARPS * SM_QCAP);   // vector index in the tile
    uint16_t* s_qrk = s_qid + SM_WARPS * SM_QCAP;                      // rank of the vector's first terminator in the warp
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_wtot[SM_WARPS];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_excl;
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        s_tile = atomicAdd(a.ticket, 1u);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    // aligned coordinates: byte `off` of the aligned stream is buffer position off - head
    const uint64_t A = ((uint64_t)a.head + a.n + 15) & ~15ull;         // aligned stream length
    const uint64_t t_lo = (uint64_t)tile * SM_TILE;
    const uint64_t ld_lo = t_lo >= SM_FRONT ? t_lo - SM_FRONT : 0;
    const uint64_t ld_hi = min(A, t_lo + SM_TILE + SM_BACK);
    const uint32_t s_off0 = tile == 0 ? SM_FRONT : 0;                  // where ld_lo lands in s_bytes
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)(ld_hi - ld_lo);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(bytes) : "memory");
        const uint8_t* src = a.buf - a.head + ld_lo;
        uint32_t done = 0;
        while (done < bytes) {
            const uint32_t part = min(bytes - done, 16384u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(s_bytes + s_off0 + done)),
                         "l"(__cvta_generic_to_global(src + done)), "r"(part), "r"(smem_u32(&s_bar))
                         : "memory");
            done += part;
        }
    }
    {   // wait for the bytes
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(smem_u32(&s_bar)), "r"(0) : "memory");
        }
    }
    // tile byte j (0 <= j < SM_TILE) lives at s_bytes[SM_FRONT + j]; its buffer position is t_lo + j - head
    const uint8_t* tb = s_bytes + SM_FRONT;
    const int64_t pos0 = (int64_t)t_lo - a.head;                       // buffer position of tile byte 0
    const int64_t nn = (int64_t)a.n;
    auto byte_at = [&](int64_t j) -> uint32_t {                         // tile-relative byte, 0 outside the buffer
        const int64_t p = pos0 + j;
        return (p >= 0 && p < nn) ? tb[j] : 0u;
    };

    uint16_t* q_id = s_qid + wib * SM_QCAP;
    uint32_t* q_mk = s_qmk + wib * SM_QCAP;
    uint16_t* q_rk = s_qrk + wib * SM_QCAP;
    const bool interior = pos0 >= 1 && pos0 + SM_TILE + 16 <= nn;   // CTA-uniform
    const uint32_t lt = (1u << lane) - 1;
    uint32_t qn = 0, wcount = 0;
    bool bad = false, q_over = false;
    // ---- M1: which 16-byte vectors can hold a terminator?  ('\n' and '\r' are < 0x10) -----------
    // Warp w owns tile bytes [w*4K, (w+1)*4K) as 8 coalesced 512-byte rows.  (x - 0x10) | x has bit 7
    // set in every byte that is < 0x10 or >= 0x80 (a borrow can only add false positives next to a true
    // one), so two instructions per word decide.  Candidate vectors are appended, in (row, lane) order,
    // to the warp's queue: everything after this loop runs on a dense list.
#pragma unroll
    for (int it = 0; it < SM_ITERS; it++) {
        const uint32_t vid = (wib * SM_ITERS + it) * 32 + lane;       // vector index inside the tile
        const uint4 v = *reinterpret_cast<const uint4*>(tb + vid * 16);
        const uint32_t t = ((v.x - 0x10101010u) | v.x) | ((v.y - 0x10101010u) | v.y) | ((v.z - 0x10101010u) | v.z) |
                           ((v.w - 0x10101010u) | v.w);
        bool cand = (t & 0x80808080u) != 0;
        if (!interior) {                                               // first / last tile of the chunk
            const int64_t p = pos0 + (int64_t)vid * 16;
            const bool full = p >= 0 && p + 16 <= nn;
            if (!full) cand = p < nn && p + 16 > 0;                    // partial vector: M2 masks the outside bytes
        }
        const uint32_t bm = __ballot_sync(0xFFFFFFFFu, cand);
        if (cand) {
            const uint32_t at = qn + __popc(bm & lt);
            if (at < (uint32_t)SM_QCAP) q_id[at] = (uint16_t)vid;
        }
        qn += __popc(bm);
    }
    q_over = qn > (uint32_t)SM_QCAP;
    if (q_over) qn = SM_QCAP;
    __syncwarp();
    // ---- M2: exact terminator / crlf masks of the candidates + their ranks inside the warp -----
    for (uint32_t i0 = 0; i0 < qn; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint32_t term = 0, crlf = 0;
        if (i < qn) {
            const uint32_t vid = q_id[i];
            const uint32_t j = vid * 16;
            const uint4 v = *reinterpret_cast<const uint4*>(tb + j);
            uint32_t valid = 0xFFFFu;
            if (!interior) {
                const int64_t p = pos0 + j;
                if (p < 0) valid &= 0xFFFFu << (uint32_t)(-p);
                if (p + 16 > nn) valid &= 0xFFFFu >> (uint32_t)(p + 16 - nn);
            }
            masks_from_vec(v, valid, interior ? (uint32_t)tb[j + 16] : byte_at((int64_t)j + 16),
                           interior ? (uint32_t)tb[(int)j - 1] : byte_at((int64_t)j - 1), bad, term, crlf);
            q_mk[i] = term | (crlf << 16);
        }
        const uint32_t c = __popc(term);
        uint32_t inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        if (i < qn) q_rk[i] = (uint16_t)(wcount + inc - c);
        wcount += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    if (bad) atomicOr(&a.counters[CNT_ERR], (unsigned long long)ERRF_NON_ASCII);
    if (lane == 0) s_wtot[wib] = wcount | (q_over ? 0x80000000u : 0u);
    for (uint32_t i = threadIdx.x; i < (uint32_t)SM_MAXREC; i += SM_THREADS) s_re[i] = 0xFFFFFFFFu;   // "line end not seen in this tile"
    __syncthreads();
    uint32_t tile_total = 0, warp_base = 0;
    bool any_over = false;
#pragma unroll
    for (int w = 0; w < SM_WARPS; w++) {
        const uint32_t x = s_wtot[w];
        any_over |= (x >> 31) != 0;
        if (w < (int)wib) warp_base += x & 0x7FFFFFFFu;
        tile_total += x & 0x7FFFFFFFu;
    }
    // ---- decoupled look-back (warp 0), 128 predecessors per hop -------------------------------
    if (wib == 0) {
        volatile unsigned long long* vs = a.status;
        if (tile == 0) {
            if (lane == 0) { vs[0] = LB_INC | tile_total; s_excl = 0; }
        } else {
            if (lane == 0) vs[tile] = LB_AGG | tile_total;
            unsigned long long excl = 0;
            int64_t look = (int64_t)tile - 1;                      // closest predecessor not yet summed
            while (true) {
                unsigned long long st[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int64_t idx = look - 4 * (int64_t)lane - k;
                    st[k] = idx >= 0 ? vs[idx] : LB_INC;
                }
                while (true) {
                    bool missing = false;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if ((st[k] >> 62) == 0) {
                            st[k] = vs[look - 4 * (int64_t)lane - k];
                            missing |= (st[k] >> 62) == 0;
                        }
                    }
                    if (!__any_sync(0xFFFFFFFFu, missing)) break;
                }
                // this lane: sum up to and including its closest inclusive word, if it has one
                unsigned long long c = 0;
                bool has_inc = false;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (!has_inc) {
                        c += st[k] & LB_VAL;
                        has_inc = (st[k] >> 62) == 2;
                    }
                }
                const uint32_t inc = __ballot_sync(0xFFFFFFFFu, has_inc);
                const int first = inc ? __ffs((int)inc) - 1 : 32;
                if ((int)lane > first) c = 0;
                for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
                excl += c;
                if (inc) break;
                look -= 128;
            }
            if (lane == 0) { vs[tile] = LB_INC | (excl + tile_total); s_excl = excl; }
        }
        if (lane == 0 && tile == a.n_tiles - 1) *a.total_out = s_excl + tile_total;
    }
    __syncthreads();
    const uint64_t base = a.line_base + s_excl;                        // line number of the tile's first line
    // reads owned by this tile: sequence lines that START here = header terminators (line%4==0)
    // in the tile; record numbers are consecutive from r_own0
    const uint64_t r_own0 = (base + 3) >> 2;
    const uint64_t last_line = base + tile_total;                      // one past the tile's last terminator
    uint32_t n_own = (uint32_t)(((last_line + 3) >> 2) - r_own0);      // #{l in [base, last_line) : l%4 == 0}
    const bool chunk_starts_in_seq = tile == 0 && (a.line_base & 3) == 1;   // chunk begins with a sequence line
    const uint32_t shift = chunk_starts_in_seq ? 1u : 0u;              // that read becomes local index 0
    const uint32_t n_local = n_own + shift;
    const bool too_many = n_local > (uint32_t)SM_MAXREC || any_over;
    if (too_many) {
        if (threadIdx.x == 0) atomicOr(&a.counters[CNT_ERR], (unsigned long long)ERRF_TILE_FULL);
        return;
    }
    if (chunk_starts_in_seq && threadIdx.x == 0) s_rs[0] = (uint32_t)a.head;   // buffer position 0, tile-relative
    // ---- emission: every queue entry knows its rank -> line numbers -> read table ---------------
    for (uint32_t i = lane; i < qn; i += 32) {
        uint32_t mask = q_mk[i] & 0xFFFFu;
        const uint32_t crlf = q_mk[i] >> 16;
        const uint32_t j0 = (uint32_t)q_id[i] * 16;
        uint64_t line = base + warp_base + q_rk[i];
        while (mask) {
            const int k = __ffs((int)mask) - 1;
            mask &= mask - 1;
            const uint32_t j = j0 + k;                                   // tile-relative terminator position
            const uint32_t phase = (uint32_t)line & 3;
            if (phase == 0) {
                s_rs[(uint32_t)((line >> 2) - r_own0) + shift] = j + 1;
            } else if (phase == 1) {
                const uint64_t r = line >> 2;
                const uint32_t e = ((crlf >> k) & 1) ? j - 1 : j;
                if (r >= r_own0) s_re[(uint32_t)(r - r_own0) + shift] = e;
                else if (chunk_starts_in_seq && r + 1 == r_own0) s_re[0] = e;
                // (a sequence line that started in the previous tile is packed by that tile)
            }
            line++;
        }
    }
    __syncthreads();
    const uint64_t r_loc0 = r_own0 - shift;                            // record number of local index 0
    // ---- pack: one thread per read, the row goes from registers to HBM with 16-byte stores -- no block
    // barrier from here on.  16 bases = four words per step: code = (ascii >> 1) & 3; the only byte with
    // code k is "ACTG"[k] = 0x41 + 2k (+15 when k == 2), which is the validity test.
    constexpr int NW = RW == 12 ? 10 : RW;                             // row words that can hold bases (cap / 16)
    for (uint32_t li = threadIdx.x; li < n_local; li += SM_THREADS) {
        const uint64_t slot = r_loc0 + li - a.rec_first;
        if (slot >= a.n_slots) { atomicOr(&a.counters[CNT_ERR], (unsigned long long)ERRF_SLOTS_FULL); continue; }
        const uint32_t st = s_rs[li];
        uint32_t en = s_re[li];
        uint32_t h = 0;
        if (en == 0xFFFFFFFFu) {
            // the line ends beyond the tile: first '\n' or '\r' in the back margin, if any (four bytes per step:
            // only a word with a byte < 0x10 is looked at byte by byte)
            for (uint32_t j = SM_TILE; j < (uint32_t)(SM_TILE + SM_BACK) && en == 0xFFFFFFFFu; j += 4) {
                const uint32_t w = *reinterpret_cast<const uint32_t*>(tb + j);
                if (!(((w - 0x10101010u) | w) & 0x80808080u) && pos0 + (int64_t)j + 4 <= nn) continue;
                for (uint32_t k = 0; k < 4 && pos0 + (int64_t)(j + k) < nn; k++) {
                    const uint32_t ch = (w >> (8 * k)) & 0xFF;
                    if (ch == '\n' || ch == '\r') { en = j + k; break; }
                }
                if (pos0 + (int64_t)j + 4 > nn) break;
            }
            if (en == 0xFFFFFFFFu) h |= PH_LONG;
        }
        uint32_t rlen = (h & PH_LONG) ? 0u : en - st;
        if (rlen > a.cap) { h |= PH_LONG; rlen = 0; }
        const uint32_t a0 = SM_FRONT + st;                              // offset of the first base in s_bytes
        const uint32_t* p = reinterpret_cast<const uint32_t*>(s_bytes + (a0 & ~3u));
        const uint32_t sh = (a0 & 3) * 8;
        uint32_t carry = p[0], diff = 0;
        uint32_t rw[RW];
#pragma unroll
        for (int v = 0; v < RW; v++) {
            uint32_t word = 0;
            if (v < NW && 16u * v < rlen) {
                uint32_t x[5];
                x[0] = carry;
#pragma unroll
                for (int q = 1; q < 5; q++) x[q] = p[4 * v + q];
                carry = x[4];
                const uint32_t rem = rlen - 16u * v;                    // bases left, >= 1
                if (rem >= 16) {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t c = __funnelshift_r(x[q], x[q + 1], sh);
                        const uint32_t c2 = (c >> 1) & 0x03030303u;
                        const uint32_t expect = 0x41414141u + 2 * c2 + 15 * ((c2 >> 1) & ~c2 & 0x01010101u);
                        diff |= expect ^ c;
                        word |= ((c2 * 0x01041040u) >> 24) << (8 * q);
                    }
                } else {                                                // the read ends inside this step: mask the bytes beyond it
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t c = __funnelshift_r(x[q], x[q + 1], sh);
                        const uint32_t vm = rem >= 4u * q + 4 ? 0xFFFFFFFFu : rem <= 4u * q ? 0u : (0xFFFFFFFFu >> (8 * (4u * q + 4 - rem)));
                        const uint32_t c2 = ((c & vm) >> 1) & 0x03030303u;
                        const uint32_t expect = 0x41414141u + 2 * c2 + 15 * ((c2 >> 1) & ~c2 & 0x01010101u);
                        diff |= (expect ^ c) & vm;
                        word |= ((c2 * 0x01041040u) >> 24) << (8 * q);
                    }
                }
            }
            rw[v] = word;
        }
        if (diff) {
            // rare: some byte is not ACGT -- 'N' (pair skipped, PE_Inference.py:160) or anything else (exhaustive tier), word by word
            bool hasN = false, badc = false;
            uint32_t cy = p[0];
            for (uint32_t v4 = 0; 4 * v4 < rlen; v4++) {
                const uint32_t nx = p[v4 + 1];
                const uint32_t c = __funnelshift_r(cy, nx, sh);
                cy = nx;
                const uint32_t left = rlen - 4 * v4;
                const uint32_t vm = left >= 4 ? 0xFFFFFFFFu : (0xFFFFFFFFu >> (8 * (4 - left)));
                const uint32_t c2 = ((c & vm) >> 1) & 0x03030303u;
                const uint32_t expect = 0x41414141u + 2 * c2 + 15 * ((c2 >> 1) & ~c2 & 0x01010101u);
                uint32_t bad = (expect ^ c) & vm;                       // non-zero bytes = invalid characters
                if (bad) {
                    for (int k = 0; k < 4; k++)
                        if ((bad >> (8 * k)) & 0xFF) { if (((c >> (8 * k)) & 0xFF) == 'N') hasN = true; else badc = true; }
                }
            }
            h |= (hasN ? PH_N : 0) | (badc ? PH_BAD : 0);
        }
        h |= rlen;
        uint4* dst = reinterpret_cast<uint4*>(a.rows + slot * RW);
#pragma unroll
        for (int v = 0; v < RW; v += 4) dst[v >> 2] = make_uint4(rw[v], rw[v + 1], rw[v + 2], rw[v + 3]);
        a.hdr[slot] = h;
        a.seq_start[slot] = (uint64_t)((int64_t)st + pos0);            // chunk-relative start
        a.seq_end[slot] = en == 0xFFFFFFFFu ? ~0ull : (uint64_t)((int64_t)en + pos0);
    }
}

// Both kernels over a device-resident chunk: k_scan_rows, then k_walk over up to n_slots reads (the walk
// kernel reads the chunk's terminator count on the device, so no host synchronisation in between).
int scan_map(Ctx* c, int m, const uint8_t* d_buf, uint64_t n, uint64_t line_base, uint64_t rec_first, uint64_t n_slots, uint32_t* d_handles,
             uint64_t* d_seq_start, uint64_t* d_seq_end, uint32_t* d_rows, uint32_t* d_hdr, uint32_t* d_defer_list,
             unsigned long long* d_defer_count, uint32_t row_words, uint32_t cap) {
    if (n == 0) return VSPE_OK;
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(d_buf) & 15);
    const uint64_t n_tiles = (n + head + SM_TILE - 1) / SM_TILE;
    if (n_tiles > 0x7FFFFFFFull) { set_error("buffer too large for one scan launch"); return VSPE_ERR_LIMIT; }
    if (n_slots > 0xFFFFFFF0ull) { set_error("more than 2^32 reads in one chunk"); return VSPE_ERR_LIMIT; }
    VSPE_TRY(c->tile_base_m[m].reserve(n_tiles + 4));
    unsigned long long* status = reinterpret_cast<unsigned long long*>(c->tile_base_m[m].p);
    VSPE_CUDA(cudaMemsetAsync(status, 0, (n_tiles + 4) * 8, c->stream));
    ScanMapArgs a;
    a.buf = d_buf; a.n = n; a.head = head; a.n_tiles = (uint32_t)n_tiles; a.status = status;
    a.ticket = reinterpret_cast<unsigned int*>(status + n_tiles + 1);
    a.total_out = status + n_tiles + 2;
    a.line_base = line_base; a.rec_first = rec_first; a.n_slots = n_slots;
    a.seq_start = d_seq_start; a.seq_end = d_seq_end; a.rows = d_rows; a.hdr = d_hdr;
    a.row_words = row_words; a.cap = cap; a.counters = c->counters.p;
    if (!c->scan_map_attr_set) {
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_rows<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM));
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_rows<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM));
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_rows<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM));
        for (auto& evs : c->ev_scan) for (auto& e : evs) if (!e) VSPE_CUDA(cudaEventCreate(&e));
        c->scan_map_attr_set = true;
    }
    // both kernels are timed on their own (CUDA events on the launching stream; two launches may be in flight)
    const int k = c->scan_map_events & 1;
    auto launch_scan = [&]() {
        if (row_words == 12) k_scan_rows<12><<<(uint32_t)n_tiles, SM_THREADS, SM_SMEM, c->stream>>>(a);
        else if (row_words == 16) k_scan_rows<16><<<(uint32_t)n_tiles, SM_THREADS, SM_SMEM, c->stream>>>(a);
        else k_scan_rows<20><<<(uint32_t)n_tiles, SM_THREADS, SM_SMEM, c->stream>>>(a);
    };
    if (c->opt_dbg_scan_twice) {
        // measurement aid: the timed launch below then finds every predecessor's inclusive word already published
        // (same results), i.e. it shows the kernel without the look-back wait
        launch_scan();
        VSPE_LAUNCH_CHECK(c);
        VSPE_CUDA(cudaMemsetAsync(a.ticket, 0, 4, c->stream));
    }
    VSPE_CUDA(cudaEventRecord(c->ev_scan[0][2 * k], c->stream));
    launch_scan();
    VSPE_LAUNCH_CHECK(c);
    VSPE_CUDA(cudaEventRecord(c->ev_scan[0][2 * k + 1], c->stream));
    WalkArgs w;
    w.rows = d_rows; w.hdr = d_hdr; w.row_words = row_words; w.n_slots = n_slots; w.total = a.total_out;
    w.line_base = line_base; w.rec_first = rec_first; w.handles = d_handles; w.defer_list = d_defer_list; w.defer_count = d_defer_count; w.counters = c->counters.p;
    const IndexView ix = c->index.view();
    const LinkView lv = link_view(c);
    const uint32_t grid = (uint32_t)((n_slots + WK_THREADS - 1) / WK_THREADS);
    VSPE_CUDA(cudaEventRecord(c->ev_scan[1][2 * k], c->stream));
    if (cap <= 160) k_walk<13><<<grid, WK_THREADS, 0, c->stream>>>(w, ix, lv);
    else if (cap <= 256) k_walk<19><<<grid, WK_THREADS, 0, c->stream>>>(w, ix, lv);
    else k_walk<23><<<grid, WK_THREADS, 0, c->stream>>>(w, ix, lv);
    VSPE_LAUNCH_CHECK(c);
    VSPE_CUDA(cudaEventRecord(c->ev_scan[1][2 * k + 1], c->stream));
    c->scan_map_pending[k] = true;
    c->scan_map_events++;
    return VSPE_OK;
}

// terminators of the chunk the last scan_map launch covered (device word, read after a stream sync)
const unsigned long long* scan_map_total_ptr(Ctx* c, int m, uint64_t n, const uint8_t* d_buf) {
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(d_buf) & 15);
    const uint64_t n_tiles = (n + head + SM_TILE - 1) / SM_TILE;
    return reinterpret_cast<unsigned long long*>(c->tile_base_m[m].p) + n_tiles + 2;
}

// fold the durations of the finished k_scan_rows / k_walk launches into the stats (call after a stream sync)
void scan_map_account(Ctx* c) {
    for (int k = 0; k < 2; k++) {
        if (!c->scan_map_pending[k]) continue;
        float ms = 0;
        if (cudaEventQuery(c->ev_scan[1][2 * k + 1]) != cudaSuccess) continue;
        if (cudaEventElapsedTime(&ms, c->ev_scan[0][2 * k], c->ev_scan[0][2 * k + 1]) == cudaSuccess) { c->stats.ms_k_scan_rows += ms; c->stats.n_k_scan_rows++; }
        if (cudaEventElapsedTime(&ms, c->ev_scan[1][2 * k], c->ev_scan[1][2 * k + 1]) == cudaSuccess) { c->stats.ms_k_walk += ms; c->stats.n_k_walk++; }
        c->scan_map_pending[k] = false;
    }
    cudaGetLastError();
}

}  // namespace vspe
