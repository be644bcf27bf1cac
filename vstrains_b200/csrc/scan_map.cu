// scan_map.cu -- K1 + K2 + K4 in ONE pass over the raw FASTQ bytes: every input byte is read from
// HBM once and a read costs 4 bytes on the way out (the handle of its node list, link.cuh).
//
// Replaces `readlines()` + `[s[:-1] ...]` (reference utils/VStrains_PE_Inference.py:149-159), the
// per-character work of `fseq.count("N")` / k-mer slicing (:160, :25) and single_end_read_mapping
// (:16-48) for every read whose result the walk below can PROVE; the rest (about 3 % on the bench
// workloads: two or more sequencing errors, repeats, non-ACGT characters, very long reads) is handed,
// packed, to the list-driven tiers of map_fast.cu / map_generic.cu.
//
// Per 40 KiB tile (one CTA of 10 warps, 3 CTAs per SM, tiles handed out by a ticket counter):
//   1. one elected thread issues TMA bulk copies (cp.async.bulk, mbarrier complete_tx) of the tile
//      + a 16-byte front margin + a 512-byte back margin into shared memory;
//   2. every lane tests its 16-byte vectors for bytes < 0x10 or >= 0x80 (two instructions per
//      32-bit word); candidate vectors go to a per-warp queue (warp ballots) and only they get exact
//      terminator masks (universal newlines: '\n', "\r\n" once, lone '\r'); a warp scan ranks them;
//   3. warp totals + one block exchange give the tile's terminator count; a decoupled look-back
//      over the tiles' status words gives the line number of the tile's first line;
//   4. terminators with line%4==0 start a sequence line, line%4==1 end it -> per-tile read table;
//   5. 8 / 16 lanes per read pack each read the tile owns (its sequence line STARTS here) to
//      2 bits/base straight from the tile into a shared-memory row (SIMD-in-word ACGT validity test,
//      'N' flag);
//   6. one thread per read walks its row through the index: seed window 0 (hash + probe + verify +
//      uniq bit; on a miss the reverse complement is seeded from the other end), then a flat loop
//      whose every turn compares 32 bases + 32 uniq bits on the current diagonal and, when the
//      stretch is complete, books it and steps to the successor strand.  ONE mismatching base is
//      tolerated when the substitution-hit bit proves that the windows covering it have no posting.
//      The saturation predicate (:36-47, integer form) is applied as each stretch is booked;
//   7. the kept node list is interned (link.cuh) and its handle stored; an unresolved read is
//      stored packed for the next tier and listed.
// Why this is exact: see map_fast.cu (a window is counted without a table access only if its text
// equality and the uniq bit of that text window were both checked; it is skipped only if the
// index build already looked that k-mer up and found nothing).
#include <algorithm>

#include "link.cuh"
#include "map_common.cuh"

namespace vspe {

static constexpr int SM_WARPS = 10;
static constexpr int SM_THREADS = SM_WARPS * 32;
static constexpr int SM_ITERS = 8;                               // 512-byte warp rows per warp
static constexpr int SM_TILE = SM_WARPS * SM_ITERS * 32 * 16;    // 40 KiB
static constexpr int SM_FRONT = 16;                              // bytes kept before the tile
static constexpr int SM_BACK = 512;                              // bytes kept after the tile (>= longest packed read + 1)
static constexpr int SM_MAXREC = 576;                            // reads a tile may own (else the chunk takes the plain path)
static constexpr int SM_QCAP = 96;                               // per warp: vectors that may hold a terminator
static constexpr int SM_RPR = 160;                               // reads packed + walked per round
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
    uint32_t cr = movemask4b(__vcmpeq4(v.x, 0x0D0D0D0Du)) | (movemask4b(__vcmpeq4(v.y, 0x0D0D0D0Du)) << 4) |
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
    uint32_t* handles;               // [n_slots] node-list handle / H_N / H_SHORT / H_PENDING per read
    // written for unresolved reads only (indexed like handles):
    uint64_t* seq_start;             // chunk-relative byte range of the sequence line
    uint64_t* seq_end;
    uint32_t* rows;                  // [n_slots][row_words] packed read
    uint32_t* hdr;                   // rlen | flags << 24
    uint32_t* defer_list;            // their indices, counters[CNT_DEFER] of them
    uint32_t row_words;              // 12, 16 or 20
    uint32_t cap;                    // longest read (bases) a packed row holds
    unsigned long long* counters;
};

// The walk of step 6.  row: this thread's packed read (STRIDE words, zero padded); lst: its node list
// column (entry i at lst[i * SM_RPR]).  Returns true when every window of the read is accounted for;
// n_kept nodes that pass the saturation predicate are then at the front of the column.  On false
// the row is unchanged (a reverse complement taken for seeding is undone).
template <int STRIDE>
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
    uint32_t i0 = 0, p = L, q = 0, lim = 0, nn = 0, kept = 0;
    unsigned long long seen = 0;                               // 64-bit filter over the nodes booked so far
    int delta = 0;
    // strand of window i0 = text position tp; false if the error's windows cannot be proven to miss there
    auto enter = [&]() -> bool {
        const uint32_t s1 = __ldg(ix.strand_start + 2 * node + 1);
        const bool rcs = tp >= s1;
        q = 2 * node + (rcs ? 1u : 0u);
        const uint32_t send = rcs ? __ldg(ix.strand_start + 2 * node + 2) : s1;
        delta = (int)tp - (int)i0;
        lim = min(rlen, (uint32_t)((int)send - delta));        // read position where the strand ends
        // a strand entered after the error still holds windows covering it if it starts at or before e
        if (err && (int)i0 <= e) {
            const uint32_t te = (uint32_t)(e + delta);
            if ((__ldg(ix.subst + (te >> 3)) >> (4 * (te & 7) + rb)) & 1u) return false;
        }
        return true;
    };
    if (running && !enter()) running = false;
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
                // a node met twice (cyclic graph) needs its hits merged: next tier
                const uint32_t hb = (node * 0x9E3779B1u) >> 26;
                bool dup = false;
                if ((seen >> hb) & 1ull)
                    for (uint32_t i = 0; i < nn; i++) dup |= lst[i * SM_RPR] == node;
                if (dup) { running = false; break; }
                seen |= 1ull << hb;
                const uint32_t v = (uint32_t)(c1 + c2), kmin = (uint32_t)(mirror ? npos - 1 - last_hit : first_hit);
                if (keep_node_f(v, kmin, __ldg(ix.node_len + node), rlen, L)) kept |= 1u << nn;
                lst[nn * SM_RPR] = node;
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
    if (!resolved) {
        if (mirror) revcomp_row<NW>(row, rlen);
        return false;
    }
    uint32_t k = 0;
    for (uint32_t i = 0; i < nn; i++)
        if ((kept >> i) & 1u) { lst[k * SM_RPR] = lst[i * SM_RPR]; k++; }
    n_kept = k;
    return true;
}

template <int STRIDE>
struct ScanMapSmem {
    static constexpr uint32_t QUEUE_BYTES = SM_WARPS * SM_QCAP * 8;
    static constexpr uint32_t ROUND_BYTES = SM_RPR * (STRIDE + SM_MAXST + 1) * 4;
    static constexpr uint32_t UNION_BYTES = QUEUE_BYTES > ROUND_BYTES ? QUEUE_BYTES : ROUND_BYTES;
    static constexpr uint32_t TOTAL = SM_FRONT + SM_TILE + SM_BACK + SM_MAXREC * 8 + UNION_BYTES + 64;
};

template <int STRIDE>
__global__ void __launch_bounds__(SM_THREADS, 3)
k_scan_map(const ScanMapArgs a, const IndexView ix, const LinkView lv) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_bytes = smem;                                           // [SM_FRONT + SM_TILE + SM_BACK]
    uint32_t* s_rs = reinterpret_cast<uint32_t*>(smem + SM_FRONT + SM_TILE + SM_BACK);   // read start (tile-relative)
    uint32_t* s_re = s_rs + SM_MAXREC;                                 // read end
    uint32_t* s_un = s_re + SM_MAXREC;                                 // queues, later the rows of a round
    uint32_t* s_qmk = s_un;                                            // [SM_WARPS][SM_QCAP] term | crlf << 16
    uint16_t* s_qid = reinterpret_cast<uint16_t*>(s_qmk + SM_WARPS * SM_QCAP);   // vector index in the tile
    uint16_t* s_qrk = s_qid + SM_WARPS * SM_QCAP;                      // rank of the vector's first terminator in the warp
    uint32_t* s_rows = s_un;                                           // [SM_RPR][STRIDE]
    uint32_t* s_lst = s_rows + SM_RPR * STRIDE;                        // [SM_MAXST][SM_RPR]
    uint32_t* s_rhdr = s_lst + SM_MAXST * SM_RPR;                      // [SM_RPR] rlen | flags
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
    for (uint32_t i = threadIdx.x; i < n_local; i += blockDim.x) s_re[i] = 0xFFFFFFFFu;
    if (chunk_starts_in_seq && threadIdx.x == 0) s_rs[0] = (uint32_t)a.head;   // buffer position 0, tile-relative
    __syncthreads();
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
    __syncthreads();                                                   // the queues are dead from here on: s_un holds rows
    if (threadIdx.x == 0 && n_local) atomicAdd(&a.counters[CNT_FAST], (unsigned long long)n_local);
    const uint64_t r_loc0 = r_own0 - shift;                            // record number of local index 0
    const uint32_t L = ix.split_len;
    constexpr uint32_t LPRP = STRIDE > 19 ? 16u : 8u;                  // lanes per read in the pack step
    constexpr uint32_t gpw = 32 / LPRP;                                // reads per warp step
    const uint32_t grp = lane / LPRP, gl = lane % LPRP;
    const uint32_t gmask = (LPRP == 16 ? 0xFFFFu : 0xFFu) << (grp * LPRP);
    for (uint32_t base_li = 0; base_li < n_local; base_li += SM_RPR) {
        const uint32_t n_round = min((uint32_t)SM_RPR, n_local - base_li);
        // ---- pack: LPRP lanes per read, 32 bases (two 32-bit words) per lane -> s_rows --------------
        for (uint32_t j0 = wib * gpw; j0 < n_round; j0 += SM_WARPS * gpw) {
            const uint32_t jr = j0 + grp;                               // read of the round
            const uint32_t li = base_li + jr;
            const bool live = jr < n_round;
            const uint32_t st = live ? s_rs[li] : 0;
            uint32_t en = live ? s_re[li] : 0;
            uint32_t flags = 0;
            if (__any_sync(0xFFFFFFFFu, live && en == 0xFFFFFFFFu)) {
                // some line ends beyond the tile: first '\n' or '\r' in the back margin, if any
                for (uint32_t g = 0; g < gpw; g++) {
                    const uint32_t en_g = __shfl_sync(0xFFFFFFFFu, en, g * LPRP);
                    const bool live_g = __shfl_sync(0xFFFFFFFFu, (uint32_t)live, g * LPRP) != 0;
                    if (!live_g || en_g != 0xFFFFFFFFu) continue;
                    uint32_t found = 0xFFFFFFFFu;
                    for (uint32_t k = 0; k < (uint32_t)SM_BACK && found == 0xFFFFFFFFu; k += 32) {
                        const uint32_t j = SM_TILE + k + lane;
                        const uint32_t ch = tb[j];
                        const bool hit = (pos0 + j < nn) && (ch == '\n' || ch == '\r');
                        const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
                        if (m) found = SM_TILE + k + (uint32_t)(__ffs((int)m) - 1);
                    }
                    if (grp == g) en = found;
                }
                if (live && en == 0xFFFFFFFFu) flags |= PH_LONG;
                if (live && gl == 0) s_re[li] = en;                     // the walk step reports the resolved end
            }
            uint32_t rlen = (!live || (flags & PH_LONG)) ? 0u : en - st;
            if (rlen > a.cap) { flags |= PH_LONG; rlen = 0; }
            bool hasN = false, badc = false;
            const uint32_t b0 = 32 * gl;
            uint32_t w0 = 0, w1 = 0;
            if (b0 < rlen) {
                const uint32_t nb = min(32u, rlen - b0);
                const uint32_t jb = SM_FRONT + st + b0;                     // offset in s_bytes (16-byte aligned base)
                const uint32_t* p = reinterpret_cast<const uint32_t*>(s_bytes + (jb & ~3u));
                const uint32_t sh = (jb & 3) * 8;
                uint32_t x[9];
#pragma unroll
                for (int q = 0; q < 9; q++) x[q] = p[q];
                uint32_t diff = 0;
                uint32_t pk[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const uint32_t c = __funnelshift_r(x[q], x[q + 1], sh);
                    // mask of the valid bytes of this word: 4 + 4q - nb of its top bytes lie beyond the lane's bases
                    const uint32_t vm = __funnelshift_rc(0xFFFFFFFFu, 0u, 8u * (uint32_t)max(4 + 4 * q - (int)nb, 0));
                    const uint32_t c2 = ((c & vm) >> 1) & 0x03030303u;
                    // the only byte with code k is "ACTG"[k] = 0x41 + 2k (+15 when k == 2)
                    const uint32_t expect = 0x41414141u + 2 * c2 + 15 * ((c2 >> 1) & ~c2 & 0x01010101u);
                    diff |= (expect ^ c) & vm;
                    pk[q] = (c2 * 0x01041040u) >> 24;
                }
                w0 = pk[0] | (pk[1] << 8) | (pk[2] << 16) | (pk[3] << 24);
                w1 = pk[4] | (pk[5] << 8) | (pk[6] << 16) | (pk[7] << 24);
                if (diff) {                                                 // rare: some byte is not ACGT
                    for (uint32_t q = 0; q < nb; q++) {
                        const uint32_t c = s_bytes[jb + q];
                        if (c == 'N') hasN = true;
                        else if (!is_acgt(c)) badc = true;
                    }
                }
            }
            if (live) {
                uint32_t* row = s_rows + jr * STRIDE;
                if (2 * gl < (uint32_t)STRIDE) row[2 * gl] = w0;
                if (2 * gl + 1 < (uint32_t)STRIDE) row[2 * gl + 1] = w1;
                if (2 * LPRP < (uint32_t)STRIDE && 2 * LPRP + gl < (uint32_t)STRIDE) row[2 * LPRP + gl] = 0;
            }
            const uint32_t bN = __ballot_sync(0xFFFFFFFFu, hasN) & gmask, bB = __ballot_sync(0xFFFFFFFFu, badc) & gmask;
            if (live && gl == 0) s_rhdr[jr] = rlen | flags | (bN ? PH_N : 0) | (bB ? PH_BAD : 0);
        }
        __syncthreads();
        // ---- walk: one thread per read of the round ---------------------------------------------------
        if (threadIdx.x < n_round) {
            const uint32_t jr = threadIdx.x, li = base_li + jr;
            const uint64_t slot = r_loc0 + li - a.rec_first;
            if (slot < a.n_slots) {
                const uint32_t h = s_rhdr[jr];
                const uint32_t rlen = h & 0xFFFFFF;
                uint32_t* row = s_rows + jr * STRIDE;
                uint32_t handle = H_PENDING;
                bool defer = (h & (PH_LONG | PH_BAD)) != 0;
                if (!(h & PH_LONG)) {
                    if (h & PH_N) handle = H_N;                            // 'N' before the length (PE_Inference.py:160-163)
                    else if (rlen < L) handle = H_SHORT;
                }
                if (handle == H_PENDING && !defer) {
                    uint32_t n_kept = 0;
                    if (walk_read<STRIDE>(ix, row, rlen, s_lst + jr, n_kept)) handle = intern_list(lv, n_kept, s_lst + jr, SM_RPR);
                    else defer = true;
                }
                if (handle == H_PENDING && defer) {
                    // unresolved: the packed row, its header and the byte range go to the list-driven tiers
                    uint4* dst = reinterpret_cast<uint4*>(a.rows + slot * a.row_words);
                    for (uint32_t w = 0; w < a.row_words; w += 4) {
                        uint4 v;
                        v.x = w < (uint32_t)STRIDE ? row[w] : 0u;
                        v.y = w + 1 < (uint32_t)STRIDE ? row[w + 1] : 0u;
                        v.z = w + 2 < (uint32_t)STRIDE ? row[w + 2] : 0u;
                        v.w = w + 3 < (uint32_t)STRIDE ? row[w + 3] : 0u;
                        dst[w >> 2] = v;
                    }
                    a.hdr[slot] = h;
                    const uint32_t st = s_rs[li], en = s_re[li];
                    a.seq_start[slot] = (uint64_t)((int64_t)st + pos0);    // chunk-relative start
                    a.seq_end[slot] = en == 0xFFFFFFFFu ? ~0ull : (uint64_t)((int64_t)en + pos0);
                    a.defer_list[atomicAdd(&a.counters[CNT_DEFER], 1ull)] = (uint32_t)slot;
                }
                a.handles[slot] = handle;
            } else if (li < n_local) {
                atomicOr(&a.counters[CNT_ERR], (unsigned long long)ERRF_SLOTS_FULL);
            }
        }
        __syncthreads();                                               // the rows are reused by the next round
    }
}

// One launch over a device-resident chunk.  Returns the terminator count and the kernels' error flags
// (transient *_FULL scan flags are cleared on the device; the caller repeats or falls back).
int scan_map(Ctx* c, const uint8_t* d_buf, uint64_t n, uint64_t line_base, uint64_t rec_first, uint64_t n_slots, uint32_t* d_handles,
             uint64_t* d_seq_start, uint64_t* d_seq_end, uint32_t* d_rows, uint32_t* d_hdr, uint32_t* d_defer_list, uint32_t row_words,
             uint32_t cap) {
    if (n == 0) return VSPE_OK;
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(d_buf) & 15);
    const uint64_t n_tiles = (n + head + SM_TILE - 1) / SM_TILE;
    if (n_tiles > 0x7FFFFFFFull) { set_error("buffer too large for one scan launch"); return VSPE_ERR_LIMIT; }
    if (n_slots > 0xFFFFFFF0ull) { set_error("more than 2^32 reads in one chunk"); return VSPE_ERR_LIMIT; }
    VSPE_TRY(c->tile_base.reserve(n_tiles + 4));
    unsigned long long* status = reinterpret_cast<unsigned long long*>(c->tile_base.p);
    VSPE_CUDA(cudaMemsetAsync(status, 0, (n_tiles + 4) * 8, c->stream));
    ScanMapArgs a;
    a.buf = d_buf; a.n = n; a.head = head; a.n_tiles = (uint32_t)n_tiles; a.status = status;
    a.ticket = reinterpret_cast<unsigned int*>(status + n_tiles + 1);
    a.total_out = status + n_tiles + 2;
    a.line_base = line_base; a.rec_first = rec_first; a.n_slots = n_slots; a.handles = d_handles;
    a.seq_start = d_seq_start; a.seq_end = d_seq_end; a.rows = d_rows; a.hdr = d_hdr; a.defer_list = d_defer_list;
    a.row_words = row_words; a.cap = cap; a.counters = c->counters.p;
    if (!c->scan_map_attr_set) {
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_map<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, ScanMapSmem<13>::TOTAL));
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_map<19>, cudaFuncAttributeMaxDynamicSharedMemorySize, ScanMapSmem<19>::TOTAL));
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_map<23>, cudaFuncAttributeMaxDynamicSharedMemorySize, ScanMapSmem<23>::TOTAL));
        for (auto& e : c->ev_scan[0]) if (!e) VSPE_CUDA(cudaEventCreate(&e));
        c->scan_map_attr_set = true;
    }
    const IndexView ix = c->index.view();
    const LinkView lv = link_view(c);
    // the dominant kernel is timed on its own (CUDA events on the launching stream)
    cudaEvent_t e0 = c->ev_scan[0][c->scan_map_events & 1 ? 2 : 0], e1 = c->ev_scan[0][c->scan_map_events & 1 ? 3 : 1];
    VSPE_CUDA(cudaEventRecord(e0, c->stream));
    if (cap <= 160) k_scan_map<13><<<(uint32_t)n_tiles, SM_THREADS, ScanMapSmem<13>::TOTAL, c->stream>>>(a, ix, lv);
    else if (cap <= 256) k_scan_map<19><<<(uint32_t)n_tiles, SM_THREADS, ScanMapSmem<19>::TOTAL, c->stream>>>(a, ix, lv);
    else k_scan_map<23><<<(uint32_t)n_tiles, SM_THREADS, ScanMapSmem<23>::TOTAL, c->stream>>>(a, ix, lv);
    VSPE_LAUNCH_CHECK(c);
    VSPE_CUDA(cudaEventRecord(e1, c->stream));
    c->scan_map_pending[c->scan_map_events & 1] = true;
    c->scan_map_events++;
    return VSPE_OK;
}

// terminators of the chunk the last scan_map launch covered (device word, read after a stream sync)
const unsigned long long* scan_map_total_ptr(Ctx* c, uint64_t n, const uint8_t* d_buf) {
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(d_buf) & 15);
    const uint64_t n_tiles = (n + head + SM_TILE - 1) / SM_TILE;
    return reinterpret_cast<unsigned long long*>(c->tile_base.p) + n_tiles + 2;
}

// fold the durations of the finished k_scan_map launches into the stats (call after a stream sync)
void scan_map_account(Ctx* c) {
    for (int k = 0; k < 2; k++) {
        if (!c->scan_map_pending[k]) continue;
        cudaEvent_t e0 = c->ev_scan[0][k ? 2 : 0], e1 = c->ev_scan[0][k ? 3 : 1];
        float ms = 0;
        if (cudaEventQuery(e1) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) {
            c->stats.ms_k_scan_pack += ms;
            c->stats.n_k_scan_pack++;
            c->scan_map_pending[k] = false;
        }
    }
    cudaGetLastError();
}

}  // namespace vspe
