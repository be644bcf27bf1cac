// scan_pack.cu -- K1 + K2: the raw FASTQ bytes -> record table -> 2-bit rows.
//
// Replaces `readlines()` + `[s[:-1] ...]` (reference utils/VStrains_PE_Inference.py:149-159) and
// the per-character work of `fseq.count("N")` / k-mer slicing (:160, :25).
//
// Per 48 KiB tile (one CTA of 10 warps, 3 CTAs per SM):
//   1. one elected thread issues TMA bulk copies (cp.async.bulk, mbarrier complete_tx) of the
//      tile + a 16-byte front margin + a 512-byte back margin into shared memory;
//   2. every lane tests its 16-byte vectors for bytes < 0x10; candidate vectors go to a per-warp
//      queue (warp ballots) and only they get exact terminator masks (universal newlines: '\n',
//      "\r\n" once, lone '\r'); a warp scan ranks the terminators;
//   3. warp totals + one block exchange give the tile's terminator count;
//   4. the line number of the tile's first line: scanned tile counts (two-kernel mode) or a
//      decoupled look-back (fused mode);
//   5. terminators with line%4==0 start a sequence line, line%4==1 end it -> per-tile read
//      table in shared memory (+ seq_start/seq_end in HBM for the exhaustive tier);
//   6. half-warps pack each read the tile owns (its sequence line STARTS here) to 2 bits/base
//      straight from the tile: 64-byte rows in HBM + one header word (length | flags).
// Default (scan_mode 0): steps 1-3 as a count pass that also saves the queues (no tile waits on
// another), a device scan of the tile counts, then steps 1, 4-6 as a pack pass with exactly
// sized outputs.  scan_mode 3 runs 1-6 in one kernel with the look-back.  Both are limited by
// instruction issue, not by HBM.  The map kernels never touch the raw bytes.
#include "ctx.cuh"

namespace vspe {

static constexpr int SP_WARPS = 10;
static constexpr int SP_ITERS = 8;                               // 512-byte warp rows per warp
static constexpr int SP_TILE = SP_WARPS * SP_ITERS * 32 * 16;    // 48 KiB
static constexpr int SP_FRONT = 16;                              // bytes kept before the tile
static constexpr int SP_BACK = 512;                              // bytes kept after the tile (>= longest packed read + 1)
static constexpr int SP_MAXREC = 576;                            // reads a tile may own (else fallback path)
static constexpr int SP_QCAP = 96;                               // per warp: vectors that may hold a terminator
static constexpr uint32_t SP_QBYTES = (4 + SP_QCAP * 6 + 15) / 16 * 16;   // header word + masks (u32) + vector ids (u16)
static constexpr uint32_t SP_SMEM = SP_FRONT + SP_TILE + SP_BACK + SP_MAXREC * 8 + SP_WARPS * SP_QCAP * 8 + 64;

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

struct ScanPackArgs {
    const uint8_t* buf;              // chunk start (may be misaligned)
    uint64_t n;                      // chunk bytes
    uint32_t head;                   // address of buf mod 16
    unsigned long long* status;      // look-back words, one per tile (zeroed)
    unsigned int* ticket;
    unsigned long long* total_out;   // terminators in the chunk
    uint64_t line_base;              // lines before this chunk
    uint64_t rec_first;              // record number of slot 0 of the outputs
    uint64_t n_slots;                // capacity of the per-read outputs
    uint64_t* seq_start;             // [n_slots] chunk-relative
    uint64_t* seq_end;
    uint32_t* rows;                  // [n_slots][row_words] packed reads
    uint32_t* hdr;                   // [n_slots] rlen | flags << 24
    uint32_t row_words;              // 12, 16 or 20
    uint32_t cap;                    // longest read (bases) the map kernel's packed rows hold
    uint32_t n_tiles;
    unsigned long long* counters;
    unsigned long long* dbg;         // optional [n_tiles][8] globaltimer stamps (profiling aid)
    uint8_t* qbuf;                   // two-kernel mode: per (tile, warp) candidate queue, SP_QBYTES each
    unsigned long long* tile_tot;    // two-kernel mode: terminators per tile (count pass) ...
    const unsigned long long* tile_excl;   // ... and their exclusive prefix (pack pass)
    unsigned long long* tile_full;   // two-kernel mode: set by the count pass when a warp queue overflows
};


__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define SP_STAMP(k) do { if (a.dbg && threadIdx.x == 0) a.dbg[(size_t)tile * 8 + (k)] = gtime(); } while (0)

// MODE 0: fused single pass with decoupled look-back.
// MODE 1: count pass of the two-kernel variant -- TMA + terminator masks only; saves each warp's
//         candidate queue and the tile's terminator count, so no tile ever waits on another.
// MODE 2: pack pass -- TMA + saved masks + scanned tile prefix -> read table -> 2-bit rows.
template <int MODE>
__global__ void __launch_bounds__(SP_WARPS * 32)
k_scan_pack(ScanPackArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_bytes = smem;                                           // [SP_FRONT + SP_TILE + SP_BACK]
    uint32_t* s_rs = reinterpret_cast<uint32_t*>(smem + SP_FRONT + SP_TILE + SP_BACK);   // read start (tile-relative + SP_FRONT)
    uint32_t* s_re = s_rs + SP_MAXREC;                                 // read end
    uint32_t* s_qmk = s_re + SP_MAXREC;                                // [SP_WARPS][SP_QCAP] term | crlf << 16
    uint16_t* s_qid = reinterpret_cast<uint16_t*>(s_qmk + SP_WARPS * SP_QCAP);   // vector index in the tile
    uint16_t* s_qrk = s_qid + SP_WARPS * SP_QCAP;                      // rank of the vector's first terminator in the warp
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_wtot[SP_WARPS];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_excl;
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        s_tile = MODE == 0 ? atomicAdd(a.ticket, 1u) : blockIdx.x;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    SP_STAMP(0);
    // aligned coordinates: byte `off` of the aligned stream is buffer position off - head
    const uint64_t A = ((uint64_t)a.head + a.n + 15) & ~15ull;         // aligned stream length
    const uint64_t t_lo = (uint64_t)tile * SP_TILE;
    const uint64_t ld_lo = t_lo >= SP_FRONT ? t_lo - SP_FRONT : 0;
    const uint64_t ld_hi = min(A, t_lo + SP_TILE + SP_BACK);
    const uint32_t s_off0 = tile == 0 ? SP_FRONT : 0;                  // where ld_lo lands in s_bytes
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
    SP_STAMP(1);
    // tile byte j (0 <= j < SP_TILE) lives at s_bytes[SP_FRONT + j]; its buffer position is t_lo + j - head
    const uint8_t* tb = s_bytes + SP_FRONT;
    const int64_t pos0 = (int64_t)t_lo - a.head;                       // buffer position of tile byte 0
    const int64_t nn = (int64_t)a.n;
    auto byte_at = [&](int64_t j) -> uint32_t {                         // tile-relative byte, 0 outside the buffer
        const int64_t p = pos0 + j;
        return (p >= 0 && p < nn) ? tb[j] : 0u;
    };

    uint16_t* q_id = s_qid + wib * SP_QCAP;
    uint32_t* q_mk = s_qmk + wib * SP_QCAP;
    uint16_t* q_rk = s_qrk + wib * SP_QCAP;
    const bool interior = pos0 >= 1 && pos0 + SP_TILE + 16 <= nn;   // CTA-uniform
    const uint32_t lt = (1u << lane) - 1;
    uint32_t qn = 0, wcount = 0;
    bool bad = false, q_over = false;
    uint8_t* gq = a.qbuf ? a.qbuf + ((uint64_t)tile * SP_WARPS + wib) * SP_QBYTES : nullptr;
    if (MODE == 2) {
        // the count pass already found the candidates and their exact masks
        const uint32_t hdrw = *reinterpret_cast<const uint32_t*>(gq);
        qn = hdrw & 0xFFFF;
        const uint32_t* gmk = reinterpret_cast<const uint32_t*>(gq + 4);
        const uint16_t* gid = reinterpret_cast<const uint16_t*>(gq + 4 + 4 * SP_QCAP);
        for (uint32_t i0 = 0; i0 < qn; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint32_t mk = 0;
            if (i < qn) { mk = gmk[i]; q_mk[i] = mk; q_id[i] = gid[i]; }
            const uint32_t c = __popc(mk & 0xFFFFu);
            uint32_t inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= (uint32_t)d) inc += y;
            }
            if (i < qn) q_rk[i] = (uint16_t)(wcount + inc - c);
            wcount += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        __syncwarp();
    } else {
    // ---- M1: which 16-byte vectors can hold a terminator?  ('\n' and '\r' are < 0x10) -----------
    // Warp w owns tile bytes [w*8K, (w+1)*8K) as 16 coalesced 512-byte rows.  Candidate vectors are
    // appended, in (row, lane) order, to the warp's queue, so everything after this loop runs on a
    // dense list instead of diverging on every row.
    uint32_t ora = 0;
#pragma unroll
    for (int it = 0; it < SP_ITERS; it++) {
        const uint32_t vid = (wib * SP_ITERS + it) * 32 + lane;       // vector index inside the tile
        const uint4 v = *reinterpret_cast<const uint4*>(tb + vid * 16);
        bool cand, full = true;
        if (!interior) {                                               // first / last tile of the chunk
            const int64_t p = pos0 + (int64_t)vid * 16;
            full = p >= 0 && p + 16 <= nn;
            cand = !full && p < nn && p + 16 > 0;                      // partial vector: M2 masks the outside bytes
        }
        if (full) {
            ora |= v.x | v.y | v.z | v.w;
            const uint32_t low = ((v.x - 0x10101010u) & ~v.x) | ((v.y - 0x10101010u) & ~v.y) |
                                 ((v.z - 0x10101010u) & ~v.z) | ((v.w - 0x10101010u) & ~v.w);
            cand = (low & 0x80808080u) != 0;
        }
        const uint32_t bm = __ballot_sync(0xFFFFFFFFu, cand);
        if (cand) {
            const uint32_t at = qn + __popc(bm & lt);
            if (at < SP_QCAP) q_id[at] = (uint16_t)vid;
        }
        qn += __popc(bm);
    }
    bad = (ora & 0x80808080u) != 0;
    q_over = qn > SP_QCAP;
    if (q_over) qn = SP_QCAP;
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
    }
    if (bad) atomicOr(&a.counters[CNT_ERR], (unsigned long long)ERRF_NON_ASCII);
    if (lane == 0) s_wtot[wib] = wcount | (q_over ? 0x80000000u : 0u);
    __syncthreads();
    uint32_t tile_total = 0, warp_base = 0;
    bool any_over = false;
#pragma unroll
    for (int w = 0; w < SP_WARPS; w++) {
        const uint32_t x = s_wtot[w];
        any_over |= (x >> 31) != 0;
        if (w < (int)wib) warp_base += x & 0x7FFFFFFFu;
        tile_total += x & 0x7FFFFFFFu;
    }
    SP_STAMP(2);
    if (MODE == 1) {
        // count pass: save the queue (ids, masks) and the tile total; the pack pass does the rest
        uint32_t* gmk = reinterpret_cast<uint32_t*>(gq + 4);
        uint16_t* gid = reinterpret_cast<uint16_t*>(gq + 4 + 4 * SP_QCAP);
        for (uint32_t i = lane; i < qn; i += 32) { gmk[i] = q_mk[i]; gid[i] = q_id[i]; }
        if (lane == 0) *reinterpret_cast<uint32_t*>(gq) = qn;
        if (threadIdx.x == 0) {
            a.tile_tot[tile] = tile_total;
            if (any_over) atomicOr(a.tile_full, 1ull);
        }
        return;
    }
    // ---- decoupled look-back (warp 0), 128 predecessors per hop -------------------------------
    // The inclusive-prefix frontier can only advance by one window per L2 round trip, so the
    // window width bounds the kernel's throughput: 4 status words per lane.
    if (MODE == 2) {
        if (threadIdx.x == 0) s_excl = a.tile_excl[tile];
    } else if (wib == 0) {
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
    SP_STAMP(3);
    const uint64_t base = a.line_base + s_excl;                        // line number of the tile's first line
    // reads owned by this tile: sequence lines that START here = header terminators (line%4==0)
    // in the tile; record numbers are consecutive from r_own0
    const uint64_t r_own0 = (base + 3) >> 2;
    const uint64_t last_line = base + tile_total;                      // one past the tile's last terminator
    uint32_t n_own = (uint32_t)(((last_line + 3) >> 2) - r_own0);      // #{l in [base, last_line) : l%4 == 0}
    const bool chunk_starts_in_seq = tile == 0 && (a.line_base & 3) == 1;   // chunk begins with a sequence line
    const uint32_t shift = chunk_starts_in_seq ? 1u : 0u;              // that read becomes local index 0
    const uint32_t n_local = n_own + shift;
    const bool too_many = n_local > SP_MAXREC || any_over;
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
    __syncthreads();
    SP_STAMP(4);
    // ---- pack: LPRP lanes per read, 32 bases (two 32-bit words) per lane ---------------------------
    const uint64_t r_loc0 = r_own0 - shift;                            // record number of local index 0
    const uint32_t RW = a.row_words;
    const uint32_t cap = a.cap;
    const uint32_t LPRP = RW > 16 ? 16u : 8u;                          // lanes per read (CTA-uniform)
    const uint32_t gpw = 32 / LPRP;                                    // reads per warp step
    const uint32_t grp = lane / LPRP, gl = lane % LPRP;
    const uint32_t gmask = (LPRP == 16 ? 0xFFFFu : 0xFFu) << (grp * LPRP);
    for (uint32_t li0 = wib * gpw; li0 < n_local; li0 += SP_WARPS * gpw) {
        const uint32_t li = li0 + grp;
        const bool live = li < n_local;
        const uint64_t slot = r_loc0 + li - a.rec_first;
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
                for (uint32_t k = 0; k < SP_BACK && found == 0xFFFFFFFFu; k += 32) {
                    const uint32_t j = SP_TILE + k + lane;
                    const uint32_t ch = tb[j];
                    const bool hit = (pos0 + j < nn) && (ch == '\n' || ch == '\r');
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
                    if (m) found = SP_TILE + k + (uint32_t)(__ffs((int)m) - 1);
                }
                if (grp == g) en = found;
            }
            if (live && en == 0xFFFFFFFFu) flags |= PH_LONG;
        }
        uint32_t rlen = (!live || (flags & PH_LONG)) ? 0u : en - st;
        if (rlen > cap) { flags |= PH_LONG; rlen = 0; }
        const bool fits = live && slot < a.n_slots;
        if (live && !fits && gl == 0) atomicOr(&a.counters[CNT_ERR], (unsigned long long)ERRF_SLOTS_FULL);
        bool hasN = false, badc = false;
        const uint32_t b0 = 32 * gl;
        uint32_t w0 = 0, w1 = 0;
        if (b0 < rlen) {
            const uint32_t nb = min(32u, rlen - b0);
            const uint32_t jb = SP_FRONT + st + b0;                     // offset in s_bytes (16-byte aligned base)
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
        if (fits && 2 * gl < RW) *reinterpret_cast<uint2*>(a.rows + slot * RW + 2 * gl) = make_uint2(w0, w1);
        const uint32_t bN = __ballot_sync(0xFFFFFFFFu, hasN) & gmask, bB = __ballot_sync(0xFFFFFFFFu, badc) & gmask;
        if (fits && gl == 0) {
            a.hdr[slot] = rlen | flags | (bN ? PH_N : 0) | (bB ? PH_BAD : 0);
            a.seq_start[slot] = (uint64_t)((int64_t)st + pos0);          // chunk-relative start
            a.seq_end[slot] = en == 0xFFFFFFFFu ? ~0ull : (uint64_t)((int64_t)en + pos0);
        }
    }
    SP_STAMP(5);
}

static int scan_pack_setup(Ctx* c);

int scan_pack(Ctx* c, const uint8_t* d_buf, uint64_t n, uint64_t line_base, uint64_t rec_first, uint64_t n_slots,
              uint64_t* d_seq_start, uint64_t* d_seq_end, uint32_t* d_rows, uint32_t* d_hdr, uint32_t row_words, uint32_t cap,
              uint64_t* n_terms, unsigned long long* err_flags) {
    *n_terms = 0;
    *err_flags = 0;
    if (n == 0) return VSPE_OK;
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(d_buf) & 15);
    const uint64_t n_tiles = (n + head + SP_TILE - 1) / SP_TILE;
    if (n_tiles > 0x7FFFFFFFull) { set_error("buffer too large for one scan launch"); return VSPE_ERR_LIMIT; }
    VSPE_TRY(c->tile_base.reserve(n_tiles + 4));
    unsigned long long* status = reinterpret_cast<unsigned long long*>(c->tile_base.p);
    VSPE_CUDA(cudaMemsetAsync(status, 0, (n_tiles + 4) * 8, c->stream));
    ScanPackArgs a;
    a.buf = d_buf; a.n = n; a.head = head; a.status = status;
    a.ticket = reinterpret_cast<unsigned int*>(status + n_tiles + 1);
    a.total_out = status + n_tiles + 2;
    a.line_base = line_base; a.rec_first = rec_first; a.n_slots = n_slots;
    a.seq_start = d_seq_start; a.seq_end = d_seq_end; a.rows = d_rows; a.hdr = d_hdr; a.row_words = row_words; a.cap = cap;
    a.n_tiles = (uint32_t)n_tiles; a.counters = c->counters.p;
    a.dbg = nullptr;
    if (c->opt_dbg_times) {
        VSPE_TRY(c->dbg_times.reserve(n_tiles * 8));
        VSPE_CUDA(cudaMemsetAsync(c->dbg_times.p, 0, n_tiles * 64, c->stream));
        a.dbg = c->dbg_times.p;
        c->dbg_tiles = n_tiles;
    }
    // the dominant kernel is timed on its own (CUDA events on the launching stream)
    VSPE_TRY(scan_pack_setup(c));
    VSPE_CUDA(cudaEventRecord(c->ev_scan[0][0], c->stream));
    a.qbuf = nullptr; a.tile_tot = nullptr; a.tile_excl = nullptr;
    k_scan_pack<0><<<(uint32_t)n_tiles, SP_WARPS * 32, SP_SMEM, c->stream>>>(a);
    VSPE_LAUNCH_CHECK(c);
    VSPE_CUDA(cudaEventRecord(c->ev_scan[0][1], c->stream));
    unsigned long long h_total = 0, h_err = 0;
    VSPE_CUDA(cudaMemcpyAsync(&h_total, a.total_out, 8, cudaMemcpyDeviceToHost, c->stream));
    VSPE_CUDA(cudaMemcpyAsync(&h_err, c->counters.p + CNT_ERR, 8, cudaMemcpyDeviceToHost, c->stream));
    VSPE_CUDA(cudaStreamSynchronize(c->stream));
    *n_terms = h_total;
    *err_flags = h_err;
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->ev_scan[0][0], c->ev_scan[0][1]) == cudaSuccess) { c->stats.ms_k_scan_pack += ms; c->stats.n_k_scan_pack++; }
    }
    const unsigned long long transient = ERRF_SLOTS_FULL | ERRF_TILE_FULL;
    if (h_err & transient) {
        unsigned long long cleared = h_err & ~transient;
        VSPE_CUDA(cudaMemcpyAsync(c->counters.p + CNT_ERR, &cleared, 8, cudaMemcpyHostToDevice, c->stream));
        VSPE_CUDA(cudaStreamSynchronize(c->stream));
    }
    return VSPE_OK;
}

// Two-kernel variant: count pass (no inter-tile dependency) -> device scan of the tile totals ->
// the caller sizes the outputs exactly -> pack pass.  `prepare_launch` queues the first two steps
// (no host synchronisation, so the count passes of both mates can be queued back to back),
// `prepare_collect` returns the terminator count, `finish` launches the pack pass.  Scratch and
// timing events are per mate: mate 1's count pass may run before mate 0's pack pass.
void scan_pack_account(Ctx* c);

static int scan_pack_setup(Ctx* c) {
    if (!c->scan_pack_attr_set) {
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_pack<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_pack<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_pack<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
        c->scan_pack_attr_set = true;
    }
    for (int m = 0; m < 2; m++)
        for (int k = 0; k < 4; k++)
            if (!c->ev_scan[m][k]) VSPE_CUDA(cudaEventCreate(&c->ev_scan[m][k]));
    return VSPE_OK;
}

int scan_pack_prepare_launch(Ctx* c, int m, const uint8_t* d_buf, uint64_t n) {
    if (n == 0) return VSPE_OK;
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(d_buf) & 15);
    const uint64_t n_tiles = (n + head + SP_TILE - 1) / SP_TILE;
    if (n_tiles > 0x7FFFFFFFull) { set_error("buffer too large for one scan launch"); return VSPE_ERR_LIMIT; }
    VSPE_TRY(c->scan_tiles[m].reserve(2 * n_tiles + n_tiles / 1024 + 16));
    VSPE_TRY(c->scan_q[m].reserve(n_tiles * SP_WARPS * SP_QBYTES + 64));
    VSPE_TRY(scan_pack_setup(c));
    unsigned long long* tot = reinterpret_cast<unsigned long long*>(c->scan_tiles[m].p);
    unsigned long long* excl = tot + n_tiles;
    unsigned long long* sums = excl + n_tiles;                 // [n_tiles / 2048 + 1] scratch, then [.. + 8] the grand total
    unsigned long long* d_total = sums + n_tiles / 1024 + 8;
    ScanPackArgs a = {};
    a.buf = d_buf; a.n = n; a.head = head; a.n_tiles = (uint32_t)n_tiles; a.counters = c->counters.p;
    a.qbuf = c->scan_q[m].p; a.tile_tot = tot; a.tile_excl = excl;
    a.row_words = 16; a.cap = 256;
    a.tile_full = d_total + 2;
    VSPE_CUDA(cudaMemsetAsync(a.tile_full, 0, 8, c->stream));
    VSPE_CUDA(cudaEventRecord(c->ev_scan[m][2], c->stream));
    k_scan_pack<1><<<(uint32_t)n_tiles, SP_WARPS * 32, SP_SMEM, c->stream>>>(a);
    VSPE_LAUNCH_CHECK(c);
    VSPE_CUDA(cudaEventRecord(c->ev_scan[m][3], c->stream));
    VSPE_TRY(device_scan_u64(c, tot, excl, n_tiles, sums, d_total));
    c->scan_count_pending[m] = true;
    return VSPE_OK;
}

int scan_pack_prepare_collect(Ctx* c, int m, const uint8_t* d_buf, uint64_t n, uint64_t* n_terms, unsigned long long* err_flags) {
    *n_terms = 0;
    *err_flags = 0;
    if (n == 0) return VSPE_OK;
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(d_buf) & 15);
    const uint64_t n_tiles = (n + head + SP_TILE - 1) / SP_TILE;
    unsigned long long* tot = reinterpret_cast<unsigned long long*>(c->scan_tiles[m].p);
    unsigned long long* d_total = tot + 2 * n_tiles + n_tiles / 1024 + 8;
    unsigned long long h[3] = {0, 0, 0};                       // grand total, (unused), queue-overflow flag of this mate's count pass
    VSPE_CUDA(cudaMemcpyAsync(h, d_total, 24, cudaMemcpyDeviceToHost, c->stream));
    VSPE_CUDA(cudaStreamSynchronize(c->stream));
    *n_terms = h[0];
    *err_flags = h[2] ? ERRF_TILE_FULL : 0;
    scan_pack_account(c);                                      // everything queued before is done by now
    VSPE_TRY(adapt_map_variant(c));
    return VSPE_OK;
}

int scan_pack_finish(Ctx* c, int m, const uint8_t* d_buf, uint64_t n, uint64_t line_base, uint64_t rec_first, uint64_t n_slots,
                     uint64_t* d_seq_start, uint64_t* d_seq_end, uint32_t* d_rows, uint32_t* d_hdr, uint32_t row_words, uint32_t cap) {
    if (n == 0) return VSPE_OK;
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(d_buf) & 15);
    const uint64_t n_tiles = (n + head + SP_TILE - 1) / SP_TILE;
    unsigned long long* tot = reinterpret_cast<unsigned long long*>(c->scan_tiles[m].p);
    ScanPackArgs a = {};
    a.buf = d_buf; a.n = n; a.head = head; a.n_tiles = (uint32_t)n_tiles; a.counters = c->counters.p;
    a.qbuf = c->scan_q[m].p; a.tile_tot = tot; a.tile_excl = tot + n_tiles;
    a.line_base = line_base; a.rec_first = rec_first; a.n_slots = n_slots;
    a.seq_start = d_seq_start; a.seq_end = d_seq_end; a.rows = d_rows; a.hdr = d_hdr; a.row_words = row_words; a.cap = cap;
    a.total_out = tot + 2 * n_tiles + n_tiles / 1024 + 9;      // unused scratch word
    VSPE_CUDA(cudaEventRecord(c->ev_scan[m][0], c->stream));
    k_scan_pack<2><<<(uint32_t)n_tiles, SP_WARPS * 32, SP_SMEM, c->stream>>>(a);
    VSPE_LAUNCH_CHECK(c);
    VSPE_CUDA(cudaEventRecord(c->ev_scan[m][1], c->stream));
    c->scan_pack_pending[m] = true;
    return VSPE_OK;
}

// fold the durations of the scan launches whose events have completed into the stats (call after
// a stream sync; a launch still in flight stays pending)
void scan_pack_account(Ctx* c) {
    for (int m = 0; m < 2; m++) {
        float ms = 0;
        if (c->scan_count_pending[m] && cudaEventQuery(c->ev_scan[m][3]) == cudaSuccess) {
            if (cudaEventElapsedTime(&ms, c->ev_scan[m][2], c->ev_scan[m][3]) == cudaSuccess) { c->stats.ms_k_scan_count += ms; c->stats.n_k_scan_count++; }
            c->scan_count_pending[m] = false;
        }
        if (c->scan_pack_pending[m] && cudaEventQuery(c->ev_scan[m][1]) == cudaSuccess) {
            if (cudaEventElapsedTime(&ms, c->ev_scan[m][0], c->ev_scan[m][1]) == cudaSuccess) { c->stats.ms_k_scan_pack += ms; c->stats.n_k_scan_pack++; }
            c->scan_pack_pending[m] = false;
        }
    }
    cudaGetLastError();                                        // cudaEventQuery on an unfinished event is not an error here
}

}  // namespace vspe
