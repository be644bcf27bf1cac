// scan.cu -- K1: FASTQ record split with universal-newline semantics.
//
// Replaces `readlines()` + `[s[:-1] for s in reads[4r : 4r+4]]` of reference
// utils/VStrains_PE_Inference.py:149-159: a line ends at '\n', at "\r\n" (one terminator)
// or at a lone '\r'; the sequence of record r is the content of line 4r+1.
//
// A *terminator* is a byte position p with  buf[p]=='\n'  or  (buf[p]=='\r' and buf[p+1]!='\n').
// The line index of a terminator is the number of terminators before it.
#include "ctx.cuh"

namespace vspe {

static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITERS = 4;                       // 16-byte vectors per thread
static constexpr int SCAN_TILE = SCAN_THREADS * 16 * SCAN_ITERS;   // 16 KiB per block

__device__ __forceinline__ uint32_t movemask4(uint32_t cmp) {   // 0xFF/0x00 bytes -> 4 bits
    return ((cmp & 0x80808080u) * 0x00204081u) >> 28;
}

// Terminator bitmask of the 16 bytes at [p0, p0+16) of buf (positions outside [0,n) excluded).
// Also reports non-ASCII bytes.
__device__ __forceinline__ uint32_t term_mask16(const uint8_t* __restrict__ buf, uint64_t n, int64_t p0,
                                                bool& non_ascii, uint4& v) {
    if (p0 >= 0 && (uint64_t)p0 + 16 <= n) {
        v = __ldg(reinterpret_cast<const uint4*>(buf + p0));
    } else {
        uint32_t w[4] = {0, 0, 0, 0};
        for (int i = 0; i < 16; i++) {
            int64_t p = p0 + i;
            uint32_t c = (p >= 0 && (uint64_t)p < n) ? buf[p] : 0u;
            w[i >> 2] |= c << (8 * (i & 3));
        }
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if ((v.x | v.y | v.z | v.w) & 0x80808080u) non_ascii = true;
    uint32_t nl = movemask4(__vcmpeq4(v.x, 0x0A0A0A0Au)) | (movemask4(__vcmpeq4(v.y, 0x0A0A0A0Au)) << 4) |
                  (movemask4(__vcmpeq4(v.z, 0x0A0A0A0Au)) << 8) | (movemask4(__vcmpeq4(v.w, 0x0A0A0A0Au)) << 12);
    uint32_t cr = movemask4(__vcmpeq4(v.x, 0x0D0D0D0Du)) | (movemask4(__vcmpeq4(v.y, 0x0D0D0D0Du)) << 4) |
                  (movemask4(__vcmpeq4(v.z, 0x0D0D0D0Du)) << 8) | (movemask4(__vcmpeq4(v.w, 0x0D0D0D0Du)) << 12);
    uint32_t term = nl;
    if (cr) {
        int64_t pn = p0 + 16;
        uint32_t next_nl = (pn >= 0 && (uint64_t)pn < n && buf[pn] == '\n') ? 0x8000u : 0u;
        term |= cr & ~((nl >> 1) | next_nl);
    }
    return term;     // bytes outside [0,n) were loaded as 0 -> never terminators
}

// block-wide exclusive scan of one value per thread (SCAN_THREADS threads); returns total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t x, uint32_t* s_warp, uint32_t& total) {
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = x;
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += y;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        uint32_t t = s_warp[w];
        if (w < wid) base += t;
        tot += t;
    }
    __syncthreads();
    total = tot;
    return base + inc - x;
}

// The buffer is processed in 16-byte vectors aligned to the ABSOLUTE address, so shards that
// start at arbitrary byte offsets still get aligned 128-bit loads.  Vector g covers buffer
// positions [16g - head, 16g - head + 16) where head = address of buf mod 16.
__global__ void __launch_bounds__(SCAN_THREADS)
k_count_terms(const uint8_t* __restrict__ buf, uint64_t n, uint32_t head, uint32_t* __restrict__ tile_counts,
              unsigned long long* __restrict__ counters) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    uint32_t cnt = 0;
    bool bad = false;
    for (int it = 0; it < SCAN_ITERS; it++) {
        uint64_t g = ((uint64_t)blockIdx.x * SCAN_ITERS + it) * SCAN_THREADS + threadIdx.x;
        int64_t p0 = (int64_t)(g * 16) - head;
        if (p0 < (int64_t)n) {
            uint4 v;
            cnt += __popc(term_mask16(buf, n, p0, bad, v));
        }
    }
    uint32_t total;
    block_excl_scan(cnt, s_warp, total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
    if (bad) atomicOr(&counters[CNT_ERR], (unsigned long long)ERRF_NON_ASCII);
}

// single-block exclusive scan of tile counts -> tile_base (u64) and grand total
__global__ void __launch_bounds__(1024)
k_scan_tiles(const uint32_t* __restrict__ counts, uint64_t n_tiles, uint64_t* __restrict__ base,
             unsigned long long* __restrict__ total_out) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint64_t off = 0; off < n_tiles; off += 1024) {
        uint64_t i = off + threadIdx.x;
        uint64_t x = i < n_tiles ? counts[i] : 0, inc = x;
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += y;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        uint64_t wbase = 0, tot = 0;
        for (int w = 0; w < 32; w++) {
            uint64_t t = s_warp[w];
            if (w < (int)wid) wbase += t;
            tot += t;
        }
        uint64_t carry = s_carry;
        if (i < n_tiles) base[i] = carry + wbase + inc - x;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = s_carry;
}

// For every terminator: line = line_base + rank.  line%4==0 ends a header -> the sequence of
// record line/4 starts at p+1; line%4==1 ends a sequence line -> its content ends at p
// (or p-1 for "\r\n").
__global__ void __launch_bounds__(SCAN_THREADS)
k_index_records(const uint8_t* __restrict__ buf, uint64_t n, uint32_t head, const uint64_t* __restrict__ tile_base,
                uint64_t line_base, uint64_t rec_first, uint64_t n_slots,
                uint64_t* __restrict__ seq_start, uint64_t* __restrict__ seq_end) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    uint64_t running = line_base + tile_base[blockIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 0 && (line_base & 3) == 1 && n_slots > 0) seq_start[0] = 0;
    for (int it = 0; it < SCAN_ITERS; it++) {
        uint64_t g = ((uint64_t)blockIdx.x * SCAN_ITERS + it) * SCAN_THREADS + threadIdx.x;
        int64_t p0 = (int64_t)(g * 16) - head;
        uint32_t mask = 0;
        uint4 v = make_uint4(0, 0, 0, 0);
        bool bad = false;
        if (p0 < (int64_t)n) mask = term_mask16(buf, n, p0, bad, v);
        uint32_t total;
        uint32_t excl = block_excl_scan(__popc(mask), s_warp, total);
        uint64_t line = running + excl;
        while (mask) {
            int i = __ffs(mask) - 1;
            mask &= mask - 1;
            uint64_t p = (uint64_t)(p0 + i);
            uint32_t phase = (uint32_t)line & 3;
            if (phase == 0) {
                uint64_t idx = (line >> 2) - rec_first;
                if (idx < n_slots) seq_start[idx] = p + 1;
            } else if (phase == 1) {
                uint64_t idx = (line >> 2) - rec_first;
                uint32_t wv = i < 4 ? v.x : i < 8 ? v.y : i < 12 ? v.z : v.w;
                uint32_t c = (wv >> (8 * (i & 3))) & 0xFF;
                uint64_t e = p;
                if (c == '\n' && p > 0 && buf[p - 1] == '\r') e = p - 1;
                if (idx < n_slots) seq_end[idx] = e;
            }
            line++;
        }
        running += total;
    }
}

// ---------------------------------------------------------------------------------------
// Single-pass variant: decoupled look-back over 16 KiB tiles (one read of the input).
// Every tile publishes {flag, count} in one 64-bit word: flag 1 = tile aggregate, 2 = inclusive
// prefix.  Tiles take their index from an atomic ticket so every predecessor is already running.
// ---------------------------------------------------------------------------------------
#define LB_AGG (1ull << 62)
#define LB_INC (2ull << 62)
#define LB_VAL ((1ull << 62) - 1)

__device__ __forceinline__ void masks16(const uint8_t* __restrict__ buf, uint64_t n, int64_t p0, bool& non_ascii,
                                        uint32_t& term, uint32_t& crlf_nl) {
    // term: terminator positions; crlf_nl: '\n' terminators preceded by '\r' (content ends one earlier)
    uint4 v;
    if (p0 >= 0 && (uint64_t)p0 + 16 <= n) {
        v = __ldg(reinterpret_cast<const uint4*>(buf + p0));
    } else {
        uint32_t w[4] = {0, 0, 0, 0};
        for (int i = 0; i < 16; i++) {
            int64_t p = p0 + i;
            uint32_t c = (p >= 0 && (uint64_t)p < n) ? buf[p] : 0u;
            w[i >> 2] |= c << (8 * (i & 3));
        }
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if ((v.x | v.y | v.z | v.w) & 0x80808080u) non_ascii = true;
    // bytes <= 0x0F are rare in FASTQ: test for them before building exact masks
    const uint32_t lowx = ~(v.x | (v.x >> 1) | (v.x >> 2) | (v.x >> 3)) , lowy = ~(v.y | (v.y >> 1) | (v.y >> 2) | (v.y >> 3)),
                   lowz = ~(v.z | (v.z >> 1) | (v.z >> 2) | (v.z >> 3)), loww = ~(v.w | (v.w >> 1) | (v.w >> 2) | (v.w >> 3));
    term = 0;
    crlf_nl = 0;
    if (((lowx | lowy | lowz | loww) & 0x10101010u) == 0) return;     // no byte with a zero high nibble
    uint32_t nl = movemask4(__vcmpeq4(v.x, 0x0A0A0A0Au)) | (movemask4(__vcmpeq4(v.y, 0x0A0A0A0Au)) << 4) |
                  (movemask4(__vcmpeq4(v.z, 0x0A0A0A0Au)) << 8) | (movemask4(__vcmpeq4(v.w, 0x0A0A0A0Au)) << 12);
    uint32_t cr = movemask4(__vcmpeq4(v.x, 0x0D0D0D0Du)) | (movemask4(__vcmpeq4(v.y, 0x0D0D0D0Du)) << 4) |
                  (movemask4(__vcmpeq4(v.z, 0x0D0D0D0Du)) << 8) | (movemask4(__vcmpeq4(v.w, 0x0D0D0D0Du)) << 12);
    // positions outside [0,n) were loaded as 0 and are neither
    term = nl;
    if (cr) {
        int64_t pn = p0 + 16;
        uint32_t next_nl = (pn >= 0 && (uint64_t)pn < n && buf[pn] == '\n') ? 0x8000u : 0u;
        term |= cr & ~((nl >> 1) | next_nl);
    }
    crlf_nl = nl & (cr << 1);
    if ((nl & 1) && p0 > 0 && buf[p0 - 1] == '\r') crlf_nl |= 1;
}

static constexpr int SR_WARPS = 8;                       // warps per block
static constexpr int SR_ITERS = 16;                      // 512-byte warp rows per warp
static constexpr int SR_TILE = SR_WARPS * SR_ITERS * 32 * 16;   // 64 KiB per block

// Layout inside a tile: warp w owns the contiguous 8 KiB [w*8K, (w+1)*8K); in iteration `it` its
// lanes read one coalesced 512-byte row.  Line order = (warp, iteration, lane), so the ranks come
// from warp scans (shuffles) plus ONE block-level exchange of the 8 warp totals.
__global__ void __launch_bounds__(SR_WARPS * 32)
k_scan_records(const uint8_t* __restrict__ buf, uint64_t n, uint32_t head, unsigned long long* __restrict__ status,
               unsigned int* __restrict__ ticket, uint64_t line_base, uint64_t rec_first, uint64_t n_slots,
               uint64_t* __restrict__ seq_start, uint64_t* __restrict__ seq_end, unsigned long long* __restrict__ total_out,
               uint32_t n_tiles, unsigned long long* __restrict__ counters) {
    __shared__ uint32_t s_wtot[SR_WARPS];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t g0 = ((uint64_t)tile * SR_WARPS + wib) * SR_ITERS * 32 + lane;    // first 16-byte vector of this lane
    uint32_t mk[SR_ITERS], cnt = 0;                           // terminator mask | crlf mask << 16
    bool bad = false;
#pragma unroll
    for (int it = 0; it < SR_ITERS; it++) {
        const int64_t p0 = (int64_t)((g0 + (uint64_t)it * 32) * 16) - head;
        uint32_t term = 0, crlf = 0;
        if (p0 < (int64_t)n) masks16(buf, n, p0, bad, term, crlf);
        mk[it] = term | (crlf << 16);
        cnt += __popc(term);
    }
    if (bad) atomicOr(&counters[CNT_ERR], (unsigned long long)ERRF_NON_ASCII);
    const uint32_t wtot = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if (lane == 0) s_wtot[wib] = wtot;
    __syncthreads();
    uint32_t tile_total = 0, warp_base = 0;
#pragma unroll
    for (int w = 0; w < SR_WARPS; w++) {
        const uint32_t x = s_wtot[w];
        if (w < (int)wib) warp_base += x;
        tile_total += x;
    }
    // publish the aggregate, look back for the exclusive prefix (warp 0)
    if (wib == 0) {
        volatile unsigned long long* vs = status;
        if (tile == 0) {
            if (lane == 0) { vs[0] = LB_INC | tile_total; s_excl = 0; }
        } else {
            if (lane == 0) vs[tile] = LB_AGG | tile_total;
            unsigned long long excl = 0;
            int64_t look = (int64_t)tile - 1;
            while (true) {
                const int64_t idx = look - lane;
                unsigned long long st = idx >= 0 ? vs[idx] : LB_INC;
                while (__any_sync(0xFFFFFFFFu, (st >> 62) == 0)) {
                    if ((st >> 62) == 0) st = vs[idx];
                }
                const uint32_t inc = __ballot_sync(0xFFFFFFFFu, (st >> 62) == 2);
                const int first = inc ? __ffs((int)inc) - 1 : 32;
                unsigned long long c = (int)lane <= first ? (st & LB_VAL) : 0ull;
                for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
                excl += c;
                if (inc) break;
                look -= 32;
            }
            if (lane == 0) { vs[tile] = LB_INC | (excl + tile_total); s_excl = excl; }
        }
        if (lane == 0 && tile == n_tiles - 1) *total_out = s_excl + tile_total;
    }
    __syncthreads();
    uint64_t running = line_base + s_excl + warp_base;
    if (tile == 0 && threadIdx.x == 0 && (line_base & 3) == 1 && n_slots > 0) seq_start[0] = 0;
    bool overflow = false;
#pragma unroll
    for (int it = 0; it < SR_ITERS; it++) {
        uint32_t mask = mk[it] & 0xFFFFu;
        const uint32_t crlf = mk[it] >> 16;
        const uint32_t c = __popc(mask);
        uint32_t inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        const uint32_t row_total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        uint64_t line = running + inc - c;
        running += row_total;
        if (!mask) continue;
        const int64_t p0 = (int64_t)((g0 + (uint64_t)it * 32) * 16) - head;
        while (mask) {
            const int i = __ffs((int)mask) - 1;
            mask &= mask - 1;
            const uint64_t p = (uint64_t)(p0 + i);
            const uint32_t phase = (uint32_t)line & 3;
            if (phase < 2) {
                const uint64_t idx = (line >> 2) - rec_first;
                if (idx < n_slots) {
                    if (phase == 0) seq_start[idx] = p + 1;
                    else seq_end[idx] = ((crlf >> i) & 1) ? p - 1 : p;
                } else if (phase == 1) {
                    overflow = true;
                }
            }
            line++;
        }
    }
    if (overflow) atomicOr(&counters[CNT_ERR], (unsigned long long)ERRF_SLOTS_FULL);
}

static inline uint32_t head_of(const uint8_t* p) { return (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15); }

int scan_count_lines(Ctx* c, const uint8_t* d_buf, uint64_t n, uint64_t* n_terms) {
    *n_terms = 0;
    if (n == 0) return VSPE_OK;
    uint32_t head = head_of(d_buf);
    uint64_t n_tiles = (n + head + SCAN_TILE - 1) / SCAN_TILE;
    if (n_tiles > 0x7FFFFFFFull) { set_error("buffer too large for one scan launch"); return VSPE_ERR_LIMIT; }
    VSPE_TRY(c->tile_counts.reserve(n_tiles));
    VSPE_TRY(c->tile_base.reserve(n_tiles + 1));
    k_count_terms<<<(uint32_t)n_tiles, SCAN_THREADS, 0, c->stream>>>(d_buf, n, head, c->tile_counts.p, c->counters.p);
    VSPE_LAUNCH_CHECK(c);
    k_scan_tiles<<<1, 1024, 0, c->stream>>>(c->tile_counts.p, n_tiles, c->tile_base.p,
                                            (unsigned long long*)(c->tile_base.p + n_tiles));
    VSPE_LAUNCH_CHECK(c);
    unsigned long long total = 0;
    VSPE_CUDA(cudaMemcpyAsync(&total, c->tile_base.p + n_tiles, 8, cudaMemcpyDeviceToHost, c->stream));
    VSPE_CUDA(cudaStreamSynchronize(c->stream));
    *n_terms = total;
    return VSPE_OK;
}

// single pass: terminator count AND record table; *overflow is set when n_slots was too small
// (the caller then retries with the exact size, which is known from *n_terms).
int scan_records_single_pass(Ctx* c, const uint8_t* d_buf, uint64_t n, uint64_t line_base, uint64_t rec_first,
                             uint64_t n_slots, uint64_t* d_seq_start, uint64_t* d_seq_end, uint64_t* n_terms, bool* overflow) {
    *n_terms = 0;
    *overflow = false;
    if (n == 0) return VSPE_OK;
    uint32_t head = head_of(d_buf);
    uint64_t n_tiles = (n + head + SR_TILE - 1) / SR_TILE;
    if (n_tiles > 0x7FFFFFFFull) { set_error("buffer too large for one scan launch"); return VSPE_ERR_LIMIT; }
    VSPE_TRY(c->tile_base.reserve(n_tiles + 4));
    unsigned long long* status = reinterpret_cast<unsigned long long*>(c->tile_base.p);
    VSPE_CUDA(cudaMemsetAsync(status, 0, (n_tiles + 4) * 8, c->stream));
    unsigned int* ticket = reinterpret_cast<unsigned int*>(status + n_tiles + 1);
    unsigned long long* total = status + n_tiles + 2;
    k_scan_records<<<(uint32_t)n_tiles, SR_WARPS * 32, 0, c->stream>>>(d_buf, n, head, status, ticket, line_base, rec_first, n_slots,
                                                                      d_seq_start, d_seq_end, total, (uint32_t)n_tiles, c->counters.p);
    VSPE_LAUNCH_CHECK(c);
    unsigned long long h_total = 0, h_err = 0;
    VSPE_CUDA(cudaMemcpyAsync(&h_total, total, 8, cudaMemcpyDeviceToHost, c->stream));
    VSPE_CUDA(cudaMemcpyAsync(&h_err, c->counters.p + CNT_ERR, 8, cudaMemcpyDeviceToHost, c->stream));
    VSPE_CUDA(cudaStreamSynchronize(c->stream));
    *n_terms = h_total;
    if (h_err & ERRF_SLOTS_FULL) {
        *overflow = true;
        unsigned long long cleared = h_err & ~(unsigned long long)ERRF_SLOTS_FULL;
        VSPE_CUDA(cudaMemcpyAsync(c->counters.p + CNT_ERR, &cleared, 8, cudaMemcpyHostToDevice, c->stream));
        VSPE_CUDA(cudaStreamSynchronize(c->stream));
    }
    return VSPE_OK;
}

int scan_index_records(Ctx* c, const uint8_t* d_buf, uint64_t n, uint64_t line_base, uint64_t rec_first,
                       uint64_t n_slots, uint64_t* d_seq_start, uint64_t* d_seq_end) {
    if (n == 0) return VSPE_OK;
    uint32_t head = head_of(d_buf);
    uint64_t n_tiles = (n + head + SCAN_TILE - 1) / SCAN_TILE;
    k_index_records<<<(uint32_t)n_tiles, SCAN_THREADS, 0, c->stream>>>(d_buf, n, head, c->tile_base.p, line_base,
                                                                       rec_first, n_slots, d_seq_start, d_seq_end);
    VSPE_LAUNCH_CHECK(c);
    return VSPE_OK;
}

}  // namespace vspe
