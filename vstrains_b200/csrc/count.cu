// count.cu -- K5 + K6: link keys and their deterministic counting.
//
// Replaces the accumulation loops of reference utils/VStrains_PE_Inference.py:160-188:
//   N / short filtering (:160-163), short_mat[a][b] += 1 for a <= b within each mate's node
//   list (:174-184), node_mat[i][j] += 1 for i in lefts, j in rights (:186-188).
//
// A key is the dense offset into [node_mat | short_mat]:  key = mat*N*N + i*N + j.
// Counting is a two-digit most-significant-digit radix sort whose last digit is fused with the
// run-length reduce:
//   pass 1  (k_pair_count)   per-bucket key histogram (bucket = key >> low_bits) + pair counters
//   pass 2  (k_bucket_scan)  exclusive scan -> bucket segments
//   pass 3  (k_pair_emit)    keys scattered into their bucket segment (radix partition)
//   pass 4  (k_bucket_hist)  one CTA per bucket: counting sort of the low digit in shared
//                            memory; the per-value counts ARE the run lengths, added to the
//                            dense matrices with plain (non-atomic) stores -- each matrix cell
//                            is owned by exactly one CTA, so the result is order independent.
#include <algorithm>

#include "ctx.cuh"

namespace vspe {

static constexpr int PAIR_THREADS = 256;
static constexpr int PAIR_WARPS = PAIR_THREADS / 32;
static constexpr int STAGE_STRIDE = 17;                               // words per staged slot: 16 + 1 pad => per-thread reads hit 32 banks
static constexpr int STAGE_WORDS = PAIR_WARPS * 64 * STAGE_STRIDE;    // per block: every warp stages 32 pairs x 2 mates
static constexpr uint32_t HIST_SLICE = 16384;                         // keys per k_bucket_hist work item

struct ListRef {
    const uint32_t* ids;
    uint32_t n;
};

// list of a slot staged in shared memory (words: hdr, ids[15]); long lists live in the spill pool
__device__ __forceinline__ ListRef list_of(const uint32_t* s, const uint32_t* __restrict__ spill) {
    ListRef r;
    r.n = s[0] >> 8;
    r.ids = r.n <= SLOT_IDS ? s + 1 : spill + s[1];
    return r;
}

// keys are dense offsets < 2 * N * N <= 2^32 (count_pairs checks), so 32-bit arithmetic is exact
template <class F>
__device__ __forceinline__ void for_each_key(const ListRef& l, const ListRef& r, uint32_t N, F fn) {
    const uint32_t NN = N * N;
    for (uint32_t a = 0; a < l.n; a++) {
        const uint32_t row = NN + l.ids[a] * N;
        for (uint32_t b = a; b < l.n; b++) fn(row + l.ids[b]);
    }
    for (uint32_t a = 0; a < r.n; a++) {
        const uint32_t row = NN + r.ids[a] * N;
        for (uint32_t b = a; b < r.n; b++) fn(row + r.ids[b]);
    }
    for (uint32_t a = 0; a < l.n; a++) {
        const uint32_t row = l.ids[a] * N;
        for (uint32_t b = 0; b < r.n; b++) fn(row + r.ids[b]);
    }
}

// 0 used, 1 N-pair, 2 short pair  (N takes precedence: PE_Inference.py:160 before :162)
__device__ __forceinline__ uint32_t pair_class(uint32_t hf, uint32_t hr) {
    uint32_t sf = hf & 0xFF, sr = hr & 0xFF;
    if (sf == ST_N || sr == ST_N) return 1;
    if (sf == ST_SHORT || sr == ST_SHORT) return 2;
    return 0;
}

// The ReadSlots of 32 consecutive pairs are 2 KiB contiguous per mate: the warp copies them to
// its shared-memory area with coalesced 16-byte loads (all eight issued before the first store).
__device__ __forceinline__ void stage_slots(const ReadSlot* __restrict__ f, const ReadSlot* __restrict__ r, uint64_t p0,
                                            uint64_t n_pairs, uint32_t* st, uint32_t lane) {
    uint4 v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint32_t c = lane + 32 * (q & 3);                       // 16-byte chunk of the mate's 2 KiB
        const uint4* src = reinterpret_cast<const uint4*>((q < 4 ? f : r) + p0);
        v[q] = p0 + (c >> 2) < n_pairs ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint32_t c = lane + 32 * (q & 3);
        uint32_t* d = st + ((q < 4 ? 0u : 32u) + (c >> 2)) * STAGE_STRIDE + (c & 3) * 4;
        d[0] = v[q].x; d[1] = v[q].y; d[2] = v[q].z; d[3] = v[q].w;
    }
    __syncwarp();
}

// Every block owns the contiguous pairs [blockIdx.x * ppb, + ppb) -- in BOTH kernels, so the
// per-block bucket histogram k_pair_count stores is exactly what k_pair_emit needs for its
// cursors.  A few hundred resident blocks => the histogram epilogue / cursor prologue (one pass
// over all buckets each) is paid once per ~2000 pairs.
__global__ void __launch_bounds__(PAIR_THREADS)
k_pair_count(const ReadSlot* __restrict__ f, const ReadSlot* __restrict__ r, uint64_t n_pairs, uint32_t ppb, uint64_t N,
             const uint32_t* __restrict__ spill, uint32_t low_bits, uint32_t n_buckets,
             unsigned long long* __restrict__ g_hist, unsigned long long* __restrict__ counters,
             uint32_t* __restrict__ blk_hist) {
    extern __shared__ uint32_t s_dyn[];
    uint32_t* s_hist = s_dyn;                                         // [n_buckets]
    uint32_t* s_stage = s_dyn + n_buckets;                            // [PAIR_WARPS][64][STAGE_STRIDE]
    __shared__ unsigned long long s_cnt[4];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (uint32_t b = threadIdx.x; b < n_buckets; b += PAIR_THREADS) s_hist[b] = 0;
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t lo = (uint64_t)blockIdx.x * ppb;
    const uint64_t hi = lo + ppb < n_pairs ? lo + ppb : n_pairs;
    uint32_t* st = s_stage + wib * 64 * STAGE_STRIDE;
    uint32_t c_used = 0, c_n = 0, c_short = 0;
    unsigned long long c_keys = 0;
    for (uint64_t p0 = lo + wib * 32; p0 < hi; p0 += PAIR_THREADS) {
        __syncwarp();                                                 // the previous round's reads are done
        stage_slots(f, r, p0, hi, st, lane);
        if (p0 + lane < hi) {
            const uint32_t* sf = st + lane * STAGE_STRIDE;
            const uint32_t* sr = st + (32 + lane) * STAGE_STRIDE;
            const uint32_t cls = pair_class(sf[0], sr[0]);
            if (cls == 0) {
                ListRef l = list_of(sf, spill), rr = list_of(sr, spill);
                uint32_t m = 0;
                for_each_key(l, rr, (uint32_t)N, [&](uint32_t key) { atomicAdd(&s_hist[key >> low_bits], 1u); m++; });
                c_keys += m;
                c_used++;
            } else if (cls == 1) {
                c_n++;
            } else {
                c_short++;
            }
        }
    }
    if (c_used) atomicAdd(&s_cnt[0], (unsigned long long)c_used);
    if (c_n) atomicAdd(&s_cnt[1], (unsigned long long)c_n);
    if (c_short) atomicAdd(&s_cnt[2], (unsigned long long)c_short);
    if (c_keys) atomicAdd(&s_cnt[3], c_keys);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < n_buckets; b += PAIR_THREADS) {
        uint32_t c = s_hist[b];
        if (c) atomicAdd(&g_hist[b], (unsigned long long)c);
        blk_hist[(uint64_t)blockIdx.x * n_buckets + b] = c;           // k_pair_emit reuses it
    }
    if (threadIdx.x == 0) {
        if (s_cnt[0]) atomicAdd(&counters[CNT_USED], s_cnt[0]);
        if (s_cnt[1]) atomicAdd(&counters[CNT_N], s_cnt[1]);
        if (s_cnt[2]) atomicAdd(&counters[CNT_SHORT], s_cnt[2]);
        if (s_cnt[3]) atomicAdd(&counters[CNT_KEYS], s_cnt[3]);
    }
}

// Exclusive scan of the bucket histogram (n <= 8192) -> start[0..n], cursor = start; and of the
// number of HIST_SLICE-key work items per bucket -> item_start[0..n] (k_bucket_hist's work list).
__global__ void __launch_bounds__(1024)
k_bucket_scan(const unsigned long long* __restrict__ hist, uint32_t n, unsigned long long* __restrict__ start,
              unsigned long long* __restrict__ cursor, unsigned long long* __restrict__ item_start) {
    __shared__ unsigned long long s_warp[2][32];
    __shared__ unsigned long long s_carry[2];
    if (threadIdx.x < 2) s_carry[threadIdx.x] = 0;
    __syncthreads();
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t off = 0; off < n; off += 1024) {
        uint32_t i = off + threadIdx.x;
        unsigned long long x = i < n ? hist[i] : 0, inc = x;
        unsigned long long y = (x + HIST_SLICE - 1) / HIST_SLICE, yinc = y;
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, yinc, d);
            if (lane >= d) { inc += t; yinc += u; }
        }
        if (lane == 31) { s_warp[0][wid] = inc; s_warp[1][wid] = yinc; }
        __syncthreads();
        unsigned long long wbase = 0, tot = 0, ywbase = 0, ytot = 0;
        for (uint32_t w = 0; w < 32; w++) {
            unsigned long long t = s_warp[0][w], u = s_warp[1][w];
            if (w < wid) { wbase += t; ywbase += u; }
            tot += t;
            ytot += u;
        }
        unsigned long long carry = s_carry[0], ycarry = s_carry[1];
        if (i < n) {
            start[i] = carry + wbase + inc - x;
            cursor[i] = carry + wbase + inc - x;
            item_start[i] = ycarry + ywbase + yinc - y;
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_carry[0] = carry + tot; s_carry[1] = ycarry + ytot; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { start[n] = s_carry[0]; item_start[n] = s_carry[1]; }
}

__global__ void __launch_bounds__(PAIR_THREADS)
k_pair_emit(const ReadSlot* __restrict__ f, const ReadSlot* __restrict__ r, uint64_t n_pairs, uint32_t ppb, uint64_t N,
            const uint32_t* __restrict__ spill, uint32_t low_bits, uint32_t n_buckets,
            unsigned long long* __restrict__ g_cursor, uint32_t* __restrict__ keys, const uint32_t* __restrict__ blk_hist) {
    extern __shared__ unsigned long long s_mem[];
    unsigned long long* s_base = s_mem;                                   // [n_buckets]
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_mem + n_buckets);    // [n_buckets]
    uint32_t* s_stage = s_hist + n_buckets;                               // [PAIR_WARPS][64][STAGE_STRIDE]
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // this block's keys per bucket (stored by k_pair_count: same block -> pairs mapping) reserve
    // its segment of every bucket
    for (uint32_t b = threadIdx.x; b < n_buckets; b += PAIR_THREADS) {
        const uint32_t c = __ldg(blk_hist + (uint64_t)blockIdx.x * n_buckets + b);
        if (c) s_base[b] = atomicAdd(&g_cursor[b], (unsigned long long)c);
        s_hist[b] = 0;
    }
    __syncthreads();
    const uint64_t lo = (uint64_t)blockIdx.x * ppb;
    const uint64_t hi = lo + ppb < n_pairs ? lo + ppb : n_pairs;
    uint32_t* st = s_stage + wib * 64 * STAGE_STRIDE;
    for (uint64_t p0 = lo + wib * 32; p0 < hi; p0 += PAIR_THREADS) {
        __syncwarp();
        stage_slots(f, r, p0, hi, st, lane);
        if (p0 + lane < hi) {
            const uint32_t* sf = st + lane * STAGE_STRIDE;
            const uint32_t* sr = st + (32 + lane) * STAGE_STRIDE;
            if (pair_class(sf[0], sr[0]) == 0) {
                ListRef l = list_of(sf, spill), rr = list_of(sr, spill);
                for_each_key(l, rr, (uint32_t)N, [&](uint32_t key) {
                    const uint32_t b = key >> low_bits;
                    const uint32_t o = atomicAdd(&s_hist[b], 1u);
                    keys[s_base[b] + o] = key;
                });
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-flat key enumeration (option count_flat, default off -- candidate for the next round, see
// DESIGN.md section 10): the nested loops of for_each_key run as many turns as the pair with the most
// keys in the warp.  Here the warp first learns how many keys each of its 32 staged pairs has
// (a closed form of the two list lengths), scans those counts, and then walks the M keys of the
// round 32 at a time: key t belongs to the pair whose prefix interval holds t (binary search over
// the 32 prefixes in shared memory) and its local index decodes to (a, b).  Same keys, same
// per-block bucket counts, balanced lanes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tri(uint32_t n) { return n * (n + 1) / 2; }

template <bool EMIT>
__global__ void __launch_bounds__(PAIR_THREADS)
k_pair_flat(const ReadSlot* __restrict__ f, const ReadSlot* __restrict__ r, uint64_t n_pairs, uint32_t ppb, uint32_t N,
            const uint32_t* __restrict__ spill, uint32_t low_bits, uint32_t n_buckets,
            unsigned long long* __restrict__ g_hist_or_cursor, unsigned long long* __restrict__ counters,
            uint32_t* __restrict__ blk_hist, uint32_t* __restrict__ keys) {
    extern __shared__ unsigned long long s_mem[];
    unsigned long long* s_base = s_mem;                                   // [n_buckets] (EMIT)
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_mem + n_buckets);    // [n_buckets]
    uint32_t* s_stage = s_hist + n_buckets;                               // [PAIR_WARPS][64][STAGE_STRIDE]
    __shared__ uint32_t s_woff[PAIR_WARPS][32], s_wn[PAIR_WARPS][32];
    __shared__ unsigned long long s_cnt[4];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (uint32_t b = threadIdx.x; b < n_buckets; b += PAIR_THREADS) {
        if (EMIT) {
            const uint32_t c = __ldg(blk_hist + (uint64_t)blockIdx.x * n_buckets + b);
            if (c) s_base[b] = atomicAdd(&g_hist_or_cursor[b], (unsigned long long)c);
        }
        s_hist[b] = 0;
    }
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t lo = (uint64_t)blockIdx.x * ppb;
    const uint64_t hi = lo + ppb < n_pairs ? lo + ppb : n_pairs;
    uint32_t* st = s_stage + wib * 64 * STAGE_STRIDE;
    const uint32_t NN = N * N;
    uint32_t c_used = 0, c_n = 0, c_short = 0;
    unsigned long long c_keys = 0;
    for (uint64_t p0 = lo + wib * 32; p0 < hi; p0 += PAIR_THREADS) {
        __syncwarp();                                                 // the previous round's reads are done
        stage_slots(f, r, p0, hi, st, lane);
        uint32_t ln = 0, rn = 0;
        if (p0 + lane < hi) {
            const uint32_t hf = st[lane * STAGE_STRIDE], hr = st[(32 + lane) * STAGE_STRIDE];
            const uint32_t cls = pair_class(hf, hr);
            if (cls == 0) { ln = hf >> 8; rn = hr >> 8; c_used++; }
            else if (cls == 1) c_n++;
            else c_short++;
        }
        const uint32_t m = tri(ln) + tri(rn) + ln * rn;
        c_keys += m;
        uint32_t inc = m;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        const uint32_t M = __shfl_sync(0xFFFFFFFFu, inc, 31);
        s_woff[wib][lane] = inc - m;
        s_wn[wib][lane] = ln | (rn << 16);
        __syncwarp();
        for (uint32_t t = lane; t < M; t += 32) {
            uint32_t own = 0;                                         // largest pair with prefix <= t
#pragma unroll
            for (uint32_t step = 16; step; step >>= 1)
                if (s_woff[wib][own + step] <= t) own += step;
            uint32_t u = t - s_woff[wib][own];
            const uint32_t nn = s_wn[wib][own];
            const uint32_t oln = nn & 0xFFFF, orn = nn >> 16;
            const uint32_t* pf = st + own * STAGE_STRIDE;
            const uint32_t* pr = st + (32 + own) * STAGE_STRIDE;
            const uint32_t* Ll = oln <= (uint32_t)SLOT_IDS ? pf + 1 : spill + pf[1];
            const uint32_t* Rl = orn <= (uint32_t)SLOT_IDS ? pr + 1 : spill + pr[1];
            const uint32_t m1 = tri(oln), m2 = tri(orn);
            uint32_t key;
            if (u < m1 + m2) {                                        // short_mat: a <= b inside one mate's list
                const bool second = u >= m1;
                const uint32_t* A = second ? Rl : Ll;
                const uint32_t n = second ? orn : oln;
                uint32_t v = second ? u - m1 : u, a = 0;
                while (v >= n - a) { v -= n - a; a++; }
                key = NN + A[a] * N + A[a + v];
            } else {                                                  // node_mat: every left node x every right node
                const uint32_t v = u - m1 - m2, a = v / orn;
                key = Ll[a] * N + Rl[v - a * orn];
            }
            const uint32_t b = key >> low_bits;
            if (EMIT) {
                const uint32_t o = atomicAdd(&s_hist[b], 1u);
                keys[s_base[b] + o] = key;
            } else {
                atomicAdd(&s_hist[b], 1u);
            }
        }
    }
    if (EMIT) return;
    if (c_used) atomicAdd(&s_cnt[0], (unsigned long long)c_used);
    if (c_n) atomicAdd(&s_cnt[1], (unsigned long long)c_n);
    if (c_short) atomicAdd(&s_cnt[2], (unsigned long long)c_short);
    if (c_keys) atomicAdd(&s_cnt[3], c_keys);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < n_buckets; b += PAIR_THREADS) {
        const uint32_t c = s_hist[b];
        if (c) atomicAdd(&g_hist_or_cursor[b], (unsigned long long)c);
        blk_hist[(uint64_t)blockIdx.x * n_buckets + b] = c;           // the emit pass reuses it
    }
    if (threadIdx.x == 0) {
        if (s_cnt[0]) atomicAdd(&counters[CNT_USED], s_cnt[0]);
        if (s_cnt[1]) atomicAdd(&counters[CNT_N], s_cnt[1]);
        if (s_cnt[2]) atomicAdd(&counters[CNT_SHORT], s_cnt[2]);
        if (s_cnt[3]) atomicAdd(&counters[CNT_KEYS], s_cnt[3]);
    }
}

// One work item = up to HIST_SLICE keys of one bucket (item_start from k_bucket_scan): counting
// sort of the low digit in shared memory; the per-value counts ARE the run lengths.  A bucket that
// is a single item owns its matrix cells (plain stores); the items of a hot bucket add their partial
// run lengths with integer atomics -- one per touched cell and item, the sum is order independent.
__global__ void __launch_bounds__(256)
k_bucket_hist(const uint32_t* __restrict__ keys, const unsigned long long* __restrict__ start,
              const unsigned long long* __restrict__ item_start, uint32_t n_buckets, uint32_t low_bits,
              uint64_t n_cells, uint64_t* __restrict__ mats) {
    extern __shared__ uint32_t s_bins[];
    __shared__ uint32_t s_bucket;
    const unsigned long long item = blockIdx.x;
    if (item >= item_start[n_buckets]) return;
    if (threadIdx.x == 0) {
        uint32_t lo = 0, hi = n_buckets;                              // largest b with item_start[b] <= item
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (item_start[mid] <= item) lo = mid; else hi = mid;
        }
        s_bucket = lo;
    }
    const uint32_t nb = 1u << low_bits, mask = nb - 1;
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) s_bins[i] = 0;
    __syncthreads();
    const uint32_t bucket = s_bucket;
    const unsigned long long s0 = start[bucket], e0 = start[bucket + 1];
    const unsigned long long n_items = item_start[bucket + 1] - item_start[bucket];
    const unsigned long long s = s0 + (item - item_start[bucket]) * HIST_SLICE;
    const unsigned long long e = s + HIST_SLICE < e0 ? s + HIST_SLICE : e0;
    const uint32_t lane = threadIdx.x & 31;
    // four independent loads per thread and step; lanes holding the same key add once
    for (unsigned long long i0 = s; i0 < e; i0 += 4ull * blockDim.x) {
        uint32_t k[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned long long i = i0 + (unsigned long long)u * blockDim.x + threadIdx.x;
            k[u] = i < e ? (__ldg(keys + i) & mask) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, k[u]);
            if (k[u] != 0xFFFFFFFFu && lane == (uint32_t)(__ffs((int)peers) - 1)) atomicAdd(&s_bins[k[u]], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    const uint64_t base = (uint64_t)bucket << low_bits;
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
        const uint32_t c = s_bins[i];
        if (c && base + i < n_cells) {
            if (n_items == 1) mats[base + i] += c;
            else atomicAdd(reinterpret_cast<unsigned long long*>(mats + base + i), (unsigned long long)c);
        }
    }
}

int count_pairs(Ctx* c, const ReadSlot* d_f, const ReadSlot* d_r, uint64_t total) {
    const uint64_t N = c->index.n_nodes;
    c->stats.total_pairs += total;
    if (total == 0) return VSPE_OK;
    const uint64_t cells = 2 * N * N;
    if (cells >= (1ull << 32)) { set_error("dense count matrices need 2*N*N <= 2^32 (N=%llu); sparse mode not built yet", (unsigned long long)N); return VSPE_ERR_LIMIT; }
    uint32_t low_bits = (uint32_t)std::max<int64_t>(7, std::min<int64_t>(15, c->opt_count_low_bits));
    while (((cells + (1ull << low_bits) - 1) >> low_bits) > 2048 && low_bits < 15) low_bits++;
    uint64_t nbk = cells ? ((cells + (1ull << low_bits) - 1) >> low_bits) : 1;
    if (nbk == 0) nbk = 1;
    if (nbk > 8192) { low_bits = 15; nbk = (cells + (1ull << 15) - 1) >> 15; }
    if (nbk > 8192) { set_error("graph too large for dense counting (N=%llu)", (unsigned long long)N); return VSPE_ERR_LIMIT; }
    const uint32_t n_buckets = (uint32_t)nbk;
    VSPE_TRY(c->bucket.reserve(4ull * (n_buckets + 1)));
    unsigned long long* g_hist = c->bucket.p;
    unsigned long long* g_start = g_hist + (n_buckets + 1);
    unsigned long long* g_cursor = g_start + (n_buckets + 1);
    unsigned long long* g_item = g_cursor + (n_buckets + 1);
    const uint32_t smem_count = n_buckets * 4 + STAGE_WORDS * 4, smem_emit = n_buckets * 12 + STAGE_WORDS * 4;
    if (!c->count_attr_set) {
        VSPE_CUDA(cudaFuncSetAttribute(k_bucket_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << 15) * 4));
        VSPE_CUDA(cudaFuncSetAttribute(k_pair_count, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 4 + STAGE_WORDS * 4));
        VSPE_CUDA(cudaFuncSetAttribute(k_pair_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12 + STAGE_WORDS * 4));
        VSPE_CUDA(cudaFuncSetAttribute(k_pair_flat<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12 + STAGE_WORDS * 4));
        VSPE_CUDA(cudaFuncSetAttribute(k_pair_flat<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12 + STAGE_WORDS * 4));
        c->count_attr_set = true;
    }
    // blocks that are all resident at once (the emit kernel needs more shared memory: it decides)
    int per_sm = 1;
    VSPE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pair_emit, PAIR_THREADS, smem_emit));
    const uint64_t target_blocks = (uint64_t)c->sm_count * (uint64_t)std::max(1, std::min(per_sm, 4));
    const uint64_t BATCH = 4ull << 20;     // pairs per key batch
    cudaStream_t st = c->stream;
    for (uint64_t off = 0; off < total; off += BATCH) {
        uint64_t n = total - off < BATCH ? total - off : BATCH;
        // pairs per block: a multiple of the block size, so that about target_blocks blocks cover the batch
        const uint32_t ppb = (uint32_t)(((n + target_blocks - 1) / target_blocks + PAIR_THREADS - 1) / PAIR_THREADS * PAIR_THREADS);
        const uint32_t grid = (uint32_t)((n + ppb - 1) / ppb);
        unsigned long long h_keys = 0, h_err = 0;
        VSPE_CUDA(cudaMemsetAsync(g_hist, 0, (n_buckets + 1) * 8, st));
        // the per-block histograms are kept for the emit kernel
        VSPE_TRY(c->blk_hist.reserve((uint64_t)grid * n_buckets));
        uint32_t* blk_hist = c->blk_hist.p;
        const bool flat = c->opt_count_flat != 0;
        if (flat)
            k_pair_flat<false><<<grid, PAIR_THREADS, smem_emit, st>>>(d_f + off, d_r + off, n, ppb, (uint32_t)N, c->spill.p, low_bits, n_buckets,
                                                                      g_hist, c->counters.p, blk_hist, nullptr);
        else
            k_pair_count<<<grid, PAIR_THREADS, smem_count, st>>>(d_f + off, d_r + off, n, ppb, N, c->spill.p, low_bits, n_buckets,
                                                                 g_hist, c->counters.p, blk_hist);
        VSPE_LAUNCH_CHECK(c);
        // one D2H + sync per batch: the cumulative key counter and the kernels' error flags
        VSPE_CUDA(cudaMemcpyAsync(&h_keys, c->counters.p + CNT_KEYS, 8, cudaMemcpyDeviceToHost, st));
        VSPE_CUDA(cudaMemcpyAsync(&h_err, c->counters.p + CNT_ERR, 8, cudaMemcpyDeviceToHost, st));
        VSPE_CUDA(cudaStreamSynchronize(st));
        VSPE_TRY(adapt_map_variant(c));
        c->last_err_flags = h_err;                             // every scan / map kernel of this call ran before
        c->err_flags_fresh = true;
        uint64_t n_keys = h_keys - c->keys_seen;
        c->keys_seen = h_keys;
        if (n_keys == 0) continue;
        if (n_keys > 0xFFFFFFF0ull) { set_error("key batch too large"); return VSPE_ERR_LIMIT; }
        VSPE_TRY(c->keys.reserve(n_keys));
        k_bucket_scan<<<1, 1024, 0, st>>>(g_hist, n_buckets, g_start, g_cursor, g_item);
        VSPE_LAUNCH_CHECK(c);
        if (flat)
            k_pair_flat<true><<<grid, PAIR_THREADS, smem_emit, st>>>(d_f + off, d_r + off, n, ppb, (uint32_t)N, c->spill.p, low_bits, n_buckets,
                                                                     g_cursor, c->counters.p, blk_hist, c->keys.p);
        else
            k_pair_emit<<<grid, PAIR_THREADS, smem_emit, st>>>(d_f + off, d_r + off, n, ppb, N, c->spill.p, low_bits, n_buckets,
                                                               g_cursor, c->keys.p, blk_hist);
        VSPE_LAUNCH_CHECK(c);
        // work items: every non-empty bucket rounds up to whole HIST_SLICE-key items
        const uint32_t n_items = (uint32_t)std::min<uint64_t>(n_keys / HIST_SLICE + n_buckets, 0x7FFFFFFFull);
        k_bucket_hist<<<n_items, 256, (1u << low_bits) * 4, st>>>(c->keys.p, g_start, g_item, n_buckets, low_bits, cells, c->mats.p);
        VSPE_LAUNCH_CHECK(c);
    }
    return VSPE_OK;
}

}  // namespace vspe
