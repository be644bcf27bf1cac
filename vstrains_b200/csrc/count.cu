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
#include "ctx.cuh"

namespace vspe {

static constexpr int PAIR_THREADS = 256;

struct ListRef {
    const uint32_t* ids;
    uint32_t n;
};

__device__ __forceinline__ ListRef list_of(const ReadSlot* s, const uint32_t* __restrict__ spill) {
    ListRef r;
    r.n = s->hdr >> 8;
    r.ids = r.n <= SLOT_IDS ? s->ids : spill + s->ids[0];
    return r;
}

template <class F>
__device__ __forceinline__ void for_each_key(const ListRef& l, const ListRef& r, uint64_t N, F fn) {
    const uint64_t NN = N * N;
    for (uint32_t a = 0; a < l.n; a++) {
        uint64_t ia = l.ids[a];
        for (uint32_t b = a; b < l.n; b++) fn(NN + ia * N + l.ids[b]);
    }
    for (uint32_t a = 0; a < r.n; a++) {
        uint64_t ia = r.ids[a];
        for (uint32_t b = a; b < r.n; b++) fn(NN + ia * N + r.ids[b]);
    }
    for (uint32_t a = 0; a < l.n; a++) {
        uint64_t ia = l.ids[a];
        for (uint32_t b = 0; b < r.n; b++) fn(ia * N + r.ids[b]);
    }
}

// 0 used, 1 N-pair, 2 short pair  (N takes precedence: PE_Inference.py:160 before :162)
__device__ __forceinline__ uint32_t pair_class(uint32_t hf, uint32_t hr) {
    uint32_t sf = hf & 0xFF, sr = hr & 0xFF;
    if (sf == ST_N || sr == ST_N) return 1;
    if (sf == ST_SHORT || sr == ST_SHORT) return 2;
    return 0;
}

__global__ void __launch_bounds__(PAIR_THREADS)
k_pair_count(const ReadSlot* __restrict__ f, const ReadSlot* __restrict__ r, uint64_t n_pairs, uint64_t N,
             const uint32_t* __restrict__ spill, uint32_t low_bits, uint32_t n_buckets,
             unsigned long long* __restrict__ g_hist, unsigned long long* __restrict__ counters,
             uint32_t* __restrict__ blk_hist) {
    extern __shared__ uint32_t s_hist[];
    __shared__ unsigned long long s_cnt[4];
    for (uint32_t b = threadIdx.x; b < n_buckets; b += PAIR_THREADS) s_hist[b] = 0;
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    uint64_t p = (uint64_t)blockIdx.x * PAIR_THREADS + threadIdx.x;
    if (p < n_pairs) {
        uint32_t cls = pair_class(f[p].hdr, r[p].hdr);
        if (cls == 0) {
            ListRef l = list_of(f + p, spill), rr = list_of(r + p, spill);
            uint32_t m = 0;
            for_each_key(l, rr, N, [&](uint64_t key) { atomicAdd(&s_hist[key >> low_bits], 1u); m++; });
            atomicAdd(&s_cnt[0], 1ull);
            atomicAdd(&s_cnt[3], (unsigned long long)m);
        } else {
            atomicAdd(&s_cnt[cls], 1ull);
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < n_buckets; b += PAIR_THREADS) {
        uint32_t c = s_hist[b];
        if (c) atomicAdd(&g_hist[b], (unsigned long long)c);
        if (blk_hist) blk_hist[(uint64_t)blockIdx.x * n_buckets + b] = c;     // k_pair_emit reuses it
    }
    if (threadIdx.x == 0) {
        if (s_cnt[0]) atomicAdd(&counters[CNT_USED], s_cnt[0]);
        if (s_cnt[1]) atomicAdd(&counters[CNT_N], s_cnt[1]);
        if (s_cnt[2]) atomicAdd(&counters[CNT_SHORT], s_cnt[2]);
        if (s_cnt[3]) atomicAdd(&counters[CNT_KEYS], s_cnt[3]);
    }
}

// exclusive scan of the bucket histogram (n <= 8192) -> start[0..n], cursor = start
__global__ void __launch_bounds__(1024)
k_bucket_scan(const unsigned long long* __restrict__ hist, uint32_t n, unsigned long long* __restrict__ start,
              unsigned long long* __restrict__ cursor) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t off = 0; off < n; off += 1024) {
        uint32_t i = off + threadIdx.x;
        unsigned long long x = i < n ? hist[i] : 0, inc = x;
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += y;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        unsigned long long wbase = 0, tot = 0;
        for (uint32_t w = 0; w < 32; w++) {
            unsigned long long t = s_warp[w];
            if (w < wid) wbase += t;
            tot += t;
        }
        unsigned long long carry = s_carry;
        if (i < n) { start[i] = carry + wbase + inc - x; cursor[i] = carry + wbase + inc - x; }
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[n] = s_carry;
}

__global__ void __launch_bounds__(PAIR_THREADS)
k_pair_emit(const ReadSlot* __restrict__ f, const ReadSlot* __restrict__ r, uint64_t n_pairs, uint64_t N,
            const uint32_t* __restrict__ spill, uint32_t low_bits, uint32_t n_buckets,
            unsigned long long* __restrict__ g_cursor, uint32_t* __restrict__ keys, const uint32_t* __restrict__ blk_hist) {
    extern __shared__ unsigned long long s_mem[];
    unsigned long long* s_base = s_mem;                              // [n_buckets]
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_mem + n_buckets);   // [n_buckets]
    // this block's keys per bucket: stored by k_pair_count (same block -> pairs mapping), or recounted
    for (uint32_t b = threadIdx.x; b < n_buckets; b += PAIR_THREADS)
        s_hist[b] = blk_hist ? __ldg(blk_hist + (uint64_t)blockIdx.x * n_buckets + b) : 0u;
    __syncthreads();
    uint64_t p = (uint64_t)blockIdx.x * PAIR_THREADS + threadIdx.x;
    bool used = p < n_pairs && pair_class(f[p].hdr, r[p].hdr) == 0;
    ListRef l = {nullptr, 0}, rr = {nullptr, 0};
    if (used) {
        l = list_of(f + p, spill);
        rr = list_of(r + p, spill);
        if (!blk_hist) for_each_key(l, rr, N, [&](uint64_t key) { atomicAdd(&s_hist[key >> low_bits], 1u); });
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < n_buckets; b += PAIR_THREADS) {
        uint32_t c = s_hist[b];
        if (c) s_base[b] = atomicAdd(&g_cursor[b], (unsigned long long)c);
        s_hist[b] = 0;
    }
    __syncthreads();
    if (used) {
        for_each_key(l, rr, N, [&](uint64_t key) {
            uint32_t b = (uint32_t)(key >> low_bits);
            uint32_t o = atomicAdd(&s_hist[b], 1u);
            keys[s_base[b] + o] = (uint32_t)key;
        });
    }
}

__global__ void __launch_bounds__(256)
k_bucket_hist(const uint32_t* __restrict__ keys, const unsigned long long* __restrict__ start, uint32_t low_bits,
              uint64_t n_cells, uint64_t* __restrict__ mats) {
    extern __shared__ uint32_t s_bins[];
    const uint32_t nb = 1u << low_bits, mask = nb - 1;
    const unsigned long long s = start[blockIdx.x], e = start[blockIdx.x + 1];
    if (s == e) return;
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) s_bins[i] = 0;
    __syncthreads();
    // four independent loads per thread and step: the loop is latency-bound otherwise
    for (unsigned long long i0 = s; i0 < e; i0 += 4ull * blockDim.x) {
        uint32_t k[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned long long i = i0 + (unsigned long long)u * blockDim.x + threadIdx.x;
            ok[u] = i < e;
            k[u] = ok[u] ? __ldg(keys + i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) if (ok[u]) atomicAdd(&s_bins[k[u] & mask], 1u);
    }
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x << low_bits;
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
        uint32_t c = s_bins[i];
        if (c && base + i < n_cells) mats[base + i] += c;
    }
}

int count_pairs(Ctx* c, const ReadSlot* d_f, const ReadSlot* d_r, uint64_t total) {
    const uint64_t N = c->index.n_nodes;
    c->stats.total_pairs += total;
    if (total == 0) return VSPE_OK;
    const uint64_t cells = 2 * N * N;
    if (cells > (1ull << 32)) { set_error("dense count matrices need 2*N*N <= 2^32 (N=%llu); sparse mode not built yet", (unsigned long long)N); return VSPE_ERR_LIMIT; }
    uint32_t low_bits = 7;
    while (((cells + (1ull << low_bits) - 1) >> low_bits) > 2048 && low_bits < 15) low_bits++;
    uint64_t nbk = cells ? ((cells + (1ull << low_bits) - 1) >> low_bits) : 1;
    if (nbk == 0) nbk = 1;
    if (nbk > 8192) { low_bits = 15; nbk = (cells + (1ull << 15) - 1) >> 15; }
    if (nbk > 8192) { set_error("graph too large for dense counting (N=%llu)", (unsigned long long)N); return VSPE_ERR_LIMIT; }
    const uint32_t n_buckets = (uint32_t)nbk;
    VSPE_TRY(c->bucket.reserve(3ull * (n_buckets + 1)));
    unsigned long long* g_hist = c->bucket.p;
    unsigned long long* g_start = g_hist + (n_buckets + 1);
    unsigned long long* g_cursor = g_start + (n_buckets + 1);
    if (!c->count_attr_set) {
        VSPE_CUDA(cudaFuncSetAttribute(k_bucket_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << 15) * 4));
        VSPE_CUDA(cudaFuncSetAttribute(k_pair_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12));
        c->count_attr_set = true;
    }
    const uint64_t BATCH = 4ull << 20;     // pairs per key batch
    cudaStream_t st = c->stream;
    for (uint64_t off = 0; off < total; off += BATCH) {
        uint64_t n = total - off < BATCH ? total - off : BATCH;
        uint32_t grid = (uint32_t)((n + PAIR_THREADS - 1) / PAIR_THREADS);
        unsigned long long h_keys = 0, h_err = 0;
        VSPE_CUDA(cudaMemsetAsync(g_hist, 0, (n_buckets + 1) * 8, st));
        // per-block histograms are kept for the emit kernel when they fit a modest scratch
        uint32_t* blk_hist = nullptr;
        if ((uint64_t)grid * n_buckets * 4 <= (256ull << 20)) {
            VSPE_TRY(c->blk_hist.reserve((uint64_t)grid * n_buckets));
            blk_hist = c->blk_hist.p;
        }
        k_pair_count<<<grid, PAIR_THREADS, n_buckets * 4, st>>>(d_f + off, d_r + off, n, N, c->spill.p, low_bits, n_buckets,
                                                                g_hist, c->counters.p, blk_hist);
        VSPE_LAUNCH_CHECK(c);
        // one D2H + sync per batch: the cumulative key counter and the kernels' error flags
        VSPE_CUDA(cudaMemcpyAsync(&h_keys, c->counters.p + CNT_KEYS, 8, cudaMemcpyDeviceToHost, st));
        VSPE_CUDA(cudaMemcpyAsync(&h_err, c->counters.p + CNT_ERR, 8, cudaMemcpyDeviceToHost, st));
        VSPE_CUDA(cudaStreamSynchronize(st));
        c->last_err_flags = h_err;                             // every scan / map kernel of this call ran before
        c->err_flags_fresh = true;
        uint64_t n_keys = h_keys - c->keys_seen;
        c->keys_seen = h_keys;
        if (n_keys == 0) continue;
        if (n_keys > 0xFFFFFFF0ull) { set_error("key batch too large"); return VSPE_ERR_LIMIT; }
        VSPE_TRY(c->keys.reserve(n_keys));
        k_bucket_scan<<<1, 1024, 0, st>>>(g_hist, n_buckets, g_start, g_cursor);
        VSPE_LAUNCH_CHECK(c);
        k_pair_emit<<<grid, PAIR_THREADS, n_buckets * 12, st>>>(d_f + off, d_r + off, n, N, c->spill.p, low_bits, n_buckets,
                                                                g_cursor, c->keys.p, blk_hist);
        VSPE_LAUNCH_CHECK(c);
        k_bucket_hist<<<n_buckets, 256, (1u << low_bits) * 4, st>>>(c->keys.p, g_start, low_bits, cells, c->mats.p);
        VSPE_LAUNCH_CHECK(c);
    }
    return VSPE_OK;
}

}  // namespace vspe
