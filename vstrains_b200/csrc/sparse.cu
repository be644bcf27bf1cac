// sparse.cu -- K6: radix sort + run-length reduce of weighted (matrix, i, j) keys.
//
// The last step of the accumulation of reference utils/VStrains_PE_Inference.py:174-188 for every
// graph size: key = mat*N*N + i*N + j (64 bit) with a weight (link.cu expands each distinct
// combination of node lists once, weighted by its multiplicity).  The north-star recipe:
//   stable LSD radix sort (8-bit digits, only the digits the key range needs)  ->  run-length
//   reduce (head flags + prefix sums of the weights)  ->  runs (key, count), keys ascending.
// Dense mode adds the runs to the N x N matrices; sparse mode (graphs whose matrices cannot exist,
// e.g. the 200 000-node stress config: 2 * N^2 * 8 B = 640 GB) keeps them as the context's sorted
// run list and merges new batches / other ranks' runs by one more sort + reduce.  No atomics on
// the counts; integer sums only, so the result does not depend on batch or rank boundaries.
#include "link.cuh"

namespace vspe {

static constexpr int SC_THREADS = 256, SC_ITEMS = 8, SC_TILE = SC_THREADS * SC_ITEMS;   // scan tile

// ---------------------------------------------------------------------------------------------
// device-wide exclusive scan (three kernels: tile sums, scan of the sums, tile scan + base)
// ---------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(SC_THREADS)
k_tile_sums(const T* __restrict__ in, uint64_t n, T* __restrict__ sums) {
    __shared__ T s_w[SC_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SC_TILE;
    T acc = 0;
    for (int k = 0; k < SC_ITEMS; k++) {
        const uint64_t i = base + (uint64_t)k * SC_THREADS + threadIdx.x;
        if (i < n) acc += in[i];
    }
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = 0;
        for (int w = 0; w < SC_THREADS / 32; w++) t += s_w[w];
        sums[blockIdx.x] = t;
    }
}

template <class T>
__global__ void __launch_bounds__(1024)
k_scan_sums(T* __restrict__ sums, uint64_t n, T* __restrict__ total) {       // in place, exclusive
    __shared__ T s_w[32];
    __shared__ T s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint64_t off = 0; off < n; off += 1024) {
        const uint64_t i = off + threadIdx.x;
        const T x = i < n ? sums[i] : (T)0;
        T inc = x;
        for (int d = 1; d < 32; d <<= 1) {
            const T y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        T wb = 0, tot = 0;
        for (uint32_t w = 0; w < 32; w++) {
            const T t = s_w[w];
            if (w < wid) wb += t;
            tot += t;
        }
        const T carry = s_carry;
        if (i < n) sums[i] = carry + wb + inc - x;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = s_carry;
}

// out[i] = sums[tile] + exclusive prefix inside the tile (blocked order: item k of thread t is
// element base + t*SC_ITEMS + k, so the order is the array order)
template <class T>
__global__ void __launch_bounds__(SC_THREADS)
k_tile_scan(const T* __restrict__ in, uint64_t n, const T* __restrict__ sums, T* __restrict__ out) {
    __shared__ T s_w[SC_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SC_TILE + (uint64_t)threadIdx.x * SC_ITEMS;
    T v[SC_ITEMS], acc = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        v[k] = base + k < n ? in[base + k] : (T)0;
        acc += v[k];
    }
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T inc = acc;
    for (int d = 1; d < 32; d <<= 1) {
        const T y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += y;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    T wb = 0;
    for (uint32_t w = 0; w < wid; w++) wb += s_w[w];
    T run = sums[blockIdx.x] + wb + inc - acc;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

template <class T>
static int device_exclusive_scan(Ctx* c, const T* in, T* out, uint64_t n, T* d_sums /* >= tiles + 1 */, T* d_total) {
    if (n == 0) {
        if (d_total) VSPE_CUDA(cudaMemsetAsync(d_total, 0, sizeof(T), c->stream));
        return VSPE_OK;
    }
    const uint64_t tiles = (n + SC_TILE - 1) / SC_TILE;
    k_tile_sums<T><<<(uint32_t)tiles, SC_THREADS, 0, c->stream>>>(in, n, d_sums);
    VSPE_LAUNCH_CHECK(c);
    k_scan_sums<T><<<1, 1024, 0, c->stream>>>(d_sums, tiles, d_total);
    VSPE_LAUNCH_CHECK(c);
    k_tile_scan<T><<<(uint32_t)tiles, SC_THREADS, 0, c->stream>>>(in, n, d_sums, out);
    VSPE_LAUNCH_CHECK(c);
    return VSPE_OK;
}

int device_scan_u64(Ctx* c, const unsigned long long* in, unsigned long long* out, uint64_t n, unsigned long long* sums,
                    unsigned long long* total) {
    return device_exclusive_scan<unsigned long long>(c, in, out, n, sums, total);
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort of (key, value) pairs, 8-bit digits.  One warp per 4096-element tile:
//   k_radix_hist     per-tile digit histogram -> hist[digit][tile]
//   (exclusive scan over hist in digit-major order = global offset of every (digit, tile))
//   k_radix_scatter  the warp walks its tile 32 elements at a time, in order; equal digits inside
//                    a step are ranked with __match_any_sync, the running per-digit offsets live in
//                    shared memory, so equal keys keep their input order (stability).
// ---------------------------------------------------------------------------------------------
static constexpr int RS_TILE = 4096;

__global__ void __launch_bounds__(32)
k_radix_hist(const unsigned long long* __restrict__ keys, uint64_t n, uint32_t shift, uint32_t n_tiles, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_h[256];
    const uint32_t lane = threadIdx.x, tile = blockIdx.x;
    for (uint32_t d = lane; d < 256; d += 32) s_h[d] = 0;
    __syncwarp();
    const uint64_t base = (uint64_t)tile * RS_TILE;
    for (uint32_t k = 0; k < RS_TILE; k += 32) {
        const uint64_t i = base + k + lane;
        if (i < n) atomicAdd(&s_h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncwarp();
    for (uint32_t d = lane; d < 256; d += 32) hist[(uint64_t)d * n_tiles + tile] = s_h[d];
}

__global__ void __launch_bounds__(32)
k_radix_scatter(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ vals, uint64_t n, uint32_t shift,
                uint32_t n_tiles, const uint32_t* __restrict__ offs, unsigned long long* __restrict__ keys_out,
                unsigned long long* __restrict__ vals_out) {
    __shared__ uint32_t s_o[256];
    const uint32_t lane = threadIdx.x, tile = blockIdx.x;
    for (uint32_t d = lane; d < 256; d += 32) s_o[d] = offs[(uint64_t)d * n_tiles + tile];
    __syncwarp();
    const uint64_t base = (uint64_t)tile * RS_TILE;
    const uint32_t lt = (1u << lane) - 1;
    for (uint32_t k = 0; k < RS_TILE; k += 32) {
        const uint64_t i = base + k + lane;
        if (base + k >= n) break;
        const bool ok = i < n;
        const unsigned long long key = ok ? keys[i] : 0ull, val = ok ? vals[i] : 0ull;
        const uint32_t d = ok ? ((uint32_t)(key >> shift) & 255u) : 256u + lane;   // inactive lanes match nobody
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        const uint32_t rank = __popc(peers & lt);
        uint32_t pos = 0;
        if (ok) pos = s_o[d] + rank;
        __syncwarp();
        if (ok && rank == (uint32_t)__popc(peers) - 1) s_o[d] = pos + 1;            // last peer advances the digit's cursor
        __syncwarp();
        if (ok) { keys_out[pos] = key; vals_out[pos] = val; }
    }
}

// ---------------------------------------------------------------------------------------------
// run-length reduce of a sorted (key, value) list
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_rle_flags(const unsigned long long* __restrict__ keys, uint64_t n, uint32_t* __restrict__ head) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// run r starts at the element whose head flag is set and whose exclusive flag sum is r
__global__ void __launch_bounds__(256)
k_rle_heads(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ head, const uint32_t* __restrict__ head_excl,
            const unsigned long long* __restrict__ val_excl, uint64_t n, unsigned long long* __restrict__ run_keys,
            unsigned long long* __restrict__ run_start) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && head[i]) {
        run_keys[head_excl[i]] = keys[i];
        run_start[head_excl[i]] = val_excl[i];
    }
}

__global__ void __launch_bounds__(256)
k_rle_counts(const unsigned long long* __restrict__ run_start, uint64_t n_runs, const unsigned long long* __restrict__ total,
             unsigned long long* __restrict__ run_count) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_runs) run_count[r] = (r + 1 < n_runs ? run_start[r + 1] : *total) - run_start[r];
}

// sort (keys, vals)[0..n) by key and reduce equal keys by summing their values.
// In / out in sp.k[0], sp.v[0]; the other buffers are scratch.  Returns the number of runs.
static uint32_t bits_for(uint64_t cells) {
    uint32_t b = 1;
    while (b < 64 && (1ull << b) < cells) b++;
    return b;
}

int sparse_sort_reduce(Ctx* c, uint64_t n, uint64_t* n_runs_out) {
    Sparse& sp = c->sparse;
    const uint32_t key_bits = bits_for(2ull * c->index.n_nodes * c->index.n_nodes);
    *n_runs_out = 0;
    if (n == 0) return VSPE_OK;
    if (n > 0xFFFFFFF0ull) { set_error("sparse batch too large"); return VSPE_ERR_LIMIT; }
    const uint32_t n_tiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    const uint64_t hist_n = 256ull * n_tiles;
    VSPE_TRY(sp.hist.reserve(hist_n + 2));
    VSPE_TRY(sp.sums32.reserve(hist_n / SC_TILE + n / SC_TILE + 8));
    VSPE_TRY(sp.sums64.reserve(n / SC_TILE + 8));
    VSPE_TRY(sp.head.reserve(2 * n + 4));
    VSPE_TRY(sp.vscan.reserve(n + 4));
    int cur = 0;
    for (uint32_t shift = 0; shift < key_bits; shift += 8) {
        k_radix_hist<<<n_tiles, 32, 0, c->stream>>>(sp.k[cur].p, n, shift, n_tiles, sp.hist.p);
        VSPE_LAUNCH_CHECK(c);
        VSPE_TRY(device_exclusive_scan<uint32_t>(c, sp.hist.p, sp.hist.p, hist_n, sp.sums32.p, nullptr));
        k_radix_scatter<<<n_tiles, 32, 0, c->stream>>>(sp.k[cur].p, sp.v[cur].p, n, shift, n_tiles, sp.hist.p, sp.k[cur ^ 1].p, sp.v[cur ^ 1].p);
        VSPE_LAUNCH_CHECK(c);
        cur ^= 1;
    }
    // run-length reduce: head flags, their prefix sum (run index), prefix sum of the values
    uint32_t* head = sp.head.p;
    uint32_t* head_excl = sp.head.p + n + 2;
    const uint32_t g = (uint32_t)((n + 255) / 256);
    k_rle_flags<<<g, 256, 0, c->stream>>>(sp.k[cur].p, n, head);
    VSPE_LAUNCH_CHECK(c);
    unsigned long long* d_tot = sp.totals.p;                  // [0] runs (as u32 in the low word), [1] value total
    VSPE_TRY(device_exclusive_scan<uint32_t>(c, head, head_excl, n, sp.sums32.p, reinterpret_cast<uint32_t*>(d_tot)));
    VSPE_TRY(device_exclusive_scan<unsigned long long>(c, sp.v[cur].p, sp.vscan.p, n, sp.sums64.p, d_tot + 1));
    unsigned long long h_tot[2] = {0, 0};
    VSPE_CUDA(cudaMemcpyAsync(h_tot, d_tot, 16, cudaMemcpyDeviceToHost, c->stream));
    VSPE_CUDA(cudaStreamSynchronize(c->stream));
    const uint64_t n_runs = (uint32_t)h_tot[0];
    // run keys -> k[cur^1], run starts -> v[cur^1], then counts back into v[cur]... keep the result in slot 0
    k_rle_heads<<<g, 256, 0, c->stream>>>(sp.k[cur].p, head, head_excl, sp.vscan.p, n, sp.k[cur ^ 1].p, sp.v[cur ^ 1].p);
    VSPE_LAUNCH_CHECK(c);
    k_rle_counts<<<(uint32_t)((n_runs + 255) / 256), 256, 0, c->stream>>>(sp.v[cur ^ 1].p, n_runs, d_tot + 1, sp.v[cur].p);
    VSPE_LAUNCH_CHECK(c);
    // result: keys in k[cur^1], counts in v[cur]; normalise to k[0], v[0]
    if ((cur ^ 1) != 0) VSPE_CUDA(cudaMemcpyAsync(sp.k[0].p, sp.k[cur ^ 1].p, n_runs * 8, cudaMemcpyDeviceToDevice, c->stream));
    if (cur != 0) VSPE_CUDA(cudaMemcpyAsync(sp.v[0].p, sp.v[cur].p, n_runs * 8, cudaMemcpyDeviceToDevice, c->stream));
    *n_runs_out = n_runs;
    return VSPE_OK;
}

// make room for `extra` more elements after the first `keep` ones (both double buffers)
int sparse_reserve(Ctx* c, uint64_t total) {
    Sparse& sp = c->sparse;
    for (int b = 0; b < 2; b++) {
        VSPE_TRY(sp.k[b].reserve(total + 8, b == 0, c->stream));
        VSPE_TRY(sp.v[b].reserve(total + 8, b == 0, c->stream));
    }
    if (!sp.totals.p) VSPE_TRY(sp.totals.reserve(4));
    return VSPE_OK;
}

// add runs from another context / rank (host arrays): append + sort + reduce
int sparse_merge_host(Ctx* c, const uint64_t* keys, const uint64_t* counts, uint64_t n) {
    Sparse& sp = c->sparse;
    if (n == 0) return VSPE_OK;
    VSPE_TRY(sparse_reserve(c, sp.n_runs + n));
    VSPE_CUDA(cudaMemcpyAsync(sp.k[0].p + sp.n_runs, keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    VSPE_CUDA(cudaMemcpyAsync(sp.v[0].p + sp.n_runs, counts, n * 8, cudaMemcpyHostToDevice, c->stream));
    uint64_t runs = 0;
    VSPE_TRY(sparse_sort_reduce(c, sp.n_runs + n, &runs));
    sp.n_runs = runs;
    return VSPE_OK;
}

// add runs that already live on this device (e.g. gathered from the other ranks over NVLink)
int sparse_merge_device(Ctx* c, const uint64_t* d_keys, const uint64_t* d_counts, uint64_t n) {
    Sparse& sp = c->sparse;
    if (n == 0) return VSPE_OK;
    VSPE_TRY(sparse_reserve(c, sp.n_runs + n));
    VSPE_CUDA(cudaMemcpyAsync(sp.k[0].p + sp.n_runs, d_keys, n * 8, cudaMemcpyDeviceToDevice, c->stream));
    VSPE_CUDA(cudaMemcpyAsync(sp.v[0].p + sp.n_runs, d_counts, n * 8, cudaMemcpyDeviceToDevice, c->stream));
    uint64_t runs = 0;
    VSPE_TRY(sparse_sort_reduce(c, sp.n_runs + n, &runs));
    sp.n_runs = runs;
    return VSPE_OK;
}

}  // namespace vspe