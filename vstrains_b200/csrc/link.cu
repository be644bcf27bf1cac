// link.cu -- K5 + K6: link keys and their deterministic counting.
//
// Replaces the accumulation loops of reference utils/VStrains_PE_Inference.py:160-188:
//   N / short filtering (:160-163), short_mat[a][b] += 1 for a <= b within each mate's node
//   list (:174-184), node_mat[i][j] += 1 for i in lefts, j in rights (:186-188).
//
// Reads arrive as list handles (link.cuh).  Per batch of pairs:
//   k_pair_agg      classify the pair (N before short, :160-163) and count it under its
//                   (handle_left, handle_right) combination in the pair table -- integer adds, so
//                   the totals do not depend on the order the pairs arrive in;
//   k_comb_weigh    every distinct combination adds its multiplicity to the `used` weight of its two
//                   lists and reports how many node_mat keys it expands to;
//   k_list_weigh    every list with a non-zero weight reports its short_mat keys;
//   k_comb_emit     combination (L, R) x c  ->  keys  i*N + j            for i in L, j in R, weight c
//   k_list_emit     list A x used           ->  keys  N*N + min*N + max  for every unordered pair of
//                   A incl. the diagonal, weight used   (= short_mat[a][b], a <= b by list position:
//                   the reference's lists ascend by node index, :36-47, so position order is index order)
//   dense mode      k_wkey_hist / k_wkey_scan / k_wkey_scatter / k_bucket_count: MSD radix partition +
//                   per-bucket counting sort whose summed weights are the run totals, added to
//                   [node_mat | short_mat] by the one CTA that owns the cells (no atomics on them);
//   sparse mode     sparse_sort_reduce (sparse.cu): stable LSD radix sort of the weighted keys +
//                   run-length reduce, merged into the context's sorted run list.
// The expansion is exact because all pairs of a combination have identical node lists (compared id
// by id when the handle was given out), so they add the same keys.
#include <algorithm>

#include "link.cuh"

namespace vspe {

static constexpr uint64_t LINK_BATCH = 32ull << 20;         // pairs per batch (weights -- up to two per pair and list -- stay below 2^32)

LinkView link_view(Ctx* c) {
    LinkView v;
    v.recs = c->list_recs.p;
    v.T = c->list_T;
    v.t_mask = c->list_T - 1;
    v.ov_cap = c->list_ov_cap;
    v.max_lists = c->list_T / 2;
    v.occ = c->list_occ.p;
    v.spill = c->spill.p;
    v.spill_cap = c->spill.cap;
    v.counters = c->counters.p;
    return v;
}

int link_setup(Ctx* c) {
    // table size from the graph: distinct node lists are a small multiple of the node count
    uint64_t want = 64ull * std::max<uint64_t>(c->index.n_nodes, 1);
    uint32_t T = 1u << 18;
    while (T < want && T < (1u << 24)) T <<= 1;
    const uint32_t ov = std::max<uint32_t>(c->list_ov_cap, 1u << 16);
    if (T != c->list_T || !c->list_recs.p) {
        c->list_recs.release();
        VSPE_TRY(c->list_recs.reserve((uint64_t)T + ov));
        VSPE_TRY(c->list_occ.reserve(T));
        c->list_T = T;
        c->list_ov_cap = (uint32_t)std::min<uint64_t>(c->list_recs.cap - T, 0x7FFFFFFFu);
    }
    if (!c->spill.p) VSPE_TRY(c->spill.reserve(4u << 20));
    return link_reset(c);
}

int link_reset(Ctx* c) {
    if (!c->list_recs.p) return VSPE_OK;
    VSPE_CUDA(cudaMemsetAsync(c->list_recs.p, 0, (uint64_t)c->list_T * sizeof(ListRec), c->stream));
    // the read memo holds list handles: it is forgotten with the lists
    if (c->memo.p) VSPE_CUDA(cudaMemsetAsync(c->memo.p, 0, c->memo.cap * sizeof(uint32_t), c->stream));
    c->memo_off = false;
    c->memo_seen = 0;
    return VSPE_OK;                                         // (the counters are zeroed by vspe_reset)
}

// More private records / spill words after a launch ran out of them; the caller repeats the launch.
int link_grow_overflow(Ctx* c) {
    VSPE_CUDA(cudaStreamSynchronize(c->stream));
    const uint64_t ov = (uint64_t)c->list_ov_cap * 4;
    if ((uint64_t)c->list_T + ov > 0xFFFFFFF0ull) { set_error("node-list records exhausted (more than 2^32)"); return VSPE_ERR_LIMIT; }
    VSPE_TRY(c->list_recs.reserve((uint64_t)c->list_T + ov, true, c->stream));
    c->list_ov_cap = (uint32_t)std::min<uint64_t>(c->list_recs.cap - c->list_T, 0x7FFFFFFFu);
    VSPE_TRY(c->spill.reserve(c->spill.cap * 4, true, c->stream));
    return VSPE_OK;
}

// ---------------------------------------------------------------------------------------------
// ReadSlots (slot-writing map tiers) <-> handles
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_intern_slots(LinkView lv, const ReadSlot* __restrict__ slots, uint64_t n_arg, const uint32_t* __restrict__ scatter,
               const unsigned long long* __restrict__ n_dev, uint32_t* __restrict__ handles) {
    const uint64_t n = n_dev ? *n_dev : n_arg;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = scatter ? scatter[i] : i;                  // deferred reads: compact slot i belongs to read scatter[i]
        const ReadSlot* s = slots + i;
        const uint32_t hdr = s->hdr, st = hdr & 0xFF, cnt = hdr >> 8;
        uint32_t h;
        if (st == ST_N) h = H_N;
        else if (st == ST_SHORT) h = H_SHORT;
        else if (cnt <= (uint32_t)SLOT_IDS) h = intern_list(lv, cnt, s->ids, 1);
        else h = intern_list(lv, cnt, lv.spill + s->ids[0], 1, s->ids[0]);
        handles[r] = h;
    }
}

int intern_slots(Ctx* c, const ReadSlot* d_slots, uint64_t n, const uint32_t* d_scatter, const unsigned long long* d_n, uint32_t* d_handles) {
    if (n == 0) return VSPE_OK;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)c->sm_count * 16);
    k_intern_slots<<<grid, 256, 0, c->stream>>>(link_view(c), d_slots, n, d_scatter, d_n, d_handles);
    VSPE_LAUNCH_CHECK(c);
    return VSPE_OK;
}

// handles -> slots with ascending ids (the order single_end_read_mapping returns, PE_Inference.py:36-47)
__global__ void __launch_bounds__(256)
k_export_slots(LinkView lv, const uint32_t* __restrict__ handles, uint64_t n, ReadSlot* __restrict__ slots) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t h = handles[i];
    ReadSlot* s = slots + i;
    if (h == H_N) { s->hdr = ST_N; return; }
    if (h == H_SHORT) { s->hdr = ST_SHORT; return; }
    const ListRef l = list_ref(lv, h);
    if (l.n > (uint32_t)LR_IDS) {                           // spill lists come from the slot-writing tiers: already ascending
        s->hdr = ST_OK | (l.n << 8);
        s->ids[0] = (uint32_t)(l.p - lv.spill);
        return;
    }
    uint32_t v[LR_IDS];
    for (uint32_t a = 0; a < l.n; a++) {                    // insertion sort, n <= 13
        const uint32_t x = l[a];
        uint32_t b = a;
        while (b > 0 && v[b - 1] > x) { v[b] = v[b - 1]; b--; }
        v[b] = x;
    }
    for (uint32_t a = 0; a < l.n; a++) s->ids[a] = v[a];
    s->hdr = ST_OK | (l.n << 8);
}

int export_slots(Ctx* c, const uint32_t* d_handles, uint64_t n, ReadSlot* d_slots) {
    if (n == 0) return VSPE_OK;
    k_export_slots<<<(uint32_t)((n + 255) / 256), 256, 0, c->stream>>>(link_view(c), d_handles, n, d_slots);
    VSPE_LAUNCH_CHECK(c);
    return VSPE_OK;
}

// ---------------------------------------------------------------------------------------------
// pair aggregation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix64to32(unsigned long long k) {
    k ^= k >> 33; k *= 0xFF51AFD7ED558CCDull; k ^= k >> 29;
    return (uint32_t)(k >> 17);
}

// Four pairs per thread and turn: the handle loads are 16-byte vectors and the four first probes are in flight
// together (the kernel is a chain of L2 round trips: table key -> claim -> count).  No shared counter is touched
// inside the loop -- new combinations are counted in a register and added once per warp -- so nothing
// serialises on one address.  max_probe bounds a probe sequence in a table that may be too small (the host
// then repeats the batch with a larger one); the largest table (load <= 1/4) is probed without a bound.
__global__ void __launch_bounds__(256)
k_pair_agg(const uint32_t* __restrict__ hf, const uint32_t* __restrict__ hr, uint64_t n_pairs, PairEnt* __restrict__ tab,
           uint32_t pmask, unsigned long long* __restrict__ counters, uint32_t h_limit, uint32_t max_probe) {
    uint32_t c_used = 0, c_n = 0, c_short = 0, c_new = 0;
    bool full = false, internal = false;
    const bool vec = ((reinterpret_cast<uintptr_t>(hf) | reinterpret_cast<uintptr_t>(hr)) & 15) == 0;
    const uint64_t span = (uint64_t)gridDim.x * blockDim.x * 4;
    for (uint64_t p0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; p0 < n_pairs; p0 += span) {
        uint32_t a[4], b[4];
        const uint32_t n_here = (uint32_t)min((uint64_t)4, n_pairs - p0);
        if (vec && n_here == 4) {
            const uint4 va = __ldg(reinterpret_cast<const uint4*>(hf + p0)), vb = __ldg(reinterpret_cast<const uint4*>(hr + p0));
            a[0] = va.x; a[1] = va.y; a[2] = va.z; a[3] = va.w;
            b[0] = vb.x; b[1] = vb.y; b[2] = vb.z; b[3] = vb.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                a[i] = (uint32_t)i < n_here ? __ldg(hf + p0 + i) : 0u;
                b[i] = (uint32_t)i < n_here ? __ldg(hr + p0 + i) : 0u;
            }
        }
        unsigned long long key1[4], k[4];
        uint32_t slot[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            key1[i] = 0;
            slot[i] = 0;
            if ((uint32_t)i >= n_here) continue;
            if (a[i] == H_N || b[i] == H_N) { c_n++; continue; }                  // N before short (PE_Inference.py:160-163)
            if (a[i] == H_SHORT || b[i] == H_SHORT) { c_short++; continue; }
            if (a[i] >= h_limit || b[i] >= h_limit) { internal = true; continue; }   // a read no tier finished: never expected
            c_used++;
            key1[i] = (((unsigned long long)a[i] << 32) | b[i]) + 1ull;
            slot[i] = mix64to32(key1[i]) & pmask;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            k[i] = 0;
            if (key1[i]) k[i] = *reinterpret_cast<volatile unsigned long long*>(&tab[slot[i]].key1);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (!key1[i]) continue;
            uint32_t s = slot[i], probes = 0;
            unsigned long long kk = k[i];
            while (true) {
                if (kk == 0) {
                    kk = atomicCAS(&tab[s].key1, 0ull, key1[i]);
                    if (kk == 0) { c_new++; kk = key1[i]; }
                }
                if (kk == key1[i]) { atomicAdd(&tab[s].count, 1u); break; }
                if (++probes > max_probe) { full = true; break; }
                s = (s + 1) & pmask;
                kk = *reinterpret_cast<volatile unsigned long long*>(&tab[s].key1);
            }
        }
    }
    if (full) atomicOr(&counters[CNT_ERR], (unsigned long long)ERRF_PAIRS_FULL);
    if (internal) atomicOr(&counters[CNT_ERR], (unsigned long long)ERRF_INTERNAL);
    // pair counters: one atomic per warp and counter
    for (int d = 16; d; d >>= 1) {
        c_used += __shfl_xor_sync(0xFFFFFFFFu, c_used, d);
        c_n += __shfl_xor_sync(0xFFFFFFFFu, c_n, d);
        c_short += __shfl_xor_sync(0xFFFFFFFFu, c_short, d);
        c_new += __shfl_xor_sync(0xFFFFFFFFu, c_new, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (c_used) atomicAdd(&counters[CNT_B_USED], (unsigned long long)c_used);
        if (c_n) atomicAdd(&counters[CNT_B_N], (unsigned long long)c_n);
        if (c_short) atomicAdd(&counters[CNT_B_SHORT], (unsigned long long)c_short);
        if (c_new) atomicAdd(&counters[CNT_PAIR_OCC], (unsigned long long)c_new);
    }
}

__device__ __forceinline__ void add_exp(unsigned long long* counters, unsigned long long exp, unsigned long long keys) {
    for (int d = 16; d; d >>= 1) {
        exp += __shfl_xor_sync(0xFFFFFFFFu, exp, d);
        keys += __shfl_xor_sync(0xFFFFFFFFu, keys, d);
    }
    if ((threadIdx.x & 31) == 0 && exp) {
        atomicAdd(&counters[CNT_EXP], exp);
        atomicAdd(&counters[CNT_KEYS], keys);
    }
}

// every distinct combination: weights of its lists, number of node_mat keys.  The whole table is visited once, here;
// the occupied entries are listed (one atomic per block and turn) so that every lane of k_comb_emit has a combination
// to expand.
__global__ void __launch_bounds__(256)
k_comb_weigh(LinkView lv, const PairEnt* __restrict__ tab, uint64_t cap, uint32_t* __restrict__ pocc, unsigned long long* __restrict__ pocc_n) {
    __shared__ uint32_t s_wc[2][8];
    __shared__ unsigned long long s_b[2];
    if (blockIdx.x == 0 && threadIdx.x < 3)                 // the batch was accepted: its pair classes count
        atomicAdd(&lv.counters[CNT_USED + (threadIdx.x == 0 ? 0 : threadIdx.x == 1 ? CNT_N - CNT_USED : CNT_SHORT - CNT_USED)],
                  lv.counters[CNT_B_USED + threadIdx.x]);
    unsigned long long exp = 0, keys = 0;
    const uint64_t span = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int turn = 0;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 < cap; i0 += span, turn ^= 1) {       // block-uniform trip count
        const uint64_t i = i0 + threadIdx.x;
        PairEnt e;
        e.key1 = 0; e.count = 0; e.pad = 0;
        if (i < cap) e = tab[i];
        const bool occ = e.key1 != 0;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, occ);
        if (lane == 0) s_wc[turn][wid] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < 8; w++) tot += s_wc[turn][w];
            s_b[turn] = tot ? atomicAdd(pocc_n, (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        if (!occ) continue;
        unsigned long long pos = s_b[turn] + __popc(m & ((1u << lane) - 1));
        for (uint32_t w = 0; w < wid; w++) pos += s_wc[turn][w];
        pocc[pos] = (uint32_t)i;
        const unsigned long long k = e.key1 - 1;
        const uint32_t a = (uint32_t)(k >> 32), b = (uint32_t)k;
        atomicAdd(&lv.recs[a].used, e.count);
        atomicAdd(&lv.recs[b].used, e.count);
        const unsigned long long mm = (unsigned long long)(lv.recs[a].nplus1 - 1) * (lv.recs[b].nplus1 - 1);
        exp += mm;
        keys += mm * e.count;
    }
    add_exp(lv.counters, exp, keys);
}

// list number i of the call: the interned ones first (insertion order), then the private records
__device__ __forceinline__ uint32_t list_handle_at(const LinkView& lv, uint64_t i, uint64_t n_tab) {
    return i < n_tab ? lv.occ[i] : lv.T + (uint32_t)(i - n_tab);
}

__global__ void __launch_bounds__(256)
k_list_weigh(LinkView lv) {
    const unsigned long long n_tab = lv.counters[CNT_LISTS], n = n_tab + lv.counters[CNT_OVF];
    unsigned long long exp = 0, keys = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const ListRec* e = lv.recs + list_handle_at(lv, i, n_tab);
        const unsigned long long u = e->used, cnt = e->nplus1 - 1, m = cnt * (cnt + 1) / 2;
        if (u) { exp += m; keys += m * u; }
    }
    add_exp(lv.counters, exp, keys);
}

// reserve m consecutive output positions for this thread (one atomic per warp)
__device__ __forceinline__ unsigned long long warp_alloc(unsigned long long* cursor, unsigned long long m) {
    const uint32_t lane = threadIdx.x & 31;
    unsigned long long inc = m;
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += y;
    }
    unsigned long long base = 0;
    const unsigned long long tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
    if (lane == 31 && tot) base = atomicAdd(cursor, tot);
    base = __shfl_sync(0xFFFFFFFFu, base, 31);
    return base + inc - m;
}

// PACKED (dense mode: cell numbers and weights both fit 32 bits): one word key << 32 | weight per key, vals unused
template <bool PACKED>
__global__ void __launch_bounds__(256)
k_comb_emit(LinkView lv, PairEnt* __restrict__ tab, const uint32_t* __restrict__ pocc, const unsigned long long* __restrict__ pocc_n, uint64_t N,
            unsigned long long* __restrict__ keys, unsigned long long* __restrict__ vals) {
    const uint64_t n = *pocc_n;
    const uint64_t span = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 < n; i0 += span) {      // warp-uniform trip count
        const uint64_t i = i0 + threadIdx.x;
        ListRef L = {nullptr, 0, 0}, R = {nullptr, 0, 0};
        unsigned long long w = 0;
        if (i < n) {
            PairEnt* e = tab + pocc[i];
            const unsigned long long k = e->key1 - 1;
            w = e->count;
            L = list_ref(lv, (uint32_t)(k >> 32));
            R = list_ref(lv, (uint32_t)k);
            e->key1 = 0;                                                // the table is clean again for the next batch
            e->count = 0;
        }
        unsigned long long o = warp_alloc(lv.counters + CNT_EXP_CURSOR, (unsigned long long)L.n * R.n);
        for (uint32_t a = 0; a < L.n; a++) {
            const unsigned long long row = (unsigned long long)L[a] * N;
            for (uint32_t b = 0; b < R.n; b++) {
                if (PACKED) keys[o++] = ((row + R[b]) << 32) | w;
                else { keys[o] = row + R[b]; vals[o++] = w; }
            }
        }
    }
}

template <bool PACKED>
__global__ void __launch_bounds__(256)
k_list_emit(LinkView lv, uint64_t N, unsigned long long* __restrict__ keys, unsigned long long* __restrict__ vals) {
    const unsigned long long n_tab = lv.counters[CNT_LISTS], n = n_tab + lv.counters[CNT_OVF];
    const uint64_t span = (uint64_t)gridDim.x * blockDim.x;
    const unsigned long long NN = N * N;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 < n; i0 += span) {
        const uint64_t i = i0 + threadIdx.x;
        ListRef A = {nullptr, 0, 0};
        unsigned long long w = 0;
        if (i < n) {
            const uint32_t h = list_handle_at(lv, i, n_tab);
            w = lv.recs[h].used;
            if (w) { A = list_ref(lv, h); lv.recs[h].used = 0; }
        }
        unsigned long long o = warp_alloc(lv.counters + CNT_EXP_CURSOR, (unsigned long long)A.n * (A.n + 1) / 2);
        for (uint32_t a = 0; a < A.n; a++) {
            const uint32_t x = A[a];
            for (uint32_t b = a; b < A.n; b++) {
                const uint32_t y = A[b];
                const unsigned long long key = NN + (unsigned long long)min(x, y) * N + max(x, y);
                if (PACKED) keys[o++] = (key << 32) | w;
                else { keys[o] = key; vals[o++] = w; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Dense mode: the weighted keys are counted by a two-digit most-significant-digit radix sort whose
// last digit is fused with the run-length reduce:
//   k_wkey_hist     keys per bucket (bucket = key >> low_bits)
//   k_wkey_scan     exclusive scan -> bucket segments
//   k_wkey_scatter  (low digit, weight) scattered into the key's bucket segment (radix partition)
//   k_bucket_count  one CTA per bucket (a large bucket: up to 16 CTAs over shares of its segment): counting sort of the
//                   low digit in shared memory; the summed weight of a digit value IS the run total of that key (of the
//                   share), added to the matrix cell -- plain read-modify-write where one CTA owns the cell, integer
//                   atomics where the bucket was shared.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_wkey_hist(const unsigned long long* __restrict__ keys, uint64_t n, uint64_t per_block, uint32_t low_bits, uint32_t n_buckets,
            uint32_t* __restrict__ g_hist) {
    extern __shared__ uint32_t s_h[];
    for (uint32_t b = threadIdx.x; b < n_buckets; b += blockDim.x) s_h[b] = 0;
    __syncthreads();
    const uint64_t lo = (uint64_t)blockIdx.x * per_block, hi = min(n, lo + per_block);
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&s_h[(uint32_t)(keys[i] >> (32 + low_bits))], 1u);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < n_buckets; b += blockDim.x) {
        const uint32_t c = s_h[b];
        if (c) atomicAdd(&g_hist[b], c);
    }
}

__global__ void __launch_bounds__(1024)
k_wkey_scan(const uint32_t* __restrict__ hist, uint32_t n, uint32_t* __restrict__ start, uint32_t* __restrict__ cursor) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t off = 0; off < n; off += 1024) {
        const uint32_t i = off + threadIdx.x;
        const uint32_t x = i < n ? hist[i] : 0;
        uint32_t inc = x;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        uint32_t wbase = 0, tot = 0;
        for (uint32_t w = 0; w < 32; w++) {
            const uint32_t t = s_warp[w];
            if (w < wid) wbase += t;
            tot += t;
        }
        const uint32_t carry = s_carry;
        if (i < n) { start[i] = carry + wbase + inc - x; cursor[i] = carry + wbase + inc - x; }
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[n] = s_carry;
}

__global__ void __launch_bounds__(256)
k_wkey_scatter(const unsigned long long* __restrict__ keys, uint64_t n, uint64_t per_block,
               uint32_t low_bits, uint32_t n_buckets, uint32_t* __restrict__ g_cursor, uint32_t* __restrict__ out_k,
               uint32_t* __restrict__ out_w) {
    extern __shared__ uint32_t s_mem[];
    uint32_t* s_cnt = s_mem;                    // [n_buckets] keys of this block per bucket, then its running cursor
    uint32_t* s_base = s_mem + n_buckets;       // [n_buckets] start of this block's part of the bucket segment
    for (uint32_t b = threadIdx.x; b < n_buckets; b += blockDim.x) s_cnt[b] = 0;
    __syncthreads();
    const uint64_t lo = (uint64_t)blockIdx.x * per_block, hi = min(n, lo + per_block);
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&s_cnt[(uint32_t)(keys[i] >> (32 + low_bits))], 1u);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < n_buckets; b += blockDim.x) {
        const uint32_t c = s_cnt[b];
        if (c) s_base[b] = atomicAdd(&g_cursor[b], c);
        s_cnt[b] = 0;
    }
    __syncthreads();
    const uint32_t mask = (1u << low_bits) - 1;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const unsigned long long kw = keys[i];                 // cell << 32 | weight
        const uint32_t k = (uint32_t)(kw >> 32);
        const uint32_t b = k >> low_bits;
        const uint32_t pos = s_base[b] + atomicAdd(&s_cnt[b], 1u);
        out_k[pos] = k & mask;
        out_w[pos] = (uint32_t)kw;
    }
}

// blockIdx.x = bucket, blockIdx.y = part: a bucket with more than BC_SPLIT keys is counted by up to gridDim.y CTAs, each
// over a contiguous share of its segment; their partial run totals are integer-added to the cells (order independent).
// A small bucket is counted by its part 0 alone, which owns the cells.
static constexpr uint32_t BC_PARTS = 16;
__global__ void __launch_bounds__(256)
k_bucket_count(const uint32_t* __restrict__ k_lo, const uint32_t* __restrict__ w, const uint32_t* __restrict__ start,
               uint32_t low_bits, uint64_t n_cells, uint64_t* __restrict__ mats, const uint32_t split) {
    extern __shared__ uint32_t s_bins[];
    const uint32_t bucket = blockIdx.x;
    const uint32_t s0 = start[bucket], e0 = start[bucket + 1];
    if (s0 == e0) return;
    const uint32_t n = e0 - s0;
    const uint32_t parts = n <= split ? 1u : min((uint32_t)gridDim.y, (n + split - 1) / split);
    if (blockIdx.y >= parts) return;
    const uint32_t per = (n + parts - 1) / parts;
    const uint32_t lo = s0 + blockIdx.y * per, hi = min(e0, lo + per);
    const uint32_t nb = 1u << low_bits;
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) s_bins[i] = 0;
    __syncthreads();
    // (folding equal cells of a warp with __match_any_sync before the atomic was measured slower on every config)
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&s_bins[k_lo[i]], w[i]);
    __syncthreads();
    const uint64_t base = (uint64_t)bucket << low_bits;
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
        const uint32_t c = s_bins[i];
        if (c && base + i < n_cells) {
            if (parts == 1) mats[base + i] += c;
            else atomicAdd(reinterpret_cast<unsigned long long*>(mats) + base + i, (unsigned long long)c);
        }
    }
}

// keys: the weighted keys of one batch, cell << 32 | weight (device array of n entries)
static int dense_accumulate(Ctx* c, const unsigned long long* keys, uint64_t n) {
    const uint64_t N = c->index.n_nodes, cells = 2 * N * N;
    if (n == 0 || cells == 0) return VSPE_OK;
    if (n > 0xFFFFFFF0ull) { set_error("key batch too large"); return VSPE_ERR_LIMIT; }
    uint32_t low_bits = 7;
    while (((cells + (1ull << low_bits) - 1) >> low_bits) > 2048 && low_bits < 15) low_bits++;
    const uint32_t n_buckets = (uint32_t)((cells + (1ull << low_bits) - 1) >> low_bits);     // <= 8192 (dense_possible)
    cudaStream_t st = c->stream;
    VSPE_TRY(c->wk_hist.reserve(3ull * (n_buckets + 1)));
    VSPE_TRY(c->wk_keys.reserve(2 * n));
    uint32_t* g_hist = c->wk_hist.p;
    uint32_t* g_start = g_hist + (n_buckets + 1);
    uint32_t* g_cursor = g_start + (n_buckets + 1);
    uint32_t* out_k = c->wk_keys.p;
    uint32_t* out_w = out_k + n;
    if (!c->link_attr_set) {
        VSPE_CUDA(cudaFuncSetAttribute(k_wkey_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
        VSPE_CUDA(cudaFuncSetAttribute(k_bucket_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << 15) * 4));
        c->link_attr_set = true;
    }
    const uint64_t blocks = std::min<uint64_t>((n + 8191) / 8192, (uint64_t)c->sm_count * 8);
    const uint64_t per_block = (n + blocks - 1) / blocks;
    VSPE_CUDA(cudaMemsetAsync(g_hist, 0, (n_buckets + 1) * 4, st));
    k_wkey_hist<<<(uint32_t)blocks, 256, n_buckets * 4, st>>>(keys, n, per_block, low_bits, n_buckets, g_hist);
    VSPE_LAUNCH_CHECK(c);
    k_wkey_scan<<<1, 1024, 0, st>>>(g_hist, n_buckets, g_start, g_cursor);
    VSPE_LAUNCH_CHECK(c);
    k_wkey_scatter<<<(uint32_t)blocks, 256, n_buckets * 8, st>>>(keys, n, per_block, low_bits, n_buckets, g_cursor, out_k, out_w);
    VSPE_LAUNCH_CHECK(c);
    k_bucket_count<<<dim3(n_buckets, BC_PARTS), 256, (1u << low_bits) * 4, st>>>(out_k, out_w, g_start, low_bits, cells, c->mats.p,
                                                                                     (uint32_t)c->opt_link_split);
    VSPE_LAUNCH_CHECK(c);
    return VSPE_OK;
}

int count_links(Ctx* c, const uint32_t* d_hf, const uint32_t* d_hr, uint64_t total) {
    const uint64_t N = c->index.n_nodes;
    c->stats.total_pairs += total;
    if (total == 0) return VSPE_OK;
    cudaStream_t st = c->stream;
    Sparse& sp = c->sparse;
    const uint32_t grid_cap = (uint32_t)c->sm_count * 8;
    for (uint64_t off = 0; off < total; off += LINK_BATCH) {
        const uint64_t n = std::min<uint64_t>(LINK_BATCH, total - off);
        // pair table: sized for the distinct combinations, not for the pairs (it should stay L2 resident): it starts
        // at 2^21 entries (32 MB) and a batch that fills it beyond one half is repeated with a table 4x larger, up to
        // 4 n entries (n distinct combinations at load 1/4).  All zero between batches (k_comb_emit cleans up).
        uint64_t cap_max = 1u << 16;
        while (cap_max < 4 * n) cap_max <<= 1;
        const uint64_t cap_first = c->opt_pair_cap_log2 > 0 ? 1ull << std::min<int64_t>(c->opt_pair_cap_log2, 40) : cap_max;
        uint64_t cap = std::max<uint64_t>(std::min<uint64_t>(cap_max, cap_first), std::min<uint64_t>(c->pair_cap, cap_max));
        const uint32_t grid = (uint32_t)std::min<uint64_t>((n + 1023) / 1024, grid_cap);      // four pairs per thread and turn
        const LinkView lv = link_view(c);
        for (;;) {
            if (cap != c->pair_cap) {
                c->pair_tab.release();
                VSPE_TRY(c->pair_tab.reserve(cap));
                VSPE_CUDA(cudaMemsetAsync(c->pair_tab.p, 0, cap * sizeof(PairEnt), st));
                c->pair_cap = cap;
            }
            VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_PAIR_OCC, 0, 6 * 8, st));     // PAIR_OCC, EXP, EXP_CURSOR, B_USED, B_N, B_SHORT
            VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_PAIR_LIST, 0, 8, st));
            k_pair_agg<<<grid, 256, 0, st>>>(d_hf + off, d_hr + off, n, c->pair_tab.p, (uint32_t)(c->pair_cap - 1), c->counters.p,
                                         c->list_T + c->list_ov_cap, cap == cap_max ? 0xFFFFFFFFu : 256u);
            VSPE_LAUNCH_CHECK(c);
            unsigned long long h_err = 0, h_occ = 0;
            VSPE_CUDA(cudaMemcpyAsync(&h_err, c->counters.p + CNT_ERR, 8, cudaMemcpyDeviceToHost, st));
            VSPE_CUDA(cudaMemcpyAsync(&h_occ, c->counters.p + CNT_PAIR_OCC, 8, cudaMemcpyDeviceToHost, st));
            VSPE_CUDA(cudaStreamSynchronize(st));
            // accepted: every probe sequence ended and the table is at most half full (the largest table is always accepted)
            if (!(h_err & ERRF_PAIRS_FULL) && (cap == cap_max || h_occ <= cap / 2)) break;
            const unsigned long long cleared = h_err & ~(unsigned long long)ERRF_PAIRS_FULL;
            VSPE_CUDA(cudaMemcpyAsync(c->counters.p + CNT_ERR, &cleared, 8, cudaMemcpyHostToDevice, st));
            VSPE_CUDA(cudaMemsetAsync(c->pair_tab.p, 0, c->pair_cap * sizeof(PairEnt), st));
            cap = std::min<uint64_t>(cap * 4, cap_max);
        }
        VSPE_TRY(c->pair_occ.reserve(std::min<uint64_t>(n, c->pair_cap) + 1));
        k_comb_weigh<<<grid_cap, 256, 0, st>>>(lv, c->pair_tab.p, c->pair_cap, c->pair_occ.p, c->counters.p + CNT_PAIR_LIST);
        VSPE_LAUNCH_CHECK(c);
        k_list_weigh<<<grid_cap, 256, 0, st>>>(lv);
        VSPE_LAUNCH_CHECK(c);
        unsigned long long h_exp = 0, h_err = 0;
        VSPE_CUDA(cudaMemcpyAsync(&h_exp, c->counters.p + CNT_EXP, 8, cudaMemcpyDeviceToHost, st));
        VSPE_CUDA(cudaMemcpyAsync(&h_err, c->counters.p + CNT_ERR, 8, cudaMemcpyDeviceToHost, st));
        VSPE_CUDA(cudaStreamSynchronize(st));
        c->last_err_flags = h_err;                          // every scan / map kernel of this call ran before
        c->err_flags_fresh = true;
        if (h_err & ERRF_FATAL) {                           // the caller reports it; leave the pair table clean
            VSPE_CUDA(cudaMemsetAsync(c->pair_tab.p, 0, c->pair_cap * sizeof(PairEnt), st));
            return VSPE_OK;
        }
        if (h_exp == 0) continue;
        // the accumulated runs (sparse mode) stay at the front of k[0] / v[0]; the new keys go right behind them
        VSPE_TRY(sparse_reserve(c, sp.n_runs + h_exp));
        unsigned long long* keys = sp.k[0].p + sp.n_runs;
        unsigned long long* vals = sp.v[0].p + sp.n_runs;
        if (sp.enabled) k_comb_emit<false><<<grid_cap, 256, 0, st>>>(lv, c->pair_tab.p, c->pair_occ.p, c->counters.p + CNT_PAIR_LIST, N, keys, vals);
        else k_comb_emit<true><<<grid_cap, 256, 0, st>>>(lv, c->pair_tab.p, c->pair_occ.p, c->counters.p + CNT_PAIR_LIST, N, keys, vals);
        VSPE_LAUNCH_CHECK(c);
        if (sp.enabled) k_list_emit<false><<<grid_cap, 256, 0, st>>>(lv, N, keys, vals);
        else k_list_emit<true><<<grid_cap, 256, 0, st>>>(lv, N, keys, vals);
        VSPE_LAUNCH_CHECK(c);
        if (sp.enabled) {
            uint64_t runs = 0;
            VSPE_TRY(sparse_sort_reduce(c, sp.n_runs + h_exp, &runs));
            sp.n_runs = runs;
        } else {
            VSPE_TRY(dense_accumulate(c, keys, h_exp));
        }
    }
    return VSPE_OK;
}

}  // namespace vspe
