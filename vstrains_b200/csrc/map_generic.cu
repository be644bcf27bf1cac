// map_generic.cu -- K4, exhaustive tier: one warp per read, one lane per (k+1)-mer position.
//
// Replaces single_end_read_mapping (reference utils/VStrains_PE_Inference.py:16-48) for ANY
// input the reference accepts: arbitrary read length, non-ACGT characters (their k-mers never
// match), repeated / palindromic k-mers with many postings, any number of nodes per read.
// It hashes straight from the ASCII bytes, so it needs no packed copy of the read.  The
// seed-and-extend tier (map_fast.cu) sends here every read it cannot prove exact.
//
// Per warp, in global memory (L2 resident for viral graphs): v[N] hit counts, kmin[N] smallest
// hit position, list[N + TOUCH_CAP] touched nodes.  The arrays are self-cleaning.
#include "ctx.cuh"

namespace vspe {

static constexpr int MG_WARPS = 8;
static constexpr uint32_t TOUCH_CAP = 256;
static constexpr uint32_t G_STAGE = 1024;   // bytes of one read staged in shared memory per warp
static constexpr uint32_t G_WORDS = 20;     // 16-base words kept per window for verification (L <= 320)

__device__ __forceinline__ bool keep_node(uint32_t v, uint32_t kmin, uint32_t len, uint32_t rlen, uint32_t L) {
    // PE_Inference.py:36-47 in integers (see oracle/pe_oracle.py:map_read)
    long long m = (long long)min((long long)len, (long long)rlen - kmin);
    long long sat = m - L + 1;
    long long ab = ((long long)min(rlen, len) - L + 1) * ((long long)rlen - L);
    return (long long)v >= sat || (long long)v * rlen >= ab;
}

__global__ void __launch_bounds__(MG_WARPS * 32)
k_map_generic(IndexView ix, const uint8_t* __restrict__ buf, const uint64_t* __restrict__ seq_start,
              const uint64_t* __restrict__ seq_end, const uint32_t* __restrict__ worklist, uint64_t n_items_arg,
              const unsigned long long* __restrict__ n_items_dev, uint64_t buf_n, ReadSlot* __restrict__ slots, uint32_t* __restrict__ scratch, uint64_t scratch_stride,
              uint32_t* __restrict__ spill, uint64_t spill_cap, unsigned long long* __restrict__ counters) {
    __shared__ uint32_t s_ntouch[MG_WARPS];
    __shared__ uint8_t s_seq[MG_WARPS][G_STAGE];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t gwarp = (uint64_t)blockIdx.x * MG_WARPS + wib, nwarps = (uint64_t)gridDim.x * MG_WARPS;
    const uint32_t N = ix.n_nodes, L = ix.split_len;
    const uint64_t n_items = n_items_dev ? *n_items_dev : n_items_arg;
    uint32_t* v = scratch + gwarp * scratch_stride;
    uint32_t* kmin = v + N;
    uint32_t* list = kmin + N;            // [TOUCH_CAP] unsorted, then [N + ...] sorted / compacted
    uint32_t* sorted = list + TOUCH_CAP;

    for (uint64_t item = gwarp; item < n_items; item += nwarps) {
        const uint64_t r = worklist ? worklist[item] : item;
        const uint64_t s = seq_start[r];
        uint64_t e = seq_end[r];
        if (e == ~0ull) {
            // the scan kernel did not see the end of this (very long) line: first '\n' or '\r'
            e = buf_n;
            for (uint64_t q = s; q < buf_n; q += 32) {
                const uint64_t i = q + lane;
                const bool hit = i < buf_n && (buf[i] == '\n' || buf[i] == '\r');
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
                if (m) { e = q + (uint32_t)(__ffs((int)m) - 1); break; }
            }
        }
        const uint64_t rlen64 = e - s;
        const uint8_t* seq = buf + s;
        // stage the read in shared memory (coalesced) so the per-window loops hit LDS, and test
        // for 'N' on the way: 'N' anywhere -> pair is skipped, checked before the length (:160-163)
        const bool staged = rlen64 <= G_STAGE;
        bool hasN = false;
        for (uint64_t i = lane; i < rlen64; i += 32) {
            const uint8_t ch = seq[i];
            hasN |= (ch == 'N');
            if (staged) s_seq[wib][i] = ch;
        }
        hasN = __any_sync(0xFFFFFFFFu, hasN);
        if (hasN || rlen64 < L) {
            if (lane == 0) slots[r].hdr = hasN ? ST_N : ST_SHORT;
            continue;
        }
        if (rlen64 > 0x7FFFFFFFull) {   // absurd line length: treat like the limit error
            if (lane == 0) { atomicOr(&counters[CNT_ERR], (unsigned long long)ERRF_SPILL_FULL); slots[r].hdr = ST_OK; }
            continue;
        }
        const uint8_t* sq = staged ? s_seq[wib] : seq;
        const uint32_t rlen = (uint32_t)rlen64, npos = rlen - L + 1;
        const uint32_t n32 = (L + 15) >> 4;
        const bool keep_words = n32 <= G_WORDS;
        if (lane == 0) s_ntouch[wib] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < npos; i += 32) {
            // hash the window from ASCII, feeding the same 16-base words hash_packed sees
            KmerHash hs;
            uint32_t w = 0, wv[G_WORDS];
            bool valid = true;
            for (uint32_t j = 0; j < L; j++) {
                const uint32_t c = sq[i + j];
                if (!is_acgt(c)) { valid = false; break; }
                w |= base_code(c) << (2 * (j & 15));
                if ((j & 15) == 15 || j == L - 1) {
                    hs.add(w);
                    if (keep_words) wv[j >> 4] = w;
                    w = 0;
                }
            }
            if (!valid) continue;
            const uint64_t h = hs.finish();
            if (ix.bloom != nullptr && !bloom_maybe(ix.bloom, ix.bloom_mask, h)) continue;   // proven miss
            uint32_t j = slot_of(h, ix.slot_mask);
            while (true) {
                const uint2 ent = __ldg(ix.slots + j);
                if (ent.x == EMPTY_TP) break;
                if (fp_match(ent.y, h, ix.node_mask)) {
                    bool eq = true;
                    if (keep_words) {
                        for (uint32_t m = 0; m < n32 && eq; m++) {
                            uint32_t tw = (uint32_t)extract64(ix.text, (uint64_t)ent.x + 16 * m);
                            const uint32_t rem = L - 16 * m;
                            if (rem < 16) tw &= (1u << (2 * rem)) - 1;
                            eq = tw == wv[m];
                        }
                    } else {
                        for (uint32_t t = 0; t < L; t++)
                            if (text_base(ix.text, (uint64_t)ent.x + t) != base_code(sq[i + t])) { eq = false; break; }
                    }
                    if (eq) {
                        const uint32_t node = ent.y & ix.node_mask;
                        const uint32_t old = atomicAdd(&v[node], 1u);
                        atomicMin(&kmin[node], i);
                        if (old == 0) {
                            const uint32_t idx = atomicAdd(&s_ntouch[wib], 1u);
                            if (idx < TOUCH_CAP) list[idx] = node;
                        }
                    }
                }
                j = (j + 1) & ix.slot_mask;
            }
        }
        __threadfence_block();
        __syncwarp();
        const uint32_t nt = s_ntouch[wib];
        uint32_t n_out = 0;
        if (nt <= TOUCH_CAP) {
            // rank sort of the distinct touched nodes
            for (uint32_t a = lane; a < nt; a += 32) {
                uint32_t x = list[a], rank = 0;
                for (uint32_t b = 0; b < nt; b++) rank += (list[b] < x);
                sorted[rank] = x;
            }
            __syncwarp();
            for (uint32_t base = 0; base < nt; base += 32) {
                uint32_t a = base + lane;
                uint32_t node = a < nt ? sorted[a] : 0;
                bool keep = false;
                if (a < nt) {
                    keep = keep_node(__ldcg(v + node), __ldcg(kmin + node), ix.node_len[node], rlen, L);
                    v[node] = 0;
                    kmin[node] = NONE32;
                }
                uint32_t m = __ballot_sync(0xFFFFFFFFu, keep);
                __syncwarp();
                if (keep) sorted[n_out + __popc(m & ((1u << lane) - 1))] = node;   // in-place: target <= a
                n_out += __popc(m);
                __syncwarp();
            }
        } else {
            // too many distinct nodes: ordered scan over all nodes, like the reference does
            for (uint32_t base = 0; base < N; base += 32) {
                uint32_t node = base + lane;
                bool keep = false;
                if (node < N) {
                    uint32_t vv = __ldcg(v + node);
                    if (vv) {
                        keep = keep_node(vv, __ldcg(kmin + node), ix.node_len[node], rlen, L);
                        v[node] = 0;
                        kmin[node] = NONE32;
                    }
                }
                uint32_t m = __ballot_sync(0xFFFFFFFFu, keep);
                if (keep) sorted[n_out + __popc(m & ((1u << lane) - 1))] = node;
                n_out += __popc(m);
            }
            __syncwarp();
        }
        __syncwarp();
        ReadSlot* out = slots + r;
        if (n_out <= SLOT_IDS) {
            if (lane < n_out) out->ids[lane] = sorted[lane];
            if (lane == 0) out->hdr = ST_OK | (n_out << 8);
        } else {
            unsigned long long off = 0;
            if (lane == 0) off = atomicAdd(&counters[CNT_SPILL_CURSOR], (unsigned long long)n_out);
            off = __shfl_sync(0xFFFFFFFFu, off, 0);
            if (off + n_out > spill_cap) {
                if (lane == 0) { atomicOr(&counters[CNT_ERR], (unsigned long long)ERRF_SPILL_FULL); out->hdr = ST_OK; }
            } else {
                for (uint32_t a = lane; a < n_out; a += 32) spill[off + a] = sorted[a];
                if (lane == 0) { out->ids[0] = (uint32_t)off; out->hdr = ST_OK | (n_out << 8); }
            }
        }
        __syncwarp();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        atomicAdd(&counters[CNT_GENERIC], (unsigned long long)n_items);
        if (worklist) atomicAdd(&counters[CNT_BAILED], (unsigned long long)n_items);
    }
}

__global__ void k_fill_u32(uint32_t* p, uint64_t n, uint32_t val) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = val;
}

// lays out [v | kmin | list] per warp and initialises v = 0, kmin = NONE
static int prepare_scratch(Ctx* c, uint32_t n_blocks, uint64_t* stride_out) {
    uint32_t N = c->index.n_nodes;
    uint64_t stride = 3ull * N + 2 * TOUCH_CAP + 32;
    uint64_t total = stride * n_blocks * MG_WARPS;
    *stride_out = stride;
    size_t before = c->warp_scratch.cap;
    VSPE_TRY(c->warp_scratch.reserve(total));
    if (c->warp_scratch.cap != before || !c->scratch_valid) {
        // v = 0 everywhere, then kmin = NONE for every warp (one strided 2-D memset)
        VSPE_CUDA(cudaMemsetAsync(c->warp_scratch.p, 0, c->warp_scratch.cap * 4, c->stream));
        if (N) {
            VSPE_CUDA(cudaMemset2DAsync(c->warp_scratch.p + N, stride * 4, 0xFF, (size_t)N * 4,
                                        (size_t)n_blocks * MG_WARPS, c->stream));
        }
        c->scratch_valid = true;
    }
    return VSPE_OK;
}

// Grid of the exhaustive tier: fixed per index (the scratch layout depends on it), bounded so the
// per-warp scratch (3N u32) stays within a few GiB for very large graphs.
static uint32_t generic_blocks(Ctx* c) {
    uint64_t cap = (uint64_t)c->sm_count * 8;
    uint64_t per_block = (3ull * c->index.n_nodes + 2 * TOUCH_CAP + 32) * MG_WARPS * 4;
    uint64_t budget = 8ull << 30;
    if (per_block * cap > budget) cap = budget / per_block ? budget / per_block : 1;
    return (uint32_t)cap;
}

int map_reads_generic_list(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                           const uint32_t* d_worklist, uint64_t n_items, ReadSlot* d_slots) {
    if (n_items == 0) return VSPE_OK;
    // the scratch layout depends on the grid, so keep the grid fixed per context size class
    uint32_t nb = generic_blocks(c);
    uint64_t stride;
    VSPE_TRY(prepare_scratch(c, nb, &stride));
    if (!c->spill.p) VSPE_TRY(c->spill.reserve(4u << 20));
    uint64_t want = (n_items + MG_WARPS - 1) / MG_WARPS;
    uint32_t grid = (uint32_t)(want < nb ? want : nb);
    k_map_generic<<<grid, MG_WARPS * 32, 0, c->stream>>>(c->index.view(), d_buf, d_seq_start, d_seq_end, d_worklist, n_items, nullptr, c->cur_buf_n,
                                                         d_slots, c->warp_scratch.p, stride, c->spill.p, c->spill.cap, c->counters.p);
    VSPE_LAUNCH_CHECK(c);
    return VSPE_OK;
}

int map_reads_generic(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                      uint64_t n_reads, ReadSlot* d_slots) {
    return map_reads_generic_list(c, d_buf, d_seq_start, d_seq_end, nullptr, n_reads, d_slots);
}

// worklist whose length lives on the device (no host round trip): a modest persistent grid
int map_reads_generic_dev(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                          const uint32_t* d_worklist, const unsigned long long* d_n_items, ReadSlot* d_slots) {
    uint32_t nb = generic_blocks(c);
    uint64_t stride;
    VSPE_TRY(prepare_scratch(c, nb, &stride));
    if (!c->spill.p) VSPE_TRY(c->spill.reserve(4u << 20));
    k_map_generic<<<nb, MG_WARPS * 32, 0, c->stream>>>(c->index.view(), d_buf, d_seq_start, d_seq_end, d_worklist, 0, d_n_items, c->cur_buf_n,
                                                       d_slots, c->warp_scratch.p, stride, c->spill.p, c->spill.cap, c->counters.p);
    VSPE_LAUNCH_CHECK(c);
    return VSPE_OK;
}

}  // namespace vspe
