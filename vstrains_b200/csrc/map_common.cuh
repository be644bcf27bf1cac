// map_common.cuh -- device helpers shared by the map kernels (scan_map.cu, map_fast.cu): the
// integer saturation predicate, packed-row access, the (k+1)-mer probe.
#pragma once
#include "ctx.cuh"

namespace vspe {

// PE_Inference.py:36-47 in integers (coords cancel out)
__device__ __forceinline__ bool keep_node_f(uint32_t v, uint32_t kmin, uint32_t len, uint32_t rlen, uint32_t L) {
    int m = min((int)len, (int)rlen - (int)kmin);
    int sat = m - (int)L + 1;
    long long ab = (long long)(min(rlen, len) - L + 1) * (long long)(rlen - L);
    return (int)v >= sat || (long long)v * rlen >= ab;
}

// 64 bits (32 bases) of a packed read row starting at base b (row has 2 pad words)
__device__ __forceinline__ uint64_t read64(const uint32_t* row, uint32_t b) {
    uint32_t w = b >> 4, s = (b & 15) * 2;
    uint32_t x0 = row[w], x1 = row[w + 1], x2 = row[w + 2];
    uint32_t lo = __funnelshift_r(x0, x1, s), hi = __funnelshift_r(x1, x2, s);
    return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ uint64_t hash_read(const uint32_t* row, uint32_t b, uint32_t L) {
    const uint32_t w0 = b >> 4, sh = (b & 15) * 2, n = (L + 15) >> 4;
    KmerHash hs;
    uint32_t x0 = row[w0];
    for (uint32_t m = 0; m < n; m++) {
        const uint32_t x1 = row[w0 + m + 1];
        uint32_t v = __funnelshift_r(x0, x1, sh);
        if (m == n - 1 && (L & 15)) v &= (1u << (2 * (L & 15))) - 1;
        hs.add(v);
        x0 = x1;
    }
    return hs.finish();
}

__device__ __forceinline__ bool read_equals_text(const uint32_t* row, uint32_t b, const uint64_t* __restrict__ text,
                                                 uint32_t tp, uint32_t L) {
    uint64_t acc = 0;
    for (uint32_t m = 0; m < L; m += 32) {
        uint64_t x = read64(row, b + m) ^ extract64(text, (uint64_t)tp + m);
        const uint32_t rem = L - m;
        if (rem < 32) x &= (1ull << (2 * rem)) - 1;
        acc |= x;
    }
    return acc == 0;
}

enum { PROBE_MISS = 0, PROBE_UNIQUE = 1, PROBE_MULTI = 2 };

// probe window b of a packed row: MISS, the UNIQUE posting, or MULTI (several postings).
// The slot walk only looks for the first fingerprint match (one load + one compare per turn); the
// base-by-base verification is straight-line code after that loop, so the lanes of a warp -- whose
// slot walks differ in length -- run it once and together.  A fingerprint match that fails the
// verification (2^-20 per occupied slot passed) continues in the out-of-line loop.
static __device__ __noinline__ int probe_window_rest(const IndexView& ix, const uint32_t* row, uint32_t b, uint64_t h, uint32_t j, uint32_t& tp, uint32_t& node) {
    const uint32_t L = ix.split_len;
    while (true) {
        const uint2 ent = __ldg(ix.slots + j);
        if (ent.x == EMPTY_TP) return PROBE_MISS;
        if (fp_match(ent.y, h, ix.node_mask) && read_equals_text(row, b, ix.text, ent.x, L)) {
            tp = ent.x;
            node = ent.y & ix.node_mask;
            return ((__ldg(ix.uniq + (ent.x >> 5)) >> (ent.x & 31)) & 1) ? PROBE_UNIQUE : PROBE_MULTI;
        }
        j = (j + 1) & ix.slot_mask;
    }
}

__device__ __forceinline__ int probe_window(const IndexView& ix, const uint32_t* row, uint32_t b, uint32_t& tp, uint32_t& node) {
    const uint32_t L = ix.split_len;
    const uint64_t h = hash_read(row, b, L);
    if (ix.bloom != nullptr && !bloom_maybe(ix.bloom, ix.bloom_mask, h)) return PROBE_MISS;
    uint32_t j = slot_of(h, ix.slot_mask);
    uint2 ent;
    while (true) {
        ent = __ldg(ix.slots + j);
        if (ent.x == EMPTY_TP || fp_match(ent.y, h, ix.node_mask)) break;
        j = (j + 1) & ix.slot_mask;
    }
    if (ent.x == EMPTY_TP) return PROBE_MISS;
    if (!read_equals_text(row, b, ix.text, ent.x, L)) return probe_window_rest(ix, row, b, h, (j + 1) & ix.slot_mask, tp, node);
    tp = ent.x;
    node = ent.y & ix.node_mask;
    return ((__ldg(ix.uniq + (ent.x >> 5)) >> (ent.x & 31)) & 1) ? PROBE_UNIQUE : PROBE_MULTI;
}

// Reverse-complement a packed read row in place (rlen bases in 16-base words): reverse the 2-bit
// groups of every word, complement, realign by the padding of the last word.  NW = data words.
template <int NW>
__device__ __forceinline__ void revcomp_row(uint32_t* row, uint32_t rlen) {
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
}

}  // namespace vspe
