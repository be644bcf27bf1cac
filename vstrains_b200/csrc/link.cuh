// link.cuh -- the interface between the map stage (K4) and the link-count stage (K5 + K6).
//
// The reference appends every read's node list to two matrices pair by pair
// (utils/VStrains_PE_Inference.py:174-188).  Here a read's result is a 32-bit HANDLE of its
// node list: the map kernels intern each distinct list once in a device-resident list table
// (open addressing, one 64-byte record per list), so
//   * a read costs 4 bytes between the stages instead of a 64-byte slot,
//   * the count stage aggregates read PAIRS by (handle_left, handle_right) first -- the distinct
//     combinations are a few 10^4..10^6 for tens of millions of pairs -- and expands every
//     combination ONCE, weighted by its multiplicity, into (matrix, i, j) keys that go through
//     the radix sort + run-length reduce (sparse.cu).
// Everything stays exact: a handle is only shared by reads whose lists were compared id by id.
#pragma once
#include "ctx.cuh"

namespace vspe {

static constexpr uint32_t H_N = 0xFFFFFFFFu;        // read contains an upper-case 'N'   (PE_Inference.py:160)
static constexpr uint32_t H_SHORT = 0xFFFFFFFEu;    // read shorter than split_len       (:162)
static constexpr uint32_t H_PENDING = 0xFFFFFFFDu;  // deferred to a later tier (never reaches the count stage)

struct LinkView {
    ListRec* recs;                    // [T] table records, then [ov_cap] overflow records
    uint32_t T, t_mask, ov_cap;
    uint32_t max_lists;               // soft load limit of the table part (T / 2)
    uint32_t* occ;                    // [T] slots of the interned lists, in insertion order
    uint32_t* spill;                  // ids of lists longer than LR_IDS
    uint64_t spill_cap;
    unsigned long long* counters;
};

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t list_hash(const uint32_t* ids, uint32_t stride, uint32_t n) {
    KmerHash hs;
    for (uint32_t i = 0; i < n; i++) hs.add(ids[i * stride] + 1u);
    hs.add(n ^ 0x5BD1E995u);
    return hs.finish();
}

// A private (not shared) record for one read: lists the table cannot hold (more than LR_IDS ids,
// table at its load limit, probe bound exceeded).  spill_off != NONE32: the ids already live in the
// spill pool (raw, ascending) and are referenced instead of copied.
static __device__ __noinline__ uint32_t list_overflow(const LinkView& lv, uint32_t n, const uint32_t* ids, uint32_t stride,
                                               uint32_t spill_off) {
    const unsigned long long idx = atomicAdd(&lv.counters[CNT_OVF], 1ull);
    if (idx >= lv.ov_cap) {
        atomicOr(&lv.counters[CNT_ERR], (unsigned long long)ERRF_LISTS_FULL);
        return H_PENDING;                                   // the host grows the pool and repeats the launch
    }
    ListRec* e = lv.recs + lv.T + idx;
    e->tag = 1;
    e->used = 0;
    if (n <= (uint32_t)LR_IDS) {
        for (uint32_t i = 0; i < n; i++) e->ids[i] = ids[i * stride] + 1u;
    } else if (spill_off != NONE32) {
        e->ids[0] = spill_off;
    } else {
        const unsigned long long off = atomicAdd(&lv.counters[CNT_SPILL_CURSOR], (unsigned long long)n);
        if (off + n > lv.spill_cap) {
            atomicOr(&lv.counters[CNT_ERR], (unsigned long long)ERRF_SPILL_FULL);
            n = 0;
        } else {
            for (uint32_t i = 0; i < n; i++) lv.spill[off + i] = ids[i * stride];
            e->ids[0] = (uint32_t)off;
        }
    }
    e->nplus1 = n + 1;
    return lv.T + (uint32_t)idx;
}

// Handle of the list ids[0], ids[stride], ... (n entries; any order, no duplicates).  Lists with the
// same ids in the same order share a handle.  Record words are written once and are non-zero once
// visible (ids are stored + 1), so a reader needs no fence: a word it sees as zero is "not yet
// known" (retry), a non-zero word is final.
__device__ __forceinline__ uint32_t intern_list(const LinkView& lv, uint32_t n, const uint32_t* ids, uint32_t stride,
                                                uint32_t spill_off = NONE32) {
    if (n <= (uint32_t)LR_IDS) {
        const uint64_t h = list_hash(ids, stride, n);
        const uint32_t tag = (uint32_t)(h >> 32) | 1u;
        uint32_t slot = (uint32_t)h & lv.t_mask;
        for (int probe = 0; probe < 48;) {
            ListRec* e = lv.recs + slot;
            uint4 w0 = __ldcg(reinterpret_cast<const uint4*>(e));           // tag, nplus1, used, ids[0]
            if (w0.x == 0) {
                if (*reinterpret_cast<volatile unsigned long long*>(lv.counters + CNT_LISTS) >= lv.max_lists) break;
                const uint32_t old = atomicCAS(&e->tag, 0u, tag);
                if (old == 0) {                                             // this thread publishes the record
                    for (uint32_t i = 0; i < n; i++) e->ids[i] = ids[i * stride] + 1u;
                    *reinterpret_cast<volatile uint32_t*>(&e->nplus1) = n + 1;
                    const unsigned long long at = atomicAdd(&lv.counters[CNT_LISTS], 1ull);
                    lv.occ[at] = slot;                                      // (at < T: every slot is claimed once)
                    return slot;
                }
                w0.x = old;
                w0.y = 0;                                                   // the rest was read before the claim: unknown
            }
            if (w0.x != tag) { slot = (slot + 1) & lv.t_mask; probe++; continue; }
            if (w0.y == 0) continue;                                        // claimed, not yet visible: look again
            if (w0.y != n + 1) { slot = (slot + 1) & lv.t_mask; probe++; continue; }
            int verdict = 1;                                                // 1 equal, 0 different, -1 not yet visible
            if (n > 0) {
                if (w0.w == 0) verdict = -1; else if (w0.w != ids[0] + 1u) verdict = 0;
            }
            for (uint32_t q = 1; verdict == 1 && 4 * q - 3 < n; q++) {      // words ids[4q-3 .. 4q]
                const uint4 w = __ldcg(reinterpret_cast<const uint4*>(e) + q);
                const uint32_t x[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t i = 4 * q - 3 + k;
                    if (i < n && verdict == 1) {
                        if (x[k] == 0) verdict = -1; else if (x[k] != ids[i * stride] + 1u) verdict = 0;
                    }
                }
            }
            if (verdict == 1) return slot;
            if (verdict == 0) { slot = (slot + 1) & lv.t_mask; probe++; }
        }
    }
    return list_overflow(lv, n, ids, stride, spill_off);
}

// ids of a list by handle
struct ListRef {
    const uint32_t* p;
    uint32_t n;
    uint32_t bias;                                          // 1: stored + 1 (inline), 0: raw (spill pool)
    __device__ __forceinline__ uint32_t operator[](uint32_t i) const { return p[i] - bias; }
};
__device__ __forceinline__ ListRef list_ref(const LinkView& lv, uint32_t h) {
    const ListRec* e = lv.recs + h;
    ListRef r;
    r.n = e->nplus1 - 1;
    if (r.n <= (uint32_t)LR_IDS) { r.p = e->ids; r.bias = 1; }
    else { r.p = lv.spill + e->ids[0]; r.bias = 0; }
    return r;
}
#endif

LinkView link_view(Ctx* c);
int link_setup(Ctx* c);                                     // after the index build: size + zero the tables
int link_reset(Ctx* c);                                     // forget every list (vspe_reset)
int link_grow_overflow(Ctx* c);                             // after ERRF_LISTS_FULL / ERRF_SPILL_FULL
// ReadSlots of the slot-writing tiers -> handles: d_handles[i] for slots [0, n), or d_handles[d_scatter[i]] for the compact slots [0, *d_n)
int intern_slots(Ctx* c, const ReadSlot* d_slots, uint64_t n, const uint32_t* d_scatter, const unsigned long long* d_n, uint32_t* d_handles);
// handles -> ReadSlots with ascending ids (vspe_map_reads)
int export_slots(Ctx* c, const uint32_t* d_handles, uint64_t n, ReadSlot* d_slots);
// K5 + K6 over pairs [0, total) of two handle arrays
int count_links(Ctx* c, const uint32_t* d_hf, const uint32_t* d_hr, uint64_t total);
// sparse.cu: sort (keys, vals)[0..n) of c->sparse by key, sum equal keys
int sparse_sort_reduce(Ctx* c, uint64_t n, uint64_t* n_runs_out);
int sparse_reserve(Ctx* c, uint64_t total);

}  // namespace vspe
