// vspe_internal.cuh -- shared layouts and device helpers for libvspe.so (sm_100a only).
//
// Domain vocabulary follows the reference (utils/VStrains_PE_Inference.py): nodes (GFA
// segments), (k+1)-mers of length split_len, postings (node ids per k-mer), read pairs,
// node_mat / short_mat.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/vspe.h"

namespace vspe {

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define VSPE_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            vspe::set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__),       \
                            __FILE__, __LINE__, #call);                                   \
            return VSPE_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define VSPE_TRY(call)                  \
    do {                                \
        int r__ = (call);               \
        if (r__ != VSPE_OK) return r__; \
    } while (0)

// ---------------------------------------------------------------------------------------
// Index layout in HBM (built by K3, read by K4)
//
//   text      2-bit packed bases, base j at bits [2(j&31), 2(j&31)+2) of word j>>5.
//             Node i with len_i >= split_len contributes its forward strand at
//             [strand_start[2i], strand_start[2i+1]) and its reverse complement at
//             [strand_start[2i+1], strand_start[2i+2]).  Shorter nodes contribute nothing.
//             Code = (ascii >> 1) & 3: A=0 C=1 T=2 G=3; complement = code ^ 2.
//   slots     open-addressing (linear probing) table of 8-byte entries {tp, meta}:
//             tp   = text position of one (k+1)-mer occurrence (0xFFFFFFFF = empty)
//             meta = (fingerprint & ~node_mask) | node index
//             Every occurrence of every (k+1)-mer of both strands is its own entry, so the
//             entries whose text equals a query are exactly the reference's postings
//             multiset (PE_Inference.py:117-135; palindromes appear twice because both
//             strands are stored).
//   uniq      1 bit per text position: the (k+1)-mer starting there has exactly one posting.
// ---------------------------------------------------------------------------------------
struct IndexView {
    const uint64_t* text;
    const uint32_t* strand_start;   // [2N+1]
    const uint32_t* node_len;       // [N] full node length (also for nodes without k-mers)
    const uint4* node_rec;          // [N] {strand_start[2n], [2n+1], [2n+2], node_len[n]}: everything a walk needs to enter a node
    const uint2* slots;
    const uint32_t* uniq;           // bitmap over text positions
    const uint32_t* succ;           // [2N][4] {text position, node} of the unique successor k-mer or {NONE32, 0}
    const uint4* succ16;            // [2N][4] the same successor with what a walk needs to enter it: {text position, strand,
                                    // strand end, node length} or {NONE32, 0, 0, 0}
    const uint32_t* bloom;          // blocked Bloom filter over the indexed (k+1)-mers (nullptr when the slot table itself is
                                    // L2 sized): one 32-byte block per hash, 3 bits per k-mer; a clear bit proves a miss
    uint32_t bloom_mask;            // blocks - 1
    const uint32_t* subst;          // 4 bits per text base (or nullptr): bit b set iff some window of the strand that
                                    // covers the base, with the base replaced by code b, has a posting
    uint32_t text_len;              // bases
    uint32_t slot_mask;
    uint32_t node_mask;             // (1 << nbits) - 1
    uint32_t split_len;
    uint32_t n_nodes;
};

static constexpr uint32_t EMPTY_TP = 0xFFFFFFFFu;
static constexpr uint64_t EMPTY_SLOT = 0xFFFFFFFFFFFFFFFFull;
static constexpr uint32_t NONE32 = 0xFFFFFFFFu;

// (k+1)-mer hash: the window is read as 32-bit words of 16 bases (the last one masked to the
// window length); two multiply-add polynomial accumulators, each finished with fmix32.
// The high half selects the slot, the low half is the fingerprint.  Every tier (packed text,
// packed reads, raw ASCII) feeds the same word sequence, so they agree bit for bit.
struct KmerHash {
    uint32_t h1 = 0x243F6A88u, h2 = 0x85A308D3u;
    __host__ __device__ __forceinline__ void add(uint32_t w) {
        h1 = h1 * 0x9E3779B1u + w;
        h2 = h2 * 0x85EBCA77u + w;
    }
    __host__ __device__ static __forceinline__ uint32_t fmix(uint32_t h) {
        h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
        return h;
    }
    __host__ __device__ __forceinline__ uint64_t finish() const {
        return ((uint64_t)fmix(h1) << 32) | fmix(h2 ^ (h1 >> 7));
    }
    // feed up to 32 bases held in a 64-bit word; rem = bases of the window left (>= 1)
    __host__ __device__ __forceinline__ void add64(uint64_t w, uint32_t rem) {
        uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
        if (rem < 16) lo &= (1u << (2 * rem)) - 1;
        add(lo);
        if (rem > 16) {
            if (rem < 32) hi &= (1u << (2 * (rem - 16))) - 1;
            add(hi);
        }
    }
};

// ascii -> 2-bit code; valid only for A,C,G,T (upper case)
__host__ __device__ __forceinline__ uint32_t base_code(uint32_t c) { return (c >> 1) & 3u; }
__host__ __device__ __forceinline__ bool is_acgt(uint32_t c) {
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

#ifdef __CUDACC__
// 32 bases (64 bits) starting at base offset b of a packed array (2 pad words required).
__device__ __forceinline__ uint64_t extract64(const uint64_t* __restrict__ t, uint64_t b) {
    uint64_t w = b >> 5;
    uint32_t s = (uint32_t)(b & 31) * 2;
    uint64_t lo = __ldg(t + w);
    if (s == 0) return lo;
    uint64_t hi = __ldg(t + w + 1);
    return (lo >> s) | (hi << (64 - s));
}

__device__ __forceinline__ uint32_t text_base(const uint64_t* __restrict__ t, uint64_t b) {
    return (uint32_t)(__ldg(t + (b >> 5)) >> ((b & 31) * 2)) & 3u;
}

// hash of the L-base window starting at base b of a packed array
__device__ __forceinline__ uint64_t hash_packed(const uint64_t* __restrict__ t, uint64_t b, uint32_t L) {
    KmerHash hs;
    for (uint32_t m = 0; m < L; m += 32) hs.add64(extract64(t, b + m), L - m);
    return hs.finish();
}

__device__ __forceinline__ uint32_t slot_of(uint64_t h, uint32_t mask) { return (uint32_t)(h >> 32) & mask; }

// Blocked Bloom filter: the block (eight 32-bit words = one 32-byte sector) and the three bit positions
// inside it come from a remix of the hash, independent of the slot index and the fingerprint.
__host__ __device__ __forceinline__ void bloom_bits(uint64_t h, uint32_t mask, uint32_t& block, uint32_t& b0, uint32_t& b1, uint32_t& b2) {
    uint32_t m = KmerHash::fmix((uint32_t)h * 0x9E3779B1u ^ (uint32_t)(h >> 32));
    block = m & mask;
    m = KmerHash::fmix(m + 0x7F4A7C15u);
    b0 = m & 255u; b1 = (m >> 8) & 255u; b2 = (m >> 16) & 255u;
}
// false: no indexed (k+1)-mer has this hash (a proof); true: look it up
__device__ __forceinline__ bool bloom_maybe(const uint32_t* __restrict__ bloom, uint32_t mask, uint64_t h) {
    uint32_t block, b0, b1, b2;
    bloom_bits(h, mask, block, b0, b1, b2);
    const uint4* p = reinterpret_cast<const uint4*>(bloom + 8 * (size_t)block);
    const uint4 lo = __ldg(p), hi = __ldg(p + 1);
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    return ((w[b0 >> 5] >> (b0 & 31)) & (w[b1 >> 5] >> (b1 & 31)) & (w[b2 >> 5] >> (b2 & 31)) & 1u) != 0;
}
__device__ __forceinline__ bool fp_match(uint32_t meta, uint64_t h, uint32_t node_mask) {
    return (((uint32_t)h ^ meta) & ~node_mask) == 0;
}

// are the L-base windows at text positions a and b equal?
__device__ __forceinline__ bool text_equal(const uint64_t* __restrict__ t, uint64_t a, uint64_t b, uint32_t L) {
    for (uint32_t m = 0; m < L; m += 32) {
        uint64_t x = extract64(t, a + m) ^ extract64(t, b + m);
        uint32_t rem = L - m;
        if (rem < 32) x &= (1ull << (2 * rem)) - 1;
        if (x) return false;
    }
    return true;
}

// strand index q (0..2N-1) containing text position tp: largest q with strand_start[q] <= tp
// among strands of non-zero length.
__device__ __forceinline__ uint32_t strand_of(const uint32_t* __restrict__ ss, uint32_t n2, uint32_t tp) {
    uint32_t lo = 0, hi = n2;     // invariant: ss[lo] <= tp < ss[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(ss + mid) <= tp) lo = mid; else hi = mid;
    }
    return lo;
}
#endif

// ---------------------------------------------------------------------------------------
// Per-read result written by K4 and read by K5: 64 bytes.
//   hdr = status | n << 8      status: 0 mapped, 1 has 'N', 2 short
//   n <= SLOT_IDS: ids[0..n) ascending node indices
//   n >  SLOT_IDS: ids[0] = offset into the spill pool where the n ids live
// ---------------------------------------------------------------------------------------
static constexpr int SLOT_IDS = 15;
struct __align__(64) ReadSlot {
    uint32_t hdr;
    uint32_t ids[SLOT_IDS];
};
static constexpr uint32_t ST_OK = 0, ST_N = 1, ST_SHORT = 2;

// ---------------------------------------------------------------------------------------
// Link stage records (link.cuh): one 64-byte record per distinct node list, one 16-byte
// entry per distinct (left list, right list) combination of a batch of read pairs.
// ---------------------------------------------------------------------------------------
static constexpr int LR_IDS = 13;
struct __align__(64) ListRec {
    uint32_t tag;                 // table part: list fingerprint | 1 (0 = free slot)
    uint32_t nplus1;              // ids + 1 (0 = not yet published)
    uint32_t used;                // used pairs of the current batch with this list as a mate (short_mat weight)
    uint32_t ids[LR_IDS];         // n <= LR_IDS: node index + 1; longer: ids[0] = offset into the spill pool
};
struct __align__(16) PairEnt {
    unsigned long long key1;      // ((handle_left << 32) | handle_right) + 1, 0 = free
    uint32_t count;               // pairs of the batch with this combination (node_mat weight)
    uint32_t pad;
};

// ---------------------------------------------------------------------------------------
// simple owning device buffer
// ---------------------------------------------------------------------------------------
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // grow to at least n elements (contents discarded unless keep)
    int reserve(size_t n, bool keep = false, cudaStream_t st = 0) {
        if (n <= cap) return VSPE_OK;
        size_t ncap = n + n / 4 + 64;
        T* q = nullptr;
        VSPE_CUDA(cudaMalloc(&q, ncap * sizeof(T)));
        if (keep && p && cap) {
            VSPE_CUDA(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st));
            VSPE_CUDA(cudaStreamSynchronize(st));
        }
        if (p) cudaFree(p);
        p = q;
        cap = ncap;
        return VSPE_OK;
    }
};

struct Index {
    DevBuf<uint64_t> text;
    DevBuf<uint32_t> strand_start;
    DevBuf<uint32_t> node_len;
    DevBuf<uint4> node_rec;
    DevBuf<uint2> slots;
    DevBuf<uint32_t> uniq;
    DevBuf<uint32_t> succ;
    DevBuf<uint4> succ16;
    DevBuf<uint32_t> bloom;
    uint32_t bloom_mask = 0;
    bool has_bloom = false;
    DevBuf<uint32_t> subst;
    bool has_subst = false;
    uint32_t text_len = 0, slot_mask = 0, node_mask = 0, split_len = 0, n_nodes = 0;
    uint64_t n_kmers = 0;
    bool built = false;
    IndexView view() const {
        IndexView v;
        v.text = text.p; v.strand_start = strand_start.p; v.node_len = node_len.p; v.node_rec = node_rec.p;
        v.slots = slots.p; v.uniq = uniq.p; v.succ = succ.p; v.succ16 = succ16.p; v.bloom = has_bloom ? bloom.p : nullptr; v.bloom_mask = bloom_mask; v.subst = has_subst ? subst.p : nullptr;
        v.text_len = text_len; v.slot_mask = slot_mask; v.node_mask = node_mask;
        v.split_len = split_len; v.n_nodes = n_nodes;
        return v;
    }
};

// per-mate record table produced by K1 for one chunk
struct Records {
    DevBuf<uint64_t> seq_start;   // chunk-relative byte offset of the 2nd line of each record (default path: of each deferred read)
    DevBuf<uint64_t> seq_end;     // one past its last content byte (~0: beyond the scan margin)
    DevBuf<uint32_t> rows;        // [tiles][slots per tile][row_words] 2-bit packed reads written by k_scan_rows
    DevBuf<uint32_t> hdr;         // [tiles][slots per tile] rlen | SH_* flags | first base << 16 (scan_map.cu)
};
static constexpr uint32_t PH_N = 1u << 24, PH_BAD = 2u << 24, PH_LONG = 4u << 24;

struct Ctx;   // defined in api.cu

// ---- kernels' host launchers (each returns VSPE_OK / error and counts its launches) ----
int index_build_device(Ctx* c, const uint8_t* seqs, const uint64_t* seq_off, uint32_t n_nodes, uint32_t split_len);

// K1: count terminators / index records of a device buffer.
int scan_count_lines(Ctx* c, const uint8_t* d_buf, uint64_t n, uint64_t* n_terms);
int scan_index_records(Ctx* c, const uint8_t* d_buf, uint64_t n, uint64_t line_base, uint64_t rec_first,
                       uint64_t n_slots, uint64_t* d_seq_start, uint64_t* d_seq_end);

int scan_records_single_pass(Ctx* c, const uint8_t* d_buf, uint64_t n, uint64_t line_base, uint64_t rec_first,
                             uint64_t n_slots, uint64_t* d_seq_start, uint64_t* d_seq_end, uint64_t* n_terms, bool* overflow);

int device_scan_u64(Ctx* c, const unsigned long long* in, unsigned long long* out, uint64_t n, unsigned long long* sums,
                    unsigned long long* total);

// K1+K2+K4 fused (scan_map.cu): one pass over a device-resident chunk -> handles + the unresolved reads, packed and listed
int scan_map(Ctx* c, int m, const uint8_t* d_buf, uint64_t n, uint64_t line_base, uint64_t rec_first, uint64_t n_slots, uint32_t* d_handles,
             unsigned long long* d_defer_count, uint32_t row_words, uint32_t cap);
const unsigned long long* scan_map_total_ptr(Ctx* c, int m, uint64_t n, const uint8_t* d_buf);
void scan_map_account(Ctx* c);

// K2+K4: map reads [0, n_reads) of a chunk into slots[rec_off + r]
int map_reads_generic(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                      uint64_t n_reads, ReadSlot* d_slots);


int sparse_merge_host(Ctx* c, const uint64_t* keys, const uint64_t* counts, uint64_t n);
int sparse_merge_device(Ctx* c, const uint64_t* d_keys, const uint64_t* d_counts, uint64_t n);
// dense matrices are kept for 2*N*N <= 2^28 cells (2 GiB of uint64); larger graphs count sparsely
inline bool dense_possible(uint64_t n_nodes) { return 2ull * n_nodes * n_nodes <= (8192ull << 15); }

}  // namespace vspe
