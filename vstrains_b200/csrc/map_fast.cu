// map_fast.cu -- K2 + K4, seed-and-extend tier.
//
// Replaces single_end_read_mapping (reference utils/VStrains_PE_Inference.py:16-48) for reads
// whose result it can PROVE; every other read goes to the exhaustive tier (map_generic.cu).
//
// The reference looks up all rlen-split_len+1 windows of a read.  Here a window is looked up
// only when nothing is known about it:
//   * seed:    hash + probe one window; accept only a verified posting whose uniq bit is set
//              (=> that posting is the window's entire postings multiset);
//   * extend:  compare the read against the packed node text 32 bases per step; every further
//              base that matches proves the next window equals the next text window, and its
//              uniq bit proves it has no other posting => one more hit for the same node,
//              no table access;
//   * walk:    at the end of a node strand the successor table (built from the index itself)
//              gives the unique window that continues with the read's next base;
//   * bail:    a verified posting without the uniq bit (repeat / palindrome), a non-ACGT
//              character, more than MAXN distinct nodes or a read longer than the packed
//              capacity sends the read to the exhaustive tier.
// Windows that miss (sequencing errors) are probed one by one, exactly like the reference.
//
// Phase 1 (K2): each warp packs 32 reads to 2 bits/base in shared memory with coalesced loads,
//               flagging 'N' (upper case: PE_Inference.py:160) and other non-ACGT bytes.
// Phase 2 (K4): one thread per read.
#include "ctx.cuh"

namespace vspe {

static constexpr int MF_THREADS = 128;
static constexpr int MAXN = 16;

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
    uint64_t h = HASH_SEED;
    for (uint32_t m = 0; m < L; m += 32) {
        uint64_t w = read64(row, b + m);
        uint32_t rem = L - m;
        if (rem < 32) w &= (1ull << (2 * rem)) - 1;
        h = hash_mix(h, w);
    }
    return hash_final(h);
}

__device__ __forceinline__ bool read_equals_text(const uint32_t* row, uint32_t b, const uint64_t* __restrict__ text,
                                                 uint32_t tp, uint32_t L) {
    for (uint32_t m = 0; m < L; m += 32) {
        uint64_t x = read64(row, b + m) ^ extract64(text, (uint64_t)tp + m);
        uint32_t rem = L - m;
        if (rem < 32) x &= (1ull << (2 * rem)) - 1;
        if (x) return false;
    }
    return true;
}

// number of equal bases of read[rb..] and text[tb..], at most max_ext
__device__ __forceinline__ uint32_t match_len(const uint32_t* row, uint32_t rb, const uint64_t* __restrict__ text,
                                              uint32_t tb, uint32_t max_ext) {
    uint32_t done = 0;
    while (done < max_ext) {
        uint64_t x = read64(row, rb + done) ^ extract64(text, (uint64_t)tb + done);
        if (x) {
            done += (uint32_t)(__ffsll((long long)x) - 1) >> 1;
            break;
        }
        done += 32;
    }
    return min(done, max_ext);
}

// length of the run of set uniq bits starting at text position p, at most n
__device__ __forceinline__ uint32_t uniq_run(const uint32_t* __restrict__ uniq, uint32_t p, uint32_t n) {
    uint32_t done = 0;
    while (done < n) {
        uint32_t q = p + done;
        uint32_t w = ~(__ldg(uniq + (q >> 5)) >> (q & 31));      // zero bits become ones
        uint32_t avail = 32 - (q & 31);
        if (avail < 32) w |= ~0u << avail;                        // bits beyond this word stop the run
        uint32_t run = w ? (uint32_t)(__ffs((int)w) - 1) : 32u;
        done += run;
        if (run < avail) break;
    }
    return min(done, n);
}

enum { PROBE_MISS = 0, PROBE_UNIQUE = 1, PROBE_MULTI = 2 };

template <int STRIDE>
__global__ void __launch_bounds__(MF_THREADS)
k_map_fast(IndexView ix, const uint8_t* __restrict__ buf, const uint64_t* __restrict__ seq_start,
           const uint64_t* __restrict__ seq_end, uint64_t n_reads, ReadSlot* __restrict__ slots,
           uint32_t* __restrict__ worklist, unsigned long long* __restrict__ counters) {
    constexpr uint32_t CAP = (STRIDE - 3) * 16;                   // bases per packed row
    __shared__ uint32_t s_read[MF_THREADS * STRIDE];
    __shared__ uint32_t s_node[MAXN][MF_THREADS];
    __shared__ uint32_t s_vk[MAXN][MF_THREADS];                    // v | kmin << 16
    __shared__ uint32_t s_len[MF_THREADS];                         // rlen | flags << 24
    constexpr uint32_t F_N = 1u << 24, F_BAD = 2u << 24, F_LONG = 4u << 24, F_NONE = 8u << 24;
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t r0 = (uint64_t)blockIdx.x * MF_THREADS;
    const uint32_t L = ix.split_len;

    // ---- phase 1: cooperative pack --------------------------------------------------------
    for (uint32_t k = 0; k < 32; k++) {
        const uint32_t t = wib * 32 + k;
        const uint64_t r = r0 + t;
        uint32_t* row = s_read + t * STRIDE;
        if (r >= n_reads) {
            if (lane == 0) s_len[t] = F_NONE;
            continue;
        }
        const uint64_t s = seq_start[r], e = seq_end[r];
        const uint64_t len64 = e - s;
        if (len64 > CAP) {
            // too long for the packed row: still need the 'N' test before the exhaustive tier
            if (lane == 0) s_len[t] = F_LONG;
            continue;
        }
        const uint32_t rlen = (uint32_t)len64;
        const uint32_t nwords = (rlen + 15) >> 4;
        bool anyN = false, anyBad = false;
        for (uint32_t w0 = 0; w0 < nwords; w0 += 32) {
            const uint32_t w = w0 + lane;
            uint32_t packed = 0;
            bool hasN = false, bad = false;
            if (w < nwords) {
                // first byte of this lane's 16 bases, as an ABSOLUTE address (shards may be misaligned)
                const uintptr_t a = reinterpret_cast<uintptr_t>(buf) + s + 16ull * w;
                const uint32_t nb = min(16u, rlen - 16 * w);
                const uint32_t sh = (uint32_t)(a & 3) * 8;
                const uint32_t* p = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
                // 5 aligned words cover 16 unaligned bytes; words entirely past the read are not touched
                uint32_t x[5];
                const uint32_t need = ((uint32_t)(a & 3) + nb + 3) >> 2;
#pragma unroll
                for (int j = 0; j < 5; j++) x[j] = (uint32_t)j < need ? __ldg(p + j) : 0u;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t c = __funnelshift_r(x[j], x[j + 1], sh);
                    const int left = (int)nb - 4 * j;               // valid bytes in this word
                    if (left <= 0) break;
                    const uint32_t vm = left >= 4 ? 0xFFFFFFFFu : ((1u << (8 * left)) - 1);
                    uint32_t ok = __vcmpeq4(c, 0x41414141u) | __vcmpeq4(c, 0x43434343u) |
                                  __vcmpeq4(c, 0x47474747u) | __vcmpeq4(c, 0x54545454u);
                    uint32_t isn = __vcmpeq4(c, 0x4E4E4E4Eu);
                    hasN |= (isn & vm) != 0;
                    bad |= (~(ok | isn) & vm) != 0;
                    uint32_t code = ((c & vm) >> 1) & 0x03030303u;
                    packed |= ((code * 0x01041040u) >> 24) << (8 * j);
                }
            }
            if (w < (uint32_t)STRIDE) row[w] = packed;
            anyN |= hasN;
            anyBad |= bad;
        }
        anyN = __any_sync(0xFFFFFFFFu, anyN);
        anyBad = __any_sync(0xFFFFFFFFu, anyBad);
        // zero the pad words read64 may touch
        for (uint32_t w = nwords + lane; w < (uint32_t)STRIDE; w += 32) row[w] = 0;
        if (lane == 0) s_len[t] = rlen | (anyN ? F_N : 0) | (anyBad ? F_BAD : 0);
    }
    __syncthreads();

    // ---- phase 2: one thread per read -------------------------------------------------------
    const uint32_t t = threadIdx.x;
    const uint64_t r = r0 + t;
    const uint32_t lf = s_len[t];
    if (lf & F_NONE) return;
    if (t == 0 && blockIdx.x == 0) atomicAdd(&counters[CNT_FAST], (unsigned long long)n_reads);
    ReadSlot* out = slots + r;
    bool bail = (lf & (F_LONG | F_BAD)) != 0;
    const uint32_t rlen = lf & 0xFFFFFF;
    if (!(lf & F_LONG)) {
        if (lf & F_N) { out->hdr = ST_N; return; }
        if (rlen < L) { out->hdr = ST_SHORT; return; }
    }
    const uint32_t* row = s_read + t * STRIDE;
    uint32_t nn = 0;
    if (!bail) {
        const uint32_t npos = rlen - L + 1;
        uint32_t i = 0;
        uint32_t walk_tp = NONE32;          // window i is already known to equal text window walk_tp (unique)
        while (i < npos && !bail) {
            uint32_t tp = walk_tp, node = 0;
            walk_tp = NONE32;
            if (tp == NONE32) {
                // ---- seed: hash + probe window i ----
                const uint64_t h = hash_read(row, i, L);
                uint32_t j = slot_of(h, ix.slot_mask);
                int res = PROBE_MISS;
                while (true) {
                    const uint2 ent = __ldg(ix.slots + j);
                    if (ent.x == EMPTY_TP) break;
                    if (fp_match(ent.y, h, ix.node_mask) && read_equals_text(row, i, ix.text, ent.x, L)) {
                        const bool u = (__ldg(ix.uniq + (ent.x >> 5)) >> (ent.x & 31)) & 1;
                        res = u ? PROBE_UNIQUE : PROBE_MULTI;
                        tp = ent.x;
                        node = ent.y & ix.node_mask;
                        break;
                    }
                    j = (j + 1) & ix.slot_mask;
                }
                if (res == PROBE_MULTI) { bail = true; break; }
                if (res == PROBE_MISS) { i++; continue; }
            } else {
                // node of a successor window: binary search over strand starts (small, cached)
                node = strand_of(ix.strand_start, 2 * ix.n_nodes, tp) >> 1;
            }
            // ---- extend along the node strand ----
            const uint32_t s0 = __ldg(ix.strand_start + 2 * node), s1 = __ldg(ix.strand_start + 2 * node + 1);
            const uint32_t q = 2 * node + (tp >= s1 ? 1u : 0u);
            const uint32_t send = tp >= s1 ? __ldg(ix.strand_start + 2 * node + 2) : s1;
            (void)s0;
            const uint32_t room_t = send - (tp + L), room_r = rlen - (i + L);
            const uint32_t max_ext = min(room_t, room_r);
            uint32_t ext = match_len(row, i + L, ix.text, tp + L, max_ext);
            if (ext) {
                const uint32_t ur = uniq_run(ix.uniq, tp + 1, ext);
                if (ur < ext) { bail = true; break; }      // a matching window with several postings
            }
            const uint32_t hits = 1 + ext;
            // ---- accumulate (node, hits, first position) ----
            {
                uint32_t a = 0;
                for (; a < nn; a++) if (s_node[a][t] == node) break;
                if (a == nn) {
                    if (nn == MAXN) { bail = true; break; }
                    s_node[a][t] = node;
                    s_vk[a][t] = hits | (i << 16);
                    nn++;
                } else {
                    s_vk[a][t] += hits;                        // kmin keeps the first (smallest) position
                }
            }
            i += hits;
            // ---- walk to the successor when the strand ended exactly here ----
            if (ext == room_t && i < npos) {
                const uint32_t nb = i + L - 1;                 // the one new base of window i
                const uint32_t b = (row[nb >> 4] >> ((nb & 15) * 2)) & 3u;
                walk_tp = __ldg(ix.succ + 4 * q + b);
            }
        }
    }
    if (bail) {
        const unsigned long long idx = atomicAdd(&counters[CNT_WORK], 1ull);
        worklist[idx] = (uint32_t)(r);
        return;
    }
    // ---- sort by node index (ascending, as enumerate(nodes) does) and apply the predicate ----
    for (uint32_t a = 1; a < nn; a++) {
        uint32_t kn = s_node[a][t], kv = s_vk[a][t];
        int b = (int)a - 1;
        while (b >= 0 && s_node[b][t] > kn) {
            s_node[b + 1][t] = s_node[b][t];
            s_vk[b + 1][t] = s_vk[b][t];
            b--;
        }
        s_node[b + 1][t] = kn;
        s_vk[b + 1][t] = kv;
    }
    uint32_t n_out = 0;
    for (uint32_t a = 0; a < nn; a++) {
        const uint32_t node = s_node[a][t], vk = s_vk[a][t];
        if (keep_node_f(vk & 0xFFFF, vk >> 16, __ldg(ix.node_len + node), rlen, L)) {
            if (n_out < (uint32_t)SLOT_IDS) out->ids[n_out] = node;
            n_out++;
        }
    }
    if (n_out > (uint32_t)SLOT_IDS) {
        // MAXN == 16 and SLOT_IDS == 15: a 16-node list does not fit the slot -> exhaustive tier
        const unsigned long long idx = atomicAdd(&counters[CNT_WORK], 1ull);
        worklist[idx] = (uint32_t)(r);
        return;
    }
    out->hdr = ST_OK | (n_out << 8);
}

int map_reads_generic_dev(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                          const uint32_t* d_worklist, const unsigned long long* d_n_items, ReadSlot* d_slots);

int map_reads_fast(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                   uint64_t n_reads, ReadSlot* d_slots) {
    if (n_reads == 0) return VSPE_OK;
    if (n_reads > 0xFFFFFFFFull) { set_error("more than 2^32 reads in one chunk"); return VSPE_ERR_LIMIT; }
    const uint32_t L = c->index.split_len;
    if (L > 0xFFFF) return map_reads_generic(c, d_buf, d_seq_start, d_seq_end, n_reads, d_slots);
    VSPE_TRY(c->worklist.reserve(n_reads));
    VSPE_CUDA(cudaMemsetAsync(c->counters.p + CNT_WORK, 0, 8, c->stream));
    const uint32_t grid = (uint32_t)((n_reads + MF_THREADS - 1) / MF_THREADS);
    // packed-row capacity 320 bases (covers 2x150, 2x250 and 2x300 runs); longer reads bail
    k_map_fast<23><<<grid, MF_THREADS, 0, c->stream>>>(c->index.view(), d_buf, d_seq_start, d_seq_end, n_reads, d_slots,
                                                       c->worklist.p, c->counters.p);
    VSPE_LAUNCH_CHECK(c);
    // the exhaustive tier consumes the worklist; its length stays on the device
    VSPE_TRY(map_reads_generic_dev(c, d_buf, d_seq_start, d_seq_end, c->worklist.p, c->counters.p + CNT_WORK, d_slots));
    return VSPE_OK;
}

}  // namespace vspe
