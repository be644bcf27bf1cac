// scan_map.cu -- K1 + K2 (k_scan_rows) and the first tier of K4 (k_walk): the default path of a
// FASTQ chunk from raw bytes to node-list handles.
//
// Replaces `readlines()` + `[s[:-1] ...]` (reference utils/VStrains_PE_Inference.py:149-159), the
// per-character work of `fseq.count("N")` / k-mer slicing (:160, :25) and single_end_read_mapping
// (:16-48) for every read whose result the walk can PROVE; the rest (about 2 % on the bench
// workloads: two or more sequencing errors, repeats, non-ACGT characters, very long reads) is
// listed for the list-driven tiers of map_fast.cu / map_generic.cu.
//
// k_scan_rows -- ONE pass over the bytes, one 40 KiB tile per CTA (10 warps, 4 CTAs per SM); NO tile waits on
// another one:
//   1. one elected thread issues TMA bulk copies (cp.async.bulk, mbarrier complete_tx) of the tile
//      + a 16-byte front margin + a 512-byte back margin into shared memory;
//   2. every lane tests its 16-byte vectors for bytes < 0x10 or >= 0x80 (two instructions per
//      32-bit word); candidate vectors go to a per-warp queue (warp ballots) and only they get exact
//      terminator masks (universal newlines: '\n', "\r\n" once, lone '\r'); a warp scan ranks them and
//      every terminator's position is stored at its rank; warp totals + one block exchange merge the
//      warps' tables into the tile's terminator table and give the tile's terminator count;
//   3. which of the tile's lines are sequence lines depends on the number of lines before the tile
//      (mod 4).  Instead of waiting for the earlier tiles (a decoupled look-back was 40 % of the
//      kernel's warp time, profiles/r02_ncu_c4_block_before.txt) the tile GUESSES that phase from its own
//      bytes ('@' / '+' at the starts of its first 32 lines), packs its reads into TILE-LOCAL slots
//      and records {terminator count, guessed phase};
//   4. one thread per read (two for rows of 16 / 20 words) packs the read the tile owns (its sequence line STARTS
//      here) to 2 bits/base straight from the tile, 16 bases per step (SIMD-in-word ACGT validity test, 'N' flag), and
//      stores the row from registers (48 / 64 / 80 bytes per read + a header word), streaming.
// k_tile_sum / k_tile_fix -- exact prefix sum of the tiles' terminator counts -> first read of every tile,
//      first tile of every k_walk block, terminator total; every guess is CHECKED against the exact
//      phase and a tile that guessed wrong (text that merely looks like FASTQ structure) is listed;
// k_scan_redo -- the listed tiles again with their exact phase (normally none): the results never
//      depend on a guess.
// k_memo -- one thread per read: finds the read's tile slot, settles 'N' / short reads, sends rows that are not plain
//      ACGT to the list-driven tiers, asks the READ MEMO (packed row -> list handle of an identical read walked before;
//      exact: the whole row is compared) and lists what is left for k_walk.
// k_walk -- list-driven, one thread per read, 128-thread blocks, 10 blocks per SM (the walk is a chain of dependent
// L2 accesses: it wants many warps and a large L1, which is why it is NOT fused into the scan kernel --
// the fused variant was built, bit-exact, and 3x slower: 15 warps per SM, 31 % issue slots):
//   6. seed window 0 (hash + probe + verify + uniq bit; on a miss the reverse complement is seeded from
//      the other end), then a flat loop whose every turn compares 32 bases + 32 uniq bits on the current
//      diagonal and, when the stretch is complete, books it and steps to the successor strand, whose
//      table entry (text position, strand, strand end, node length: 16 bytes) was requested when the
//      stretch was entered.  ONE mismatching base is tolerated when the substitution-hit bit proves that
//      the windows covering it have no posting.  The saturation predicate (:36-47, integer form) is
//      applied as each stretch is booked;
//   7. the kept node list is interned (link.cuh) and its handle stored (and entered into the memo if the read had no
//      error); an unresolved read is copied (row, header, byte range) to the compact arrays the list-driven tiers work on.
// Why this is exact: see map_fast.cu (a window is counted without a table access only if its text
// equality and the uniq bit of that text window were both checked; it is skipped only if the
// index build already looked that k-mer up and found nothing).
#include <algorithm>

#include "link.cuh"
#include "map_common.cuh"

namespace vspe {

#ifndef VSPE_SM_WARPS
#define VSPE_SM_WARPS 10
#endif
#ifndef VSPE_SM_ITERS
#define VSPE_SM_ITERS 8
#endif
#ifndef VSPE_SM_MINB
#define VSPE_SM_MINB 4
#endif
static constexpr int SM_WARPS = VSPE_SM_WARPS;
static constexpr int SM_THREADS = SM_WARPS * 32;
static constexpr int SM_ITERS = VSPE_SM_ITERS;                   // 512-byte warp rows per warp
static constexpr int SM_TILE = SM_WARPS * SM_ITERS * 32 * 16;    // 40 KiB
static constexpr int SM_FRONT = 16;                              // bytes kept before the tile
static constexpr int SM_BACK = 512;                              // bytes kept after the tile (>= longest packed read + 1)
static constexpr int SM_MAXREC = SM_TILE / 64 - 64;              // reads a tile may own (else the chunk takes the plain path): 576
static constexpr int SM_QCAP = 12 * SM_ITERS;                    // per warp: vectors that may hold a terminator
static constexpr int WK_THREADS = 128;                           // k_walk: reads per block
static constexpr int SM_MAXST = 16;                              // stretches (nodes with hits) per read in the walk

#define LB_AGG (1ull << 62)
#define LB_INC (2ull << 62)
#define LB_VAL ((1ull << 62) - 1)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t movemask4b(uint32_t cmp) { return ((cmp & 0x80808080u) * 0x00204081u) >> 28; }

// terminator mask of one 16-byte vector held in registers (universal newlines: '\n', '\r' unless a '\n' follows);
// `valid` = bitmask of the bytes that belong to the buffer; next_byte = the byte after the vector (0 if outside)
__device__ __forceinline__ void masks_from_vec(uint4 v, uint32_t valid, uint32_t next_byte, bool& non_ascii, uint32_t& term) {
    term = 0;
    if (valid == 0xFFFFu && ((v.x | v.y | v.z | v.w) & 0x80808080u)) non_ascii = true;
    uint32_t nl = movemask4b(__vcmpeq4(v.x, 0x0A0A0A0Au)) | (movemask4b(__vcmpeq4(v.y, 0x0A0A0A0Au)) << 4) |
                  (movemask4b(__vcmpeq4(v.z, 0x0A0A0A0Au)) << 8) | (movemask4b(__vcmpeq4(v.w, 0x0A0A0A0Au)) << 12);
    // '\r' is rare: one zero-byte test over the four words decides whether its mask is needed at all
    const uint32_t zx = v.x ^ 0x0D0D0D0Du, zy = v.y ^ 0x0D0D0D0Du, zz = v.z ^ 0x0D0D0D0Du, zw = v.w ^ 0x0D0D0D0Du;
    uint32_t cr = 0;
    if ((((zx - 0x01010101u) & ~zx) | ((zy - 0x01010101u) & ~zy) | ((zz - 0x01010101u) & ~zz) | ((zw - 0x01010101u) & ~zw)) & 0x80808080u)
        cr = movemask4b(__vcmpeq4(v.x, 0x0D0D0D0Du)) | (movemask4b(__vcmpeq4(v.y, 0x0D0D0D0Du)) << 4) |
             (movemask4b(__vcmpeq4(v.z, 0x0D0D0D0Du)) << 8) | (movemask4b(__vcmpeq4(v.w, 0x0D0D0D0Du)) << 12);
    if (valid != 0xFFFFu) {
        const uint32_t na = movemask4b(v.x) | (movemask4b(v.y) << 4) | (movemask4b(v.z) << 8) | (movemask4b(v.w) << 12);
        if (na & valid) non_ascii = true;
        nl &= valid;
        cr &= valid;
    }
    term = nl | (cr & ~((nl >> 1) | (next_byte == '\n' ? 0x8000u : 0u)));
}

struct ScanMapArgs {
    const uint8_t* buf;              // chunk start (may be misaligned)
    uint64_t n;                      // chunk bytes
    uint32_t head;                   // address of buf mod 16
    uint32_t n_tiles;
    uint64_t line_base;              // lines before this chunk
    unsigned long long* tile_info;   // [n_tiles] out: terminators | phase << 32 | overflow << 40
    const unsigned long long* redo;  // list-driven launch: [*n_redo] tile | exact phase << 32
    const unsigned long long* n_redo;
    uint32_t* rows;                  // [n_tiles * tcap][row_words] packed reads, tile-local slots
    uint32_t* hdr;                   // [n_tiles * tcap] rlen | SH_* flags | first base (tile-relative) << 16
    uint32_t tcap;                   // slots per tile
    uint32_t cap;                    // longest read (bases) a packed row holds
    unsigned long long* counters;
};
// header word of a tile-local slot
static constexpr uint32_t SH_RLEN = 0xFFFu, SH_N = 1u << 12, SH_BAD = 1u << 13, SH_LONG = 1u << 14;

// The walk of step 6.  row: this thread's packed read (STRIDE words, zero padded); lst: its node list
// column (entry i at lst[i * LS]).  Returns true when every window of the read is accounted for;
// n_kept nodes that pass the saturation predicate are then at the front of the column.
template <int STRIDE, int LS>
__device__ __forceinline__ bool walk_read(const IndexView& ix, uint32_t* row, const uint32_t rlen, uint32_t* lst, uint32_t& n_kept, bool& clean) {
    constexpr int NW = STRIDE - 3;
    const uint32_t L = ix.split_len;
    const int npos = (int)(rlen - L + 1);
    bool mirror = false;
    uint32_t tp = NONE32, node = 0;
    int pr = probe_window(ix, row, 0, tp, node);
    if (pr == PROBE_MISS && ix.subst != nullptr) {
        // error in the first split_len bases: seed window 0 of the reverse complement (the other end)
        revcomp_row<NW>(row, rlen);
        mirror = true;
        pr = probe_window(ix, row, 0, tp, node);
    }
    // (PROBE_MULTI, or both ends miss: a real complication, the next tier decides)
    bool running = pr == PROBE_UNIQUE, resolved = false;
    bool err = false;
    int e = 0;
    uint32_t rb = 0;
    uint32_t i0 = 0, p = L, q = 0, lim = 0, nn = 0, n_front = 0, nlen = 0;
    unsigned long long seen = 0, seen2 = 0;                    // two 64-bit filters over the nodes booked so far
    int delta = 0;
    uint4 nxt = make_uint4(NONE32, 0, 0, 0);                   // successor entry of the current strand for the read's base at lim
    // enter the strand q that holds text position tp_ (its end: send_, its node length: nlen_) at window i0;
    // false if the error's windows cannot be proven to miss there
    auto enter = [&](uint32_t tp_, uint32_t q_, uint32_t send_, uint32_t nlen_) -> bool {
        q = q_;
        node = q_ >> 1;
        nlen = nlen_;
        delta = (int)tp_ - (int)i0;
        lim = min(rlen, (uint32_t)((int)send_ - delta));       // read position where the strand ends
        if (lim < rlen) {
            // the strand ends before the read: request the successor entry for the read's next base NOW, it is
            // consumed when the stretch is booked (the chunk compares in between hide the access)
            const uint32_t b = (row[lim >> 4] >> ((lim & 15) * 2)) & 3u;
            nxt = __ldg(ix.succ16 + 4 * (size_t)q + b);
        }
        // a strand entered after the error still holds windows covering it if it starts at or before e
        if (err && (int)i0 <= e) {
            const uint32_t te = (uint32_t)(e + delta);
            if ((__ldg(ix.subst + (te >> 3)) >> (4 * (te & 7) + rb)) & 1u) return false;
        }
        return true;
    };
    if (running) {
        const uint4 nr = __ldg(ix.node_rec + node);            // {forward start, rc start, end, node length}
        const bool rcs = tp >= nr.y;
        if (!enter(tp, 2 * node + (rcs ? 1u : 0u), rcs ? nr.z : nr.y, nr.w)) running = false;
    }
    while (running) {
        if (p < lim) {
            const uint32_t n = min(32u, lim - p);
            uint64_t x = read64(row, p) ^ extract64(ix.text, (uint64_t)((int)p + delta));
            if (n < 32) x &= (1ull << (2 * n)) - 1;
            bool ok = true;
            if (x) {
                const uint32_t off = (uint32_t)(__ffsll((long long)x) - 1) >> 1;
                if (err || ix.subst == nullptr || (x & ~(3ull << (2 * off)))) ok = false;      // second mismatch
                else {
                    err = true;
                    e = (int)(p + off);
                    rb = (row[(uint32_t)e >> 4] >> (((uint32_t)e & 15) * 2)) & 3u;
                    const uint32_t te = (uint32_t)(e + delta);
                    if ((__ldg(ix.subst + (te >> 3)) >> (4 * (te & 7) + rb)) & 1u) ok = false;
                }
            }
            const uint32_t u = (uint32_t)((int)p + delta) - L + 1;     // text position of the first window ending here
            const uint32_t ub = __funnelshift_r(__ldg(ix.uniq + (u >> 5)), __ldg(ix.uniq + (u >> 5) + 1), u & 31);
            const uint32_t m32 = n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1);
            if ((ub & m32) != m32) ok = false;
            if (!ok) { running = false; break; }
            p += n;
        }
        if (p >= lim) {
            // windows [a, bw] of this node are resolved; those covering e are proven misses
            const int a = (int)i0, bw = (int)lim - (int)L;
            int c1 = bw - a + 1, c2 = 0, last_hit = bw, first_hit = a;
            if (err) {
                c1 = min(bw, e - (int)L) - a + 1;
                if (c1 < 0) c1 = 0;
                const int a2 = max(a, e + 1);
                c2 = bw - a2 + 1;
                if (c2 < 0) c2 = 0;
                first_hit = c1 > 0 ? a : a2;
                last_hit = c2 > 0 ? bw : min(bw, e - (int)L);
            }
            if (c1 + c2 > 0) {
                if (nn == (uint32_t)SM_MAXST) { running = false; break; }
                // a node met twice (cyclic graph) needs its hits merged by the next tier: two 64-bit filters gate
                // the exact comparison with the nodes booked so far
                const uint32_t hb = (node * 0x9E3779B1u) >> 26, hb2 = (node * 0x85EBCA77u) >> 26;
                if (((seen >> hb) & (seen2 >> hb2)) & 1ull) {
                    bool dup = false;
                    for (uint32_t i = 0; i < n_front; i++) dup |= lst[i * LS] == node;
                    for (uint32_t i = 0; i < nn - n_front; i++) dup |= lst[(SM_MAXST - 1 - i) * LS] == node;
                    if (dup) { running = false; break; }
                }
                seen |= 1ull << hb;
                seen2 |= 1ull << hb2;
                const uint32_t v = (uint32_t)(c1 + c2), kmin = (uint32_t)(mirror ? npos - 1 - last_hit : first_hit);
                // kept nodes fill the column from the front, the others from the back
                if (keep_node_f(v, kmin, nlen, rlen, L)) lst[(n_front++) * LS] = node;
                else lst[(SM_MAXST - 1 - (nn - n_front)) * LS] = node;
                nn++;
            }
            if (lim >= rlen) { resolved = true; running = false; break; }
            if (nxt.x == NONE32) { running = false; break; }       // no unique successor window for that base
            i0 = lim - L + 1;
            p = lim + 1;
            if (!enter(nxt.x, nxt.y, nxt.z, nxt.w)) { running = false; break; }
        }
    }
    if (!resolved) return false;
    n_kept = n_front;
    clean = !err;                                              // every base of the read equals the graph's text
    return true;
}

// ---------------------------------------------------------------------------------------------
// Read memo.  Deep sequencing (what a strain-level viral graph is built from) repeats reads: every read
// without a sequencing error is a substring of a strain, and there are only about 2 x genome length x strains
// distinct ones, each seen coverage / read length times.  The memo is a hash table packed row -> list handle,
// filled by k_walk with reads it resolved WITHOUT tolerating an error (so it stays bounded by the graph, not by
// the input) and asked by k_memo before any walk.  A hit is exact: the whole row and its length are compared,
// and the handle is the one intern_list gave that same sequence.  Entry = 4 + row_words words:
//   [0] tag (hash high half | 1; 0 = free)   [1] handle + 1 (0 = not yet published)   [2] rlen   [4..] row
// ---------------------------------------------------------------------------------------------
static constexpr int MEMO_PROBES = 4;
static constexpr uint32_t MEMO_ENTRIES = 1u << 20;
struct MemoView {
    uint32_t* tab;                     // nullptr: memo off
    uint32_t mask;                     // entries - 1
    uint32_t stride;                   // words per entry
};

template <int XW>
__device__ __forceinline__ uint64_t memo_hash(const uint32_t (&x)[XW], uint32_t rlen) {
    KmerHash hs;
#pragma unroll
    for (int w = 0; w < XW; w++) hs.add(x[w]);
    hs.add(rlen);
    return hs.finish();
}

// handle of an equal read, or H_PENDING.  The whole entry is requested at once (the lookup is one round trip to L2 / HBM
// in the common case, not header-then-row).
template <int XW>
__device__ __forceinline__ uint32_t memo_find(const MemoView& mv, const uint32_t (&x)[XW], uint32_t rlen, uint64_t h) {
    const uint32_t tag = (uint32_t)(h >> 32) | 1u;
    uint32_t slot = (uint32_t)h & mv.mask;
#pragma unroll 1
    for (int probe = 0; probe < MEMO_PROBES; probe++, slot = (slot + 1) & mv.mask) {
        const uint4* e = reinterpret_cast<const uint4*>(mv.tab + (size_t)slot * mv.stride);
        const uint4 w0 = __ldcg(e);                                             // tag, handle + 1, rlen, -
        uint4 w[XW / 4];
#pragma unroll
        for (int q = 0; q < XW / 4; q++) w[q] = __ldcg(e + 1 + q);
        if (w0.x == 0) return H_PENDING;
        if (w0.x != tag || w0.y == 0 || w0.z != rlen) continue;
        bool eq = true;
#pragma unroll
        for (int q = 0; q < XW / 4; q++) eq &= w[q].x == x[4 * q] && w[q].y == x[4 * q + 1] && w[q].z == x[4 * q + 2] && w[q].w == x[4 * q + 3];
        if (eq) return w0.y - 1;
        // (a row read while its writer had not finished differs from x: the read is walked, which is always right)
    }
    return H_PENDING;
}

template <int XW>
__device__ __forceinline__ void memo_insert(const MemoView& mv, const uint32_t (&x)[XW], uint32_t rlen, uint64_t h, uint32_t handle) {
    const uint32_t tag = (uint32_t)(h >> 32) | 1u;
    uint32_t slot = (uint32_t)h & mv.mask;
#pragma unroll 1
    for (int probe = 0; probe < MEMO_PROBES; probe++, slot = (slot + 1) & mv.mask) {
        uint32_t* e = mv.tab + (size_t)slot * mv.stride;
        uint32_t t = *reinterpret_cast<volatile uint32_t*>(e);
        if (t == 0) t = atomicCAS(e, 0u, tag);
        if (t == 0) {                                                           // this thread owns the entry
            e[2] = rlen;
#pragma unroll
            for (int q = 0; q < XW / 4; q++)
                reinterpret_cast<uint4*>(e)[1 + q] = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(e + 1) = handle + 1;         // publish
            return;
        }
        if (t == tag) return;                                                   // (very likely) the same read, entered by someone else
    }
}

// The per-read arrays shared by k_memo and k_walk.
struct WalkArgs {
    const uint32_t* rows;              // tile-local slots written by k_scan_rows
    const uint32_t* hdr;
    uint32_t row_words, tcap;
    const uint32_t* r_first;           // [n_tiles + 1] chunk-local index of each tile's first read (k_tile_fix)
    const uint32_t* blk_tile;          // [blocks of WK_THREADS reads] tile of the block's first read
    uint32_t head;                     // chunk start address mod 16 (tile t starts at buffer position t * SM_TILE - head)
    uint64_t n_slots;                  // capacity of the per-read arrays
    uint64_t r_lo, r_hi;               // reads of this k_memo / k_walk round (the memo a round fills serves the next ones)
    const unsigned long long* total;   // terminators of the chunk (written by k_tile_fix)
    uint64_t line_base, rec_first;
    uint32_t split_len;
    uint32_t* handles;
    uint2* walk_list;                  // reads the walk has to do: {chunk-local read index, slot}
    unsigned long long* walk_count;
    uint32_t* d_read;                  // deferred reads, compact: chunk-local read index,
    uint32_t* d_hdr;                   //   rlen | PH_* flags,
    uint32_t* d_rows;                  //   packed row,
    uint64_t* d_start;                 //   chunk-relative byte range of the sequence line (~0: end not seen)
    uint64_t* d_end;
    unsigned long long* defer_count;
    unsigned long long* counters;
    MemoView memo;
};

// a read no tier before the list-driven ones can finish: row, header and byte range go to their compact arrays
__device__ __forceinline__ void defer_read(const WalkArgs& a, unsigned long long d, uint32_t r, uint64_t slot, uint32_t h) {
    const uint32_t rlen = h & SH_RLEN;
    a.d_read[d] = r;
    a.d_hdr[d] = rlen | ((h & SH_N) ? PH_N : 0u) | ((h & SH_BAD) ? PH_BAD : 0u) | ((h & SH_LONG) ? PH_LONG : 0u);
    const uint64_t st = (slot / a.tcap) * SM_TILE + (h >> 16) - a.head;
    a.d_start[d] = st;
    a.d_end[d] = (h & SH_LONG) ? ~0ull : st + rlen;
    const uint4* src = reinterpret_cast<const uint4*>(a.rows + slot * a.row_words);
    uint4* dst = reinterpret_cast<uint4*>(a.d_rows + d * a.row_words);
    for (uint32_t q = 0; 4 * q < a.row_words; q++) dst[q] = __ldg(src + q);
}

// positions [base, base + n) of a device-side list for the lanes of this warp that want one (one atomic per warp)
__device__ __forceinline__ unsigned long long warp_append(unsigned long long* count, bool want) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, want), lane = threadIdx.x & 31;
    if (m == 0) return 0;
    unsigned long long base = 0;
    const int leader = __ffs((int)m) - 1;
    if ((int)lane == leader) base = atomicAdd(count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + __popc(m & ((1u << lane) - 1));
}

// k_memo -- one thread per read of the chunk: finds the read's slot, settles the pair-skipping classes ('N', short),
// sends reads that are not plain ACGT rows to the list-driven tiers, asks the memo, and lists what is left for k_walk
// (one list append per block).
static constexpr int MM_THREADS = 256;
static_assert(MM_THREADS % WK_THREADS == 0, "blk_tile is indexed by blocks of WK_THREADS reads");

// (8 blocks per SM = 32 registers: the kernel is a chain of memory round trips, occupancy is what it needs;
// measured 0.65 -> 0.59 ms per 10 M reads against the unconstrained 44 registers)
#ifndef VSPE_MM_MINB
#define VSPE_MM_MINB 8
#endif
template <int RW>
__global__ void __launch_bounds__(MM_THREADS, VSPE_MM_MINB)
k_memo(const WalkArgs a) {
    __shared__ uint32_t s_cnt[MM_THREADS / 32], s_hit[MM_THREADS / 32];
    __shared__ unsigned long long s_base;
    // sequence lines of the chunk = #{l in [line_base, line_base + total) : l % 4 == 1}
    const uint64_t n_lines1 = (uint64_t)(a.line_base + *a.total + 2) / 4 - a.rec_first;
    const uint64_t n_reads = n_lines1 < a.n_slots ? n_lines1 : a.n_slots;
    const uint64_t r = a.r_lo + (uint64_t)blockIdx.x * MM_THREADS + threadIdx.x;
    // a tile overflowed: the host repeats the chunk on the plain path, nothing of this launch is used
    const bool live = r < n_reads && r < a.r_hi && !(*reinterpret_cast<const volatile unsigned long long*>(a.counters + CNT_ERR) & ERRF_TILE_FULL);
    if (r == 0 && live) atomicAdd(&a.counters[CNT_FAST], (unsigned long long)n_reads);
    bool walk = false, hit = false;
    uint64_t slot = 0;
    if (live) {
        // the read's slot: its tile (the first tile of its 128-read block, or one of the next ones) and its index there
        uint32_t t = __ldg(a.blk_tile + (r / WK_THREADS));
        uint32_t f0 = __ldg(a.r_first + t), f1 = __ldg(a.r_first + t + 1);
        while ((uint32_t)r >= f1) { t++; f0 = f1; f1 = __ldg(a.r_first + t + 1); }
        slot = (uint64_t)t * a.tcap + ((uint32_t)r - f0);
        // header and row are requested together (the row of a read that turns out to be 'N' / short / long is simply not used)
        const uint32_t h = __ldcs(a.hdr + slot);
        uint32_t x[RW];
        const uint4* src = reinterpret_cast<const uint4*>(a.rows + slot * RW);
#pragma unroll
        for (int q = 0; q < RW / 4; q++) {
            const uint4 v = __ldcs(src + q);                               // streamed: keep L2 for the memo and the tables
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
        const uint32_t rlen = h & SH_RLEN;
        uint32_t handle = H_PENDING;
        bool defer = false;
        if (h & SH_LONG) defer = true;
        else if (h & SH_N) handle = H_N;                                   // 'N' before the length (PE_Inference.py:160-163)
        else if (rlen < a.split_len) handle = H_SHORT;
        else if (h & SH_BAD) defer = true;
        else {
            walk = true;
            if (a.memo.tab != nullptr) {
                handle = memo_find<RW>(a.memo, x, rlen, memo_hash<RW>(x, rlen));
                hit = handle != H_PENDING;
                walk = !hit;
            }
        }
        if (defer) defer_read(a, atomicAdd(a.defer_count, 1ull), (uint32_t)r, slot, h);
        if (!walk) a.handles[r] = handle;                                  // (H_PENDING for a deferred read: the tiers fill it in)
    }
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, walk), mh = __ballot_sync(0xFFFFFFFFu, hit);
    if (lane == 0) { s_cnt[wid] = __popc(m); s_hit[wid] = __popc(mh); }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0, hits = 0;
        for (int w = 0; w < MM_THREADS / 32; w++) { tot += s_cnt[w]; hits += s_hit[w]; }
        s_base = tot ? atomicAdd(a.walk_count, (unsigned long long)tot) : 0ull;
        if (hits) atomicAdd(&a.counters[CNT_MEMO_HIT], (unsigned long long)hits);
    }
    __syncthreads();
    if (walk) {
        unsigned long long pos = s_base + __popc(m & ((1u << lane) - 1));
        for (uint32_t w = 0; w < wid; w++) pos += s_cnt[w];
        a.walk_list[pos] = make_uint2((uint32_t)r, (uint32_t)slot);
    }
}

// k_walk -- one thread per listed read; a fixed grid walks the device-side list.
#ifndef VSPE_WK_MINB
#define VSPE_WK_MINB 10
#endif
template <int STRIDE>
__global__ void __launch_bounds__(WK_THREADS, VSPE_WK_MINB)
k_walk(const WalkArgs a, const IndexView ix, const LinkView lv) {
    __shared__ uint32_t s_rows[WK_THREADS * STRIDE];
    __shared__ uint32_t s_lst[SM_MAXST * WK_THREADS];
    constexpr int NW = STRIDE - 3, XW = (NW + 3) / 4 * 4;
    if (*reinterpret_cast<const volatile unsigned long long*>(a.counters + CNT_ERR) & ERRF_TILE_FULL) return;
    const unsigned long long n = *a.walk_count;
    for (uint64_t i0 = (uint64_t)blockIdx.x * WK_THREADS; i0 < n; i0 += (uint64_t)gridDim.x * WK_THREADS) {
        const uint64_t i = i0 + threadIdx.x;
        bool defer = false;
        uint32_t r = 0, h = 0;
        uint64_t slot = 0;
        if (i < n) {
            const uint2 e = a.walk_list[i];
            r = e.x;
            slot = e.y;
            h = __ldg(a.hdr + slot);
            const uint32_t rlen = h & SH_RLEN;
            const uint4* src = reinterpret_cast<const uint4*>(a.rows + slot * a.row_words);
            uint32_t* row = s_rows + threadIdx.x * STRIDE;
            {
                uint32_t x[XW];
#pragma unroll
                for (int q = 0; q < XW / 4; q++) {
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if ((uint32_t)(4 * q) < a.row_words) v = __ldg(src + q);
                    x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int w = 0; w < NW; w++) row[w] = x[w];
                row[NW] = 0; row[NW + 1] = 0; row[NW + 2] = 0;
            }
            uint32_t n_kept = 0;
            bool clean = false;
            if (walk_read<STRIDE, WK_THREADS>(ix, row, rlen, s_lst + threadIdx.x, n_kept, clean)) {
                const uint32_t handle = intern_list(lv, n_kept, s_lst + threadIdx.x, WK_THREADS);
                a.handles[r] = handle;
                // only table handles outlive the call (private records are per call), only error-free reads are bounded by the graph
                if (clean && a.memo.tab != nullptr && handle < lv.T) {
                    uint32_t x[XW];                                        // (the walk may have reverse-complemented its copy)
#pragma unroll
                    for (int q = 0; q < XW / 4; q++) {
                        const uint4 v = __ldg(src + q);
                        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
                    }
                    memo_insert<XW>(a.memo, x, rlen, memo_hash<XW>(x, rlen), handle);
                }
            } else {
                defer = true;
            }
        }
        const unsigned long long d = warp_append(a.defer_count, defer);
        if (defer) {
            a.handles[r] = H_PENDING;
            defer_read(a, d, r, slot, h);
        }
    }
}

static constexpr int SM_MAXTERM = 4 * SM_MAXREC + 8;                  // terminators a tile may hold (else the chunk takes the plain path)
static constexpr int SM_WTERM = 256;                                  // ... and a warp's 4 KiB share of it
static constexpr uint32_t SM_SMEM = SM_FRONT + SM_TILE + SM_BACK + SM_MAXTERM * 2 + SM_WARPS * SM_WTERM * 2 + 64;
static_assert(SM_WARPS * SM_QCAP <= SM_MAXTERM, "the candidate queues alias the terminator table");
static_assert(SM_WARPS <= 32, "warp totals are scanned by one warp's lanes");

// One tile.  forced_phase < 0: the line phase of the tile (line number of its first line mod 4) is GUESSED from
// the bytes -- k_tile_fix checks the guess against the exact prefix sum afterwards and lists the tile for a
// second, exact launch if it was wrong; forced_phase >= 0: that phase is used as is.  parity: of the mbarrier.
template <int RW>
__device__ __forceinline__ void scan_tile(const ScanMapArgs& a, const uint32_t tile, const int forced_phase, const uint32_t parity,
                                          uint8_t* smem, unsigned long long* s_bar, uint32_t* s_wtot) {
    uint8_t* s_bytes = smem;                                           // [SM_FRONT + SM_TILE + SM_BACK]
    uint16_t* s_tp = reinterpret_cast<uint16_t*>(smem + SM_FRONT + SM_TILE + SM_BACK);   // [SM_MAXTERM] terminator positions by rank in the tile
    uint16_t* s_qid = s_tp;                                            // [SM_WARPS][SM_QCAP] candidate vectors (dead before s_tp is written)
    uint16_t* s_tpw = s_tp + SM_MAXTERM;                               // [SM_WARPS][SM_WTERM] terminator positions by rank in the warp's share
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;

    // aligned coordinates: byte `off` of the aligned stream is buffer position off - head
    const uint64_t A = ((uint64_t)a.head + a.n + 15) & ~15ull;         // aligned stream length
    const uint64_t t_lo = (uint64_t)tile * SM_TILE;
    const uint64_t ld_lo = t_lo >= SM_FRONT ? t_lo - SM_FRONT : 0;
    const uint64_t ld_hi = min(A, t_lo + SM_TILE + SM_BACK);
    const uint32_t s_off0 = tile == 0 ? SM_FRONT : 0;                  // where ld_lo lands in s_bytes
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)(ld_hi - ld_lo);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(s_bar)), "r"(bytes) : "memory");
        const uint8_t* src = a.buf - a.head + ld_lo;
        // the FASTQ bytes are read exactly once: evict-first in L2, which the k-mer table, the memo and the link tables keep
        unsigned long long policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        uint32_t done = 0;
        while (done < bytes) {
            const uint32_t part = min(bytes - done, 16384u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                             smem_u32(s_bytes + s_off0 + done)),
                         "l"(__cvta_generic_to_global(src + done)), "r"(part), "r"(smem_u32(s_bar)), "l"(policy)
                         : "memory");
            done += part;
        }
    }
    {   // wait for the bytes
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(smem_u32(s_bar)), "r"(parity) : "memory");
        }
    }
    // tile byte j (0 <= j < SM_TILE) lives at s_bytes[SM_FRONT + j]; its buffer position is t_lo + j - head
    const uint8_t* tb = s_bytes + SM_FRONT;
    const int64_t pos0 = (int64_t)t_lo - a.head;                       // buffer position of tile byte 0
    const int64_t nn = (int64_t)a.n;
    auto byte_at = [&](int64_t j) -> uint32_t {                         // tile-relative byte, 0 outside the buffer
        const int64_t p = pos0 + j;
        return (p >= 0 && p < nn) ? tb[j] : 0u;
    };

    uint16_t* q_id = s_qid + wib * SM_QCAP;
    uint16_t* tpw = s_tpw + wib * SM_WTERM;
    const bool interior = pos0 >= 1 && pos0 + SM_TILE + 16 <= nn;   // CTA-uniform
    const uint32_t lt = (1u << lane) - 1;
    uint32_t qn = 0, wcount = 0;
    bool bad = false;
    // ---- M1: which 16-byte vectors can hold a terminator?  ('\n' and '\r' are < 0x10) -----------
    // Warp w owns tile bytes [w*4K, (w+1)*4K) as 8 coalesced 512-byte rows.  (x - 0x10) | x has bit 7
    // set in every byte that is < 0x10 or >= 0x80 (a borrow can only add false positives next to a true
    // one), so two instructions per word decide.  Candidate vectors are appended, in (row, lane) order,
    // to the warp's queue: everything after this loop runs on a dense list.
#pragma unroll
    for (int it = 0; it < SM_ITERS; it++) {
        const uint32_t vid = (wib * SM_ITERS + it) * 32 + lane;       // vector index inside the tile
        const uint4 v = *reinterpret_cast<const uint4*>(tb + vid * 16);
        const uint32_t t = ((v.x - 0x10101010u) | v.x) | ((v.y - 0x10101010u) | v.y) | ((v.z - 0x10101010u) | v.z) |
                           ((v.w - 0x10101010u) | v.w);
        bool cand = (t & 0x80808080u) != 0;
        if (!interior) {                                               // first / last tile of the chunk
            const int64_t p = pos0 + (int64_t)vid * 16;
            const bool full = p >= 0 && p + 16 <= nn;
            if (!full) cand = p < nn && p + 16 > 0;                    // partial vector: M2 masks the outside bytes
        }
        const uint32_t bm = __ballot_sync(0xFFFFFFFFu, cand);
        if (cand) {
            const uint32_t at = qn + __popc(bm & lt);
            if (at < (uint32_t)SM_QCAP) q_id[at] = (uint16_t)vid;
        }
        qn += __popc(bm);
    }
    bool w_over = qn > (uint32_t)SM_QCAP;
    if (w_over) qn = SM_QCAP;
    __syncwarp();
    // ---- M2: exact terminator masks of the candidates (universal newlines); every terminator's position goes to
    // the warp's table at its rank (warp scan of the per-vector counts) ---------------------------------------
    for (uint32_t i0 = 0; i0 < qn; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint32_t term = 0, j = 0;
        if (i < qn) {
            j = (uint32_t)q_id[i] * 16;
            const uint4 v = *reinterpret_cast<const uint4*>(tb + j);
            uint32_t valid = 0xFFFFu;
            if (!interior) {
                const int64_t p = pos0 + j;
                if (p < 0) valid &= 0xFFFFu << (uint32_t)(-p);
                if (p + 16 > nn) valid &= 0xFFFFu >> (uint32_t)(p + 16 - nn);
            }
            masks_from_vec(v, valid, interior ? (uint32_t)tb[j + 16] : byte_at((int64_t)j + 16), bad, term);
        }
        const uint32_t c = __popc(term);
        uint32_t inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        uint32_t rk = wcount + inc - c;
        while (term) {
            const uint32_t k = (uint32_t)__ffs((int)term) - 1;
            term &= term - 1;
            if (rk < (uint32_t)SM_WTERM) tpw[rk] = (uint16_t)(j + k);
            rk++;
        }
        wcount += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    if (bad) atomicOr(&a.counters[CNT_ERR], (unsigned long long)ERRF_NON_ASCII);
    w_over |= wcount > (uint32_t)SM_WTERM;
    if (lane == 0) s_wtot[wib] = wcount | (w_over ? 0x80000000u : 0u);
    __syncthreads();
    // warp totals -> this warp's first rank and the tile's terminator count (one scan over SM_WARPS lanes)
    uint32_t tile_total, warp_base;
    bool too_many;
    {
        const uint32_t xw = lane < (uint32_t)SM_WARPS ? s_wtot[lane] : 0u;
        const uint32_t xc = xw & 0x7FFFFFFFu;
        uint32_t inc = xc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        warp_base = __shfl_sync(0xFFFFFFFFu, inc - xc, wib);
        tile_total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        too_many = __any_sync(0xFFFFFFFFu, (xw >> 31) != 0) || tile_total > (uint32_t)SM_MAXTERM;
    }
    if (!too_many)
        for (uint32_t k = lane; k < wcount; k += 32) s_tp[warp_base + k] = tpw[k];
    __syncthreads();
    // ---- line phase: terminator i of the tile ends line number base + i of the chunk's file; i ends a header line iff
    // (base + i) % 4 == 0.  b = base % 4 is not known here (it needs every earlier tile): it is guessed from the first
    // byte of the 32 lines that follow the tile's first terminators -- a header line starts with '@', a separator line
    // with '+' -- as the smallest b without a contradiction.  Any guess is safe: k_tile_fix verifies it.
    uint32_t b = 0;
    if (forced_phase >= 0) b = (uint32_t)forced_phase;
    else if (!too_many) {
        uint32_t viol = 0;
        if (lane < tile_total) {
            const uint32_t j = (uint32_t)s_tp[lane] + 1;
            if (pos0 + (int64_t)j < nn) {
                const uint32_t ch = tb[j];
                if (ch != '@') viol |= 1u << ((3u - lane) & 3u);       // the phase under which this line is a header
                if (ch != '+') viol |= 1u << ((1u - lane) & 3u);       // ... a separator
            }
        }
        viol = __reduce_or_sync(0xFFFFFFFFu, viol);
        b = viol == 0xFu ? 0u : (uint32_t)__ffs((int)(~viol & 0xFu)) - 1;
    }
    // reads owned by this tile: sequence lines that START here, i.e. follow a header terminator i = h0, h0 + 4, ...
    const uint32_t h0 = (4u - b) & 3u;
    const uint32_t n_own = tile_total > h0 ? (tile_total - h0 + 3) >> 2 : 0u;
    const bool chunk_starts_in_seq = tile == 0 && (a.line_base & 3) == 1;   // chunk begins with a sequence line
    const uint32_t shift = chunk_starts_in_seq ? 1u : 0u;              // that read becomes local index 0
    const uint32_t n_local = n_own + shift;
    too_many |= n_local > a.tcap;
    if (threadIdx.x == 0) a.tile_info[tile] = (unsigned long long)tile_total | ((unsigned long long)b << 32) | (too_many ? 1ull << 40 : 0ull);
    if (too_many) {
        if (threadIdx.x == 0) atomicOr(&a.counters[CNT_ERR], (unsigned long long)ERRF_TILE_FULL);
        return;
    }
    // ---- pack: the row goes from registers to HBM.  Rows of 16 / 20 words (reads of up to 256 / 320 bases) are packed by
    // TWO threads per read, each half of the words: a tile owns ~80 such reads, so this keeps 5 warps busy instead of 3 and
    // halves the phase's critical path (measured on 2x250: k_scan_rows 0.274 -> 0.237 ms per 516 MB).  Rows of 12 words
    // (up to 160 bases; a tile owns ~130 reads) stay with one thread per read: the halves would be 6 and 4 steps and the
    // split was measured 5 % slower there.  16 bases = four words per step: code = (ascii >> 1) & 3; the only byte with
    // code k is "ACTG"[k] = 0x41 + 2k (+15 when k == 2), which is the validity test.
    constexpr int NW = RW == 12 ? 10 : RW;                             // row words that can hold bases (cap / 16)
    constexpr int TPR = RW >= 16 ? 2 : 1;                              // threads per read
    constexpr int HW = RW / TPR;                                       // row words per thread
    const uint32_t half = TPR == 2 ? (threadIdx.x & 1) : 0u;
    for (uint32_t base_li = 0; base_li < n_local; base_li += SM_THREADS / TPR) {   // block-uniform trips: the exchange below is a full-warp shuffle
        const uint32_t li = base_li + (TPR == 2 ? (threadIdx.x >> 1) : threadIdx.x);
        const bool act = li < n_local;
        uint32_t st = 0, en = 0xFFFFFFFFu, h = 0, rlen = 0, diff = 0, sh = 0;
        const uint32_t* p = reinterpret_cast<const uint32_t*>(s_bytes);
        uint32_t rw[HW];
#pragma unroll
        for (int k = 0; k < HW; k++) rw[k] = 0;
        if (act) {
            if (li < shift) {
                st = a.head;                                            // buffer position 0, tile-relative
                if (tile_total > 0) en = s_tp[0];
            } else {
                const uint32_t i = h0 + 4 * (li - shift);
                st = (uint32_t)s_tp[i] + 1;
                if (i + 1 < tile_total) en = s_tp[i + 1];
            }
            if (en != 0xFFFFFFFFu) {
                if (en > st && tb[en] == '\n' && tb[en - 1] == '\r') en--;  // "\r\n": the '\r' is not part of the line
            } else {
                // the line ends beyond the tile: first '\n' or '\r' in the back margin, if any (four bytes per step:
                // only a word with a byte < 0x10 is looked at byte by byte)
                for (uint32_t j = SM_TILE; j < (uint32_t)(SM_TILE + SM_BACK) && en == 0xFFFFFFFFu; j += 4) {
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(tb + j);
                    if (!(((w - 0x10101010u) | w) & 0x80808080u) && pos0 + (int64_t)j + 4 <= nn) continue;
                    for (uint32_t k = 0; k < 4 && pos0 + (int64_t)(j + k) < nn; k++) {
                        const uint32_t ch = (w >> (8 * k)) & 0xFF;
                        if (ch == '\n' || ch == '\r') { en = j + k; break; }
                    }
                    if (pos0 + (int64_t)j + 4 > nn) break;
                }
                if (en == 0xFFFFFFFFu) h |= SH_LONG;
            }
            rlen = (h & SH_LONG) ? 0u : en - st;
            if (rlen > a.cap) { h |= SH_LONG; rlen = 0; }
            const uint32_t a0 = SM_FRONT + st;                          // offset of the first base in s_bytes
            p = reinterpret_cast<const uint32_t*>(s_bytes + (a0 & ~3u));
            sh = (a0 & 3) * 8;
            const uint32_t v0 = half * HW;                              // this thread's first row word
            if (16u * v0 < rlen) {
                uint32_t carry = p[4 * v0];
#pragma unroll
                for (int k = 0; k < HW; k++) {
                    const uint32_t v = v0 + k;
                    uint32_t word = 0;
                    if (v < (uint32_t)NW && 16u * v < rlen) {
                        uint32_t x[5];
                        x[0] = carry;
#pragma unroll
                        for (int q = 1; q < 5; q++) x[q] = p[4 * v + q];
                        carry = x[4];
                        const uint32_t rem = rlen - 16u * v;            // bases left, >= 1
                        if (rem >= 16) {
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const uint32_t c = __funnelshift_r(x[q], x[q + 1], sh);
                                const uint32_t c2 = (c >> 1) & 0x03030303u;
                                const uint32_t expect = 0x41414141u + 2 * c2 + 15 * ((c2 >> 1) & ~c2 & 0x01010101u);
                                diff |= expect ^ c;
                                word |= ((c2 * 0x01041040u) >> 24) << (8 * q);
                            }
                        } else {                                        // the read ends inside this step: mask the bytes beyond it
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const uint32_t c = __funnelshift_r(x[q], x[q + 1], sh);
                                const uint32_t vm = rem >= 4u * q + 4 ? 0xFFFFFFFFu : rem <= 4u * q ? 0u : (0xFFFFFFFFu >> (8 * (4u * q + 4 - rem)));
                                const uint32_t c2 = ((c & vm) >> 1) & 0x03030303u;
                                const uint32_t expect = 0x41414141u + 2 * c2 + 15 * ((c2 >> 1) & ~c2 & 0x01010101u);
                                diff |= (expect ^ c) & vm;
                                word |= ((c2 * 0x01041040u) >> 24) << (8 * q);
                            }
                        }
                    }
                    rw[k] = word;
                }
            }
        }
        if (TPR == 2) diff |= __shfl_xor_sync(0xFFFFFFFFu, diff, 1);    // either half saw a byte that is not ACGT
        if (!act) continue;
        const uint64_t slot = (uint64_t)tile * a.tcap + li;
        if (TPR == 2) {
            uint2* dst = reinterpret_cast<uint2*>(a.rows + slot * RW + half * HW);
#pragma unroll
            for (int k = 0; k + 1 < HW; k += 2) __stcs(dst + (k >> 1), make_uint2(rw[k], rw[k + 1]));
        } else {
            uint4* dst = reinterpret_cast<uint4*>(a.rows + slot * RW);
#pragma unroll
            for (int k = 0; k + 3 < HW; k += 4) __stcs(dst + (k >> 2), make_uint4(rw[k], rw[k + 1], rw[k + 2], rw[k + 3]));
        }
        if (half != 0) continue;
        if (diff) {
            // rare: some byte is not ACGT -- 'N' (pair skipped, PE_Inference.py:160) or anything else (exhaustive tier), word by word
            bool hasN = false, badc = false;
            uint32_t cy = p[0];
            for (uint32_t v4 = 0; 4 * v4 < rlen; v4++) {
                const uint32_t nx = p[v4 + 1];
                const uint32_t c = __funnelshift_r(cy, nx, sh);
                cy = nx;
                const uint32_t left = rlen - 4 * v4;
                const uint32_t vm = left >= 4 ? 0xFFFFFFFFu : (0xFFFFFFFFu >> (8 * (4 - left)));
                const uint32_t c2 = ((c & vm) >> 1) & 0x03030303u;
                const uint32_t expect = 0x41414141u + 2 * c2 + 15 * ((c2 >> 1) & ~c2 & 0x01010101u);
                uint32_t badm = (expect ^ c) & vm;                      // non-zero bytes = invalid characters
                if (badm) {
                    for (int k = 0; k < 4; k++)
                        if ((badm >> (8 * k)) & 0xFF) { if (((c >> (8 * k)) & 0xFF) == 'N') hasN = true; else badc = true; }
                }
            }
            h |= (hasN ? SH_N : 0) | (badc ? SH_BAD : 0);
        }
        __stcs(a.hdr + slot, h | rlen | (st << 16));
    }
}

// k_scan_rows: one tile per CTA, no tile waits on another.  k_scan_redo: the tiles k_tile_fix listed, with their exact phase
// (a fixed grid walks the device-side list; normally it is empty).
template <int RW>                                                // row words in HBM: 12, 16 or 20 (a multiple of 4 >= cap / 16)
__global__ void __launch_bounds__(SM_THREADS, VSPE_SM_MINB)
k_scan_rows(const ScanMapArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_wtot[SM_WARPS];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t tile = blockIdx.x;
    scan_tile<RW>(a, tile, tile == 0 ? (int)(a.line_base & 3) : -1, 0u, smem, &s_bar, s_wtot);
}

template <int RW>
__global__ void __launch_bounds__(SM_THREADS, VSPE_SM_MINB)
k_scan_redo(const ScanMapArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_wtot[SM_WARPS];
    const unsigned long long n = *a.n_redo;
    if (blockIdx.x >= n) return;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    for (unsigned long long i = blockIdx.x; i < n; i += gridDim.x, parity ^= 1u) {
        const unsigned long long e = a.redo[i];
        scan_tile<RW>(a, (uint32_t)e, (int)((e >> 32) & 3), parity, smem, &s_bar, s_wtot);
        __syncthreads();                                               // the tile buffer and the tables are reused by the next round
    }
}

// Exact line numbers for every tile from the tiles' terminator counts: k_tile_sum adds up 1024 tiles per CTA,
// k_tile_fix (same grid) turns the sums before its CTA into its base, scans its 1024 tiles and writes
//   r_first[t]   chunk-local index of the first read tile t owns (r_first[n_tiles]: one past the last read),
//   blk_tile[b]  the tile that holds read 128 b (where k_walk block b starts looking),
//   redo list    tiles whose guessed phase differs from the exact one,
//   total        terminators of the chunk.
struct TileFixArgs {
    const unsigned long long* tile_info;
    uint32_t n_tiles;
    uint64_t line_base, rec_first;
    unsigned long long* part;          // [ceil(n_tiles / 1024)] terminators per CTA
    uint32_t* r_first;
    uint32_t* blk_tile;
    uint32_t n_blk;
    unsigned long long* redo;
    unsigned long long* n_redo;        // zeroed before the launch
    unsigned long long* total;
};

__device__ __forceinline__ unsigned long long block_sum_1024(unsigned long long x, unsigned long long* s_warp) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, d);
    __syncthreads();
    if (lane == 0) s_warp[wid] = x;
    __syncthreads();
    unsigned long long t = 0;
    for (uint32_t w = 0; w < 32; w++) t += s_warp[w];
    return t;
}

__global__ void __launch_bounds__(1024)
k_tile_sum(const TileFixArgs a) {
    __shared__ unsigned long long s_warp[32];
    const uint32_t t = blockIdx.x * 1024 + threadIdx.x;
    const unsigned long long x = t < a.n_tiles ? (unsigned long long)(uint32_t)a.tile_info[t] : 0ull;
    const unsigned long long tot = block_sum_1024(x, s_warp);
    if (threadIdx.x == 0) a.part[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024)
k_tile_fix(const TileFixArgs a) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_scan[32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // terminators before this CTA's tiles (and, in the last CTA, the chunk's total)
    unsigned long long before = 0;
    for (uint32_t i = threadIdx.x; i < blockIdx.x; i += 1024) before += a.part[i];
    before = block_sum_1024(before, s_warp);
    const uint32_t t = blockIdx.x * 1024 + threadIdx.x;
    const unsigned long long info = t < a.n_tiles ? a.tile_info[t] : 0ull;
    const uint32_t tot = (uint32_t)info, ph = (uint32_t)(info >> 32) & 3u;
    unsigned long long inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += y;
    }
    if (lane == 31) s_scan[wid] = inc;
    __syncthreads();
    unsigned long long wbase = 0;
    for (uint32_t w = 0; w < wid; w++) wbase += s_scan[w];
    if (t >= a.n_tiles) return;
    const unsigned long long run = before + wbase + inc - tot;         // terminators before tile t
    const uint64_t base = a.line_base + run;                           // line number of the tile's first line
    const uint32_t shift = (t == 0 && (a.line_base & 3) == 1) ? 1u : 0u;
    const uint32_t rf = (uint32_t)(((base + 3) >> 2) - shift - a.rec_first);
    const uint32_t rn = (uint32_t)(((base + tot + 3) >> 2) - a.rec_first);       // = r_first[t + 1]
    a.r_first[t] = rf;
    if (t != 0 && ph != ((uint32_t)base & 3u)) a.redo[atomicAdd(a.n_redo, 1ull)] = (unsigned long long)t | ((unsigned long long)(base & 3) << 32);
    for (uint32_t bk = (rf + WK_THREADS - 1) / WK_THREADS; bk < a.n_blk && bk * WK_THREADS < rn; bk++) a.blk_tile[bk] = t;
    if (t + 1 == a.n_tiles) {
        a.r_first[t + 1] = rn;
        *a.total = run + tot;
    }
}

// The default path over a device-resident chunk: k_scan_rows -> k_tile_fix -> k_scan_redo -> k_walk over up to n_slots
// reads (every size the later kernels need is read on the device, so there is no host synchronisation in between).
int scan_map(Ctx* c, int m, const uint8_t* d_buf, uint64_t n, uint64_t line_base, uint64_t rec_first, uint64_t n_slots, uint32_t* d_handles,
             unsigned long long* d_defer_count, uint32_t row_words, uint32_t cap) {
    if (n == 0) return VSPE_OK;
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(d_buf) & 15);
    const uint64_t n_tiles = (n + head + SM_TILE - 1) / SM_TILE;
    if (n_tiles > 0x7FFFFFFFull) { set_error("buffer too large for one scan launch"); return VSPE_ERR_LIMIT; }
    if (n_slots > 0xFFFFFFF0ull) { set_error("more than 2^32 reads in one chunk"); return VSPE_ERR_LIMIT; }
    MateBuf& mb = c->mate[m];
    // slots per tile: four times what a tile holds when every read has the hinted length (shorter reads, e.g. trimmed
    // ones, make more records per tile; the slots are address space, only the used ones are ever touched); a tile that
    // still overflows sends the chunk to the plain path
    uint32_t tcap = 4 * (SM_TILE / (2 * std::max<uint32_t>(c->read_len_hint, 1) + 6) + 2);
    tcap = std::min<uint32_t>((tcap + 7) & ~7u, SM_MAXREC);
    if (n_tiles * tcap > 0xFFFFFFFFull) { set_error("buffer too large for one scan launch (tile slots exceed 2^32)"); return VSPE_ERR_LIMIT; }
    const uint64_t n_blk = (n_slots + WK_THREADS - 1) / WK_THREADS;
    VSPE_TRY(mb.rec.rows.reserve(n_tiles * tcap * row_words));
    VSPE_TRY(mb.rec.hdr.reserve(n_tiles * tcap));
    const uint64_t n_part = (n_tiles + 1023) / 1024;
    VSPE_TRY(c->tile_base_m[m].reserve(2 * n_tiles + 4 + n_part));      // tile_info, redo list, redo count, total, per-CTA sums
    VSPE_TRY(c->tile_idx_m[m].reserve(n_tiles + 1 + n_blk));            // r_first, blk_tile
    VSPE_TRY(mb.rec.seq_start.reserve(n_slots + 2));                    // compact arrays of the deferred reads
    VSPE_TRY(mb.rec.seq_end.reserve(n_slots + 2));
    VSPE_TRY(mb.d_hdr.reserve(n_slots + 2));
    VSPE_TRY(mb.d_rows.reserve((n_slots + 2) * row_words));
    VSPE_TRY(c->defer_m[m].reserve(n_slots + 2));
    VSPE_TRY(c->walk_list_m[m].reserve(n_slots + 2));
    unsigned long long* info = reinterpret_cast<unsigned long long*>(c->tile_base_m[m].p);
    ScanMapArgs a;
    a.buf = d_buf; a.n = n; a.head = head; a.n_tiles = (uint32_t)n_tiles; a.line_base = line_base;
    a.tile_info = info; a.redo = info + n_tiles; a.n_redo = info + 2 * n_tiles;
    a.rows = mb.rec.rows.p; a.hdr = mb.rec.hdr.p; a.tcap = tcap; a.cap = cap; a.counters = c->counters.p;
    TileFixArgs f;
    f.tile_info = info; f.n_tiles = (uint32_t)n_tiles; f.line_base = line_base; f.rec_first = rec_first;
    f.r_first = c->tile_idx_m[m].p; f.blk_tile = f.r_first + n_tiles + 1; f.n_blk = (uint32_t)n_blk;
    f.redo = info + n_tiles; f.n_redo = info + 2 * n_tiles; f.total = info + 2 * n_tiles + 1; f.part = info + 2 * n_tiles + 4;
    if (!c->scan_map_attr_set) {
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_rows<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM));
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_rows<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM));
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_rows<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM));
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_redo<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM));
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_redo<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM));
        VSPE_CUDA(cudaFuncSetAttribute(k_scan_redo<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM));
        for (auto& evs : c->ev_scan) for (auto& e : evs) if (!e) VSPE_CUDA(cudaEventCreate(&e));
        c->scan_map_attr_set = true;
    }
    VSPE_CUDA(cudaMemsetAsync(info + 2 * n_tiles, 0, 16, c->stream));
    // both big kernels are timed on their own (CUDA events on the launching stream; two launches may be in flight)
    const int k = c->scan_map_events & 1;
    VSPE_CUDA(cudaEventRecord(c->ev_scan[0][2 * k], c->stream));
    if (row_words == 12) k_scan_rows<12><<<(uint32_t)n_tiles, SM_THREADS, SM_SMEM, c->stream>>>(a);
    else if (row_words == 16) k_scan_rows<16><<<(uint32_t)n_tiles, SM_THREADS, SM_SMEM, c->stream>>>(a);
    else k_scan_rows<20><<<(uint32_t)n_tiles, SM_THREADS, SM_SMEM, c->stream>>>(a);
    VSPE_LAUNCH_CHECK(c);
    VSPE_CUDA(cudaEventRecord(c->ev_scan[0][2 * k + 1], c->stream));
    k_tile_sum<<<(uint32_t)n_part, 1024, 0, c->stream>>>(f);
    VSPE_LAUNCH_CHECK(c);
    k_tile_fix<<<(uint32_t)n_part, 1024, 0, c->stream>>>(f);
    VSPE_LAUNCH_CHECK(c);
    const uint32_t redo_grid = (uint32_t)std::min<uint64_t>(n_tiles, (uint64_t)c->sm_count * VSPE_SM_MINB);
    if (row_words == 12) k_scan_redo<12><<<redo_grid, SM_THREADS, SM_SMEM, c->stream>>>(a);
    else if (row_words == 16) k_scan_redo<16><<<redo_grid, SM_THREADS, SM_SMEM, c->stream>>>(a);
    else k_scan_redo<20><<<redo_grid, SM_THREADS, SM_SMEM, c->stream>>>(a);
    VSPE_LAUNCH_CHECK(c);
    // the memo lives as long as the list handles do (link_reset clears it); its entry size follows the row size
    if (c->opt_memo && !c->memo_off) {
        const uint32_t stride = 4 + row_words;
        if (!c->memo.p || c->memo_stride != stride) {
            c->memo.release();
            VSPE_TRY(c->memo.reserve((size_t)MEMO_ENTRIES * stride));
            c->memo_stride = stride;
            VSPE_CUDA(cudaMemsetAsync(c->memo.p, 0, (size_t)MEMO_ENTRIES * stride * 4, c->stream));
        }
    }
    WalkArgs w;
    w.rows = mb.rec.rows.p; w.hdr = mb.rec.hdr.p; w.row_words = row_words; w.tcap = tcap; w.r_first = f.r_first; w.blk_tile = f.blk_tile;
    w.head = head; w.n_slots = n_slots; w.total = f.total; w.line_base = line_base; w.rec_first = rec_first; w.split_len = c->index.split_len;
    w.handles = d_handles; w.walk_list = c->walk_list_m[m].p; w.walk_count = c->counters.p + (m == 0 ? CNT_WALK : CNT_WALK1);
    w.d_read = c->defer_m[m].p; w.d_hdr = mb.d_hdr.p; w.d_rows = mb.d_rows.p; w.d_start = mb.rec.seq_start.p; w.d_end = mb.rec.seq_end.p;
    w.defer_count = d_defer_count; w.counters = c->counters.p;
    w.memo.tab = (c->opt_memo && !c->memo_off) ? c->memo.p : nullptr; w.memo.mask = MEMO_ENTRIES - 1; w.memo.stride = c->memo_stride;
    const IndexView ix = c->index.view();
    const LinkView lv = link_view(c);
    VSPE_CUDA(cudaEventRecord(c->ev_scan[1][2 * k], c->stream));
    // While the memo is cold: rounds of k_memo -> k_walk over growing ranges of reads (256 Ki, 1 Mi, 4 Mi, the rest) --
    // what a round's walks enter into the memo answers the later rounds, so one large chunk warms its own memo.
    // Afterwards: one round.
    const bool memo_cold = w.memo.tab != nullptr && c->memo_seen < (2ull << 20);
    c->memo_seen += n_slots;
    for (uint64_t lo = 0, len = memo_cold ? 1ull << 18 : n_slots; lo < n_slots; lo += len, len = len < (4ull << 20) ? len * 4 : n_slots) {
        w.r_lo = lo;
        w.r_hi = std::min<uint64_t>(n_slots, lo + len);
        const uint64_t n_round = w.r_hi - w.r_lo;
        VSPE_CUDA(cudaMemsetAsync(w.walk_count, 0, 8, c->stream));
        const uint32_t mgrid = (uint32_t)((n_round + MM_THREADS - 1) / MM_THREADS);
        if (row_words == 12) k_memo<12><<<mgrid, MM_THREADS, 0, c->stream>>>(w);
        else if (row_words == 16) k_memo<16><<<mgrid, MM_THREADS, 0, c->stream>>>(w);
        else k_memo<20><<<mgrid, MM_THREADS, 0, c->stream>>>(w);
        VSPE_LAUNCH_CHECK(c);
        const uint32_t grid = (uint32_t)std::min<uint64_t>((n_round + WK_THREADS - 1) / WK_THREADS, (uint64_t)c->sm_count * VSPE_WK_MINB);
        if (cap <= 160) k_walk<13><<<grid, WK_THREADS, 0, c->stream>>>(w, ix, lv);
        else if (cap <= 256) k_walk<19><<<grid, WK_THREADS, 0, c->stream>>>(w, ix, lv);
        else k_walk<23><<<grid, WK_THREADS, 0, c->stream>>>(w, ix, lv);
        VSPE_LAUNCH_CHECK(c);
    }
    VSPE_CUDA(cudaEventRecord(c->ev_scan[1][2 * k + 1], c->stream));
    c->scan_map_pending[k] = true;
    c->scan_map_events++;
    return VSPE_OK;
}

// {tiles listed for k_scan_redo, terminators of the chunk} of the last scan_map launch (two device words, read after a stream sync)
const unsigned long long* scan_map_total_ptr(Ctx* c, int m, uint64_t n, const uint8_t* d_buf) {
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(d_buf) & 15);
    const uint64_t n_tiles = (n + head + SM_TILE - 1) / SM_TILE;
    return reinterpret_cast<unsigned long long*>(c->tile_base_m[m].p) + 2 * n_tiles;
}

// fold the durations of the finished k_scan_rows / k_walk launches into the stats (call after a stream sync)
void scan_map_account(Ctx* c) {
    for (int k = 0; k < 2; k++) {
        if (!c->scan_map_pending[k]) continue;
        float ms = 0;
        if (cudaEventQuery(c->ev_scan[1][2 * k + 1]) != cudaSuccess) continue;
        if (cudaEventElapsedTime(&ms, c->ev_scan[0][2 * k], c->ev_scan[0][2 * k + 1]) == cudaSuccess) { c->stats.ms_k_scan_rows += ms; c->stats.n_k_scan_rows++; }
        if (cudaEventElapsedTime(&ms, c->ev_scan[1][2 * k], c->ev_scan[1][2 * k + 1]) == cudaSuccess) { c->stats.ms_k_walk += ms; c->stats.n_k_walk++; }
        c->scan_map_pending[k] = false;
    }
    cudaGetLastError();
}

}  // namespace vspe
