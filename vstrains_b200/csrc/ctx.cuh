// ctx.cuh -- the per-device context behind the opaque vspe_ctx handle.
#pragma once
#include "vspe_internal.cuh"

namespace vspe {

// device-side counters (one u64 each), zeroed by vspe_reset
enum Counter {
    CNT_TOTAL = 0, CNT_N, CNT_SHORT, CNT_USED, CNT_KEYS,
    CNT_SPILL_CURSOR,        // next free entry of the spill pool
    CNT_ERR,                 // error bit flags raised by kernels
    CNT_FAST, CNT_GENERIC,   // reads resolved per tier
    CNT_WORK,                // generic-tier worklist length
    CNT_BAILED,              // reads the fast tier handed to the exhaustive tier
    CNT_DEFER,               // reads k_walk left for the list-driven tiers
    CNT_WORK2,               // reads k_map_windows left for the ASCII tier
    CNT_DEFER1,              // ... the same for the second mate file (the two mates' tiers may be in flight together)
    CNT_BIG,                 // (unused)
    CNT_LISTS,               // distinct node lists interned in the list table
    CNT_OVF,                 // private list records of this call (lists the table cannot hold)
    CNT_PAIR_OCC,            // distinct (left list, right list) combinations of the current batch
    CNT_EXP,                 // weighted keys the current batch expands to
    CNT_EXP_CURSOR,          // ... and the emit cursor over them
    CNT_B_USED, CNT_B_N, CNT_B_SHORT,   // pair classes of the current batch (added to USED / N / SHORT once the batch is accepted)
    CNT_WALK, CNT_WALK1,     // per mate: reads k_memo listed for k_walk
    CNT_MEMO_HIT,            // reads whose handle came from the read memo
    CNT_PAIR_LIST,           // occupied pair-table entries listed by k_comb_weigh for k_comb_emit
    CNT_COUNT_
};
static constexpr uint64_t ERRF_NON_ASCII = 1, ERRF_SPILL_FULL = 2, ERRF_KEYS_FULL = 4, ERRF_SLOTS_FULL = 8, ERRF_TILE_FULL = 16, ERRF_LISTS_FULL = 32, ERRF_INTERNAL = 64, ERRF_PAIRS_FULL = 128;
// flags that end the run (the *_FULL scan flags are transient: the host repeats or ignores the launch)
static constexpr uint64_t ERRF_FATAL = ERRF_NON_ASCII | ERRF_SPILL_FULL | ERRF_KEYS_FULL | ERRF_LISTS_FULL | ERRF_INTERNAL;

// sparse (COO) counting state: sorted runs (key, count) at the front of k[0] / v[0]
struct Sparse {
    bool enabled = false;
    uint64_t n_runs = 0;
    DevBuf<unsigned long long> k[2], v[2], vscan, sums64, totals, m, moff;
    DevBuf<uint32_t> hist, sums32, head;
    std::vector<uint64_t> h_keys, h_counts;      // host copies handed out by vspe_sparse_host
};

struct MateBuf {
    Records rec;
    DevBuf<ReadSlot> slots;
    DevBuf<uint32_t> d_hdr, d_rows;   // compact copies (header, packed row) of the reads k_walk deferred
    DevBuf<uint32_t> handles;  // [n_recs] list handle (or H_N / H_SHORT) per read
    uint64_t n_recs = 0;      // complete records = lines / 4
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream[2] = {nullptr, nullptr};
    cudaStream_t tier_stream = nullptr;  // list-driven tiers of one mate run here while the other mate is scanned on `stream`
    cudaEvent_t ev_walk[2] = {}, ev_tier[2] = {};
    cudaEvent_t ev[8] = {};
    cudaEvent_t ev_scan[2][4] = {};     // events around the k_scan_rows [0] / k_walk [1] launches (two in flight each)
    Index index;
    DevBuf<uint64_t> mats;            // [2][N][N] node_mat then short_mat (dense mode)
    Sparse sparse;                    // sorted (key, count) runs (sparse mode)
    int64_t opt_sparse = 0;           // force sparse counting
    DevBuf<unsigned long long> counters;
    // K1 scratch
    DevBuf<uint32_t> tile_counts;
    DevBuf<uint64_t> tile_base;
    MateBuf mate[2];
    // K4 generic-tier scratch
    DevBuf<uint32_t> warp_scratch;
    bool scratch_valid = false;       // warp_scratch initialised for the current index
    DevBuf<uint32_t> spill;
    DevBuf<uint32_t> worklist;
    DevBuf<uint32_t> defer_list;       // tier scratch: reads k_map_windows leaves for the ASCII tier
    DevBuf<uint32_t> defer_m[2];       // per mate: reads k_walk left unresolved
    DevBuf<uint64_t> tile_base_m[2];   // per mate: k_scan_rows' tile records, k_tile_fix's redo list + count, terminator total
    DevBuf<uint2> walk_list_m[2];      // per mate: {read, slot} of the reads k_memo left for k_walk
    DevBuf<uint32_t> memo;             // read memo: packed row -> list handle (scan_map.cu)
    uint32_t memo_stride = 0;          // words per memo entry (4 + row words)
    uint64_t memo_seen = 0;            // reads that went through k_memo since the memo was cleared (estimate: cold / warm)
    bool memo_off = false;             // switched off for this index: the input does not repeat reads
    DevBuf<uint32_t> tile_idx_m[2];    // per mate: first read of every tile, first tile of every k_walk block (k_tile_fix)
    // K5/K6: list table + pair table (link.cuh)
    DevBuf<ListRec> list_recs;         // [list_T] table part + [list_ov_cap] private records
    DevBuf<uint32_t> list_occ;         // [list_T]
    uint32_t list_T = 0, list_ov_cap = 0;
    DevBuf<PairEnt> pair_tab;          // [pair_cap] (power of two), all zero between batches
    DevBuf<uint32_t> pair_occ;         // occupied entries of the pair table (k_comb_weigh -> k_comb_emit)
    uint64_t pair_cap = 0;
    DevBuf<uint32_t> wk_hist, wk_keys;   // dense accumulation scratch: bucket histogram / segments, partitioned (digit, weight)
    bool link_attr_set = false;
    int mf_blocks_per_sm = 0;          // resident k_map_fast blocks per SM for the row capacity mf_blocks_cap (occupancy query)
    uint32_t mf_blocks_cap = 0;
    unsigned long long last_err_flags = 0;
    bool err_flags_fresh = false;
    cudaEvent_t ev_m[2][3] = {};       // per mate: scan start, scan end / map start, map end
    bool scan_map_attr_set = false;     // fused scan + map kernel (scan_map.cu): attributes set, launch events pending
    bool scan_map_pending[2] = {false, false};
    uint32_t scan_map_events = 0;
    // staging for host-input entry points
    uint8_t* pinned[2] = {nullptr, nullptr};
    size_t pinned_bytes = 0;
    DevBuf<uint8_t> dev_in[2];
    // host copies for vspe_map_reads / vspe_split_records
    std::vector<uint64_t> h_offsets, h_seq_start;
    std::vector<uint32_t> h_nodes, h_seq_len;
    std::vector<uint8_t> h_status;
    // options
    int64_t opt_force_generic = 0;
    int64_t opt_chunk_mb = 256;
    int64_t opt_link_split = 65536;    // k_bucket_count: keys of a bucket per CTA before the bucket is shared
    int64_t opt_memo = 1;              // ask / fill the read memo (scan_map.cu)
    int64_t opt_tier_overlap = 1;      // device-resident calls: run one mate's list-driven tiers beside the other mate's scan
    int64_t opt_pair_cap_log2 = 21;    // first size of the pair table (log2 entries); 0: size it by the pairs of the batch
    int64_t opt_stage_threads = 8;     // host threads that copy an unpinned input chunk into the pinned staging buffer
    int64_t opt_subst = 1;             // build / use the substitution-hit bitmap
    int64_t opt_scan_two_pass = 0;     // K1 as count + index passes (cross-check of the look-back kernel)
    uint64_t cur_buf_n = 0;            // bytes of the chunk being mapped (exhaustive tier bound)
    int64_t opt_scan_mode = 0;         // 0: k_scan_rows + k_walk, 1: look-back record scan + raw-byte map, 2: two-pass record scan
    uint32_t read_len_hint = 320;      // longest sequence line among the first records of the input
    // accounting
    vspe_stats stats = {};
    bool stats_overridden = false;
    uint64_t launches = 0;
    int sm_count = 148;
};

#define VSPE_LAUNCH_CHECK(c)                                                             \
    do {                                                                                 \
        (c)->launches++;                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            vspe::set_error("kernel launch failed: %s at %s:%d", cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                         \
            return VSPE_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

}  // namespace vspe

// the opaque handle of include/vspe.h is the context itself
struct vspe_ctx : public vspe::Ctx {};
