/*
 * vspe.h -- C ABI of libvspe.so: B200-native paired-end link inference for VStrains.
 *
 * The reference has no FFI: its boundary is a CLI subprocess
 * (reference utils/VStrains_SPAdes.py:118-132 shells out to
 *  utils/VStrains_PE_Inference.py, then reads <out>/aln/pe_info and st_info at :134-138).
 * This header is the surface a maintainer binds instead (ctypes stub in INTEGRATION.md);
 * each entry point names the reference lines it replaces.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a
 * negative vspe_status; vspe_last_error() gives a thread-local message.  There is no CPU
 * fallback: every compute entry point fails with VSPE_ERR_CUDA if no sm_100 device works.
 */
#ifndef VSPE_H
#define VSPE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vspe_status {
    VSPE_OK = 0,
    VSPE_ERR_ARG = -1,        /* bad argument / call order                                   */
    VSPE_ERR_NODE_SEQ = -2,   /* node sequence >= split_len has a non-ACGT char: the reference
                                 raises KeyError in reverse_seq (PE_Inference.py:9-13,122)   */
    VSPE_ERR_NON_ASCII = -3,  /* byte >= 0x80 in an input (outside the reference's contract) */
    VSPE_ERR_GFA = -4,        /* S line with fewer than 3 fields (IndexError, :108-111)      */
    VSPE_ERR_IO = -5,         /* unreadable input / unwritable output                        */
    VSPE_ERR_CUDA = -6,       /* CUDA runtime failure or no usable device                    */
    VSPE_ERR_LIMIT = -7,      /* documented size limit exceeded                              */
    VSPE_ERR_NCCL = -8
} vspe_status;

/* Mirrors the counters of PE_Inference.py:142-144,154 plus timing/traffic accounting. */
typedef struct vspe_stats {
    uint64_t total_pairs;     /* min(lines_fwd/4, lines_rve/4)            (:154) */
    uint64_t n_pairs;         /* pairs skipped for an upper-case 'N'      (:160) */
    uint64_t short_pairs;     /* pairs skipped for a mate < split_len     (:162) */
    uint64_t used_pairs;      /*                                          (:165) */
    uint64_t bytes_fwd, bytes_rve;
    uint64_t n_nodes;
    uint64_t n_kmers;         /* index insertions = 2 * sum(len - split_len + 1) (:117-135)  */
    uint64_t table_slots;
    uint64_t reads_fast;      /* reads resolved by the seed-and-extend kernel               */
    uint64_t reads_generic;   /* reads resolved by the exhaustive per-position kernel       */
    uint64_t n_keys;          /* (matrix,i,j) increments = sum over used pairs of
                                 L(L+1)/2 + R(R+1)/2 + L*R              (:174-188) */
    uint64_t kernel_launches; /* launches of this library's kernels since vspe_reset        */
    float ms_index;           /* K3 index build                                             */
    float ms_h2d;             /* host->device copies (host-input entry points only)         */
    float ms_scan;            /* K1+K2+first tier of K4: record split, pack, walk           */
    float ms_map;             /* K4 list-driven tiers on the reads the walk left unresolved */
    float ms_count;           /* K5+K6 key emit + radix partition + run-length reduce       */
    float ms_total;           /* device time of the last vspe_count_* call                  */
    float ms_k_scan_rows;     /* sum of k_scan_rows launch durations (CUDA events on its stream)  */
    uint32_t n_k_scan_rows;   /* ... and how many launches that was                               */
    float ms_k_walk;          /* same for the k_memo + k_walk rounds of a chunk (one interval per chunk)      */
    uint32_t n_k_walk;
    uint32_t scan_redo_tiles; /* tiles k_scan_redo packed again because their guessed line phase was wrong  */
    uint32_t reserved0;
    uint64_t reads_memo;      /* reads whose node list came from the read memo (an identical read was walked before) */
} vspe_stats;

typedef struct vspe_ctx vspe_ctx;

const char* vspe_last_error(void);
const char* vspe_version(void);

/* One context = one device + its streams, index and count matrices. */
int vspe_create(int device, vspe_ctx** out);
void vspe_destroy(vspe_ctx* ctx);

/* K3 -- replaces the index build of PE_Inference.py:117-135.
 * seqs: the node sequences concatenated (ASCII, as in the GFA S lines, in S-line order);
 * seq_off[n_nodes + 1]: byte offsets into seqs.  split_len = kmer_size + 1 (:114). */
int vspe_index_build(vspe_ctx* ctx, const uint8_t* seqs, const uint64_t* seq_off,
                     uint32_t n_nodes, uint32_t split_len);

/* Zero the count matrices and pair counters (a fresh run over the same index). */
int vspe_reset(vspe_ctx* ctx);

/* K1+K2+K4+K5+K6 -- replaces the pair loop of PE_Inference.py:147-188 for one pair of FASTQ
 * buffers.  Counts are ADDED to the context's matrices, so disjoint record ranges of the
 * same files may be fed by successive calls (or by different contexts / ranks and summed).
 *   _device: buffers already resident in this context's device memory (device pointers);
 *   _host:   host buffers; pinned cudaMemcpyAsync streaming happens inside. */
int vspe_count_device(vspe_ctx* ctx, const uint8_t* d_fwd, uint64_t n_fwd,
                      const uint8_t* d_rve, uint64_t n_rve);
int vspe_count_host(vspe_ctx* ctx, const uint8_t* fwd, uint64_t n_fwd,
                    const uint8_t* rve, uint64_t n_rve);

/* Device pointers of the dense count matrices: node_mat then short_mat, each
 * n_nodes*n_nodes uint64 row-major, contiguous (one allreduce covers both).
 * (node_mat / short_mat of PE_Inference.py:139-140.) */
int vspe_matrices_device(vspe_ctx* ctx, uint64_t** d_mats, uint64_t* n_elems);
/* Copy them to host (each n_nodes*n_nodes uint64, caller-allocated). */
int vspe_matrices_host(vspe_ctx* ctx, uint64_t* node_mat, uint64_t* short_mat);
int vspe_get_stats(vspe_ctx* ctx, vspe_stats* out);
/* Overwrite the pair counters (after a cross-rank reduction done by the caller). */
int vspe_set_pair_counters(vspe_ctx* ctx, uint64_t total, uint64_t n, uint64_t shrt, uint64_t used);

/* Per-read mapping of ONE FASTQ host buffer -- replaces single_end_read_mapping
 * (PE_Inference.py:16-48) applied to the 2nd line of every complete 4-line record.
 * Outputs are library-allocated, valid until the next call on this context:
 *   status[r]   0 mapped, 1 contains 'N', 2 shorter than split_len
 *   offsets[r]..offsets[r+1]  range into nodes[] (ascending node indices). */
int vspe_map_reads(vspe_ctx* ctx, const uint8_t* fq, uint64_t n_bytes, uint64_t* n_reads,
                   const uint64_t** offsets, const uint32_t** nodes, const uint8_t** status);

/* K1 alone (tests): universal-newline record split of a host buffer.  For every complete
 * record r: seq_start[r], seq_len[r] of its 2nd line (what PE_Inference.py:158 slices). */
int vspe_split_records(vspe_ctx* ctx, const uint8_t* fq, uint64_t n_bytes, uint64_t* n_lines,
                       uint64_t* n_records, const uint64_t** seq_start, const uint32_t** seq_len);

/* Dense writer -- replaces PE_Inference.py:190-207: n*n lines "id_i:id_j:count\n".
 * ids: n NUL-terminated strings. */
int vspe_write_info(const char* path, const char* const* ids, uint32_t n, const uint64_t* mat);

/* Whole drop-in path -- replaces main() of PE_Inference.py:51-211 except argument parsing
 * and the stdout banner: rm -rf out_dir, mkdir, parse GFA S lines, build index, stream both
 * FASTQ files, write out_dir/pe_info and out_dir/st_info.  n_gpus >= 1 shards pairs over
 * devices 0..n_gpus-1 and sums the matrices with one ncclAllReduce. */
int vspe_run(const char* gfa_path, const char* fwd_path, const char* rve_path, int kmer_size,
             const char* out_dir, int n_gpus, vspe_stats* stats);

/* Sparse counting -- for graphs whose N*N matrices cannot exist (dense counting needs
 * 2*N*N <= 2^28 cells, N <= 11 585); switched on automatically for larger graphs or by
 * vspe_set_option(ctx, "sparse", 1).  The context then keeps the non-zero cells as runs sorted by
 * key = mat*N*N + i*N + j (mat 0 = node_mat, 1 = short_mat of PE_Inference.py:139-140):
 * radix sort of the (matrix, i, j) keys + run-length reduce.
 *   vspe_sparse_host   library-owned host copies, valid until the next call on the context
 *   vspe_sparse_merge  add the runs of another context / rank (exact integer sums)
 *   vspe_write_info_sparse  only the non-zero lines "id_i:id_j:count\n" of matrix `mat`; the
 *                      consumer (VStrains_IO.py:598-612) zero-initialises every key, so it parses
 *                      to the same dict as the dense file (opt-in: the bytes differ). */
int vspe_is_sparse(vspe_ctx* ctx);
int vspe_sparse_host(vspe_ctx* ctx, uint64_t* n_entries, const uint64_t** keys, const uint64_t** counts);
int vspe_sparse_merge(vspe_ctx* ctx, const uint64_t* keys, const uint64_t* counts, uint64_t n_entries);
int vspe_write_info_sparse(const char* path, const char* const* ids, uint32_t n, const uint64_t* keys,
                           const uint64_t* counts, uint64_t n_entries, int mat);
/* The same run list in device memory (valid until the next counting / merge call on the context), and
 * a merge of runs that already live on this context's device -- what a multi-GPU caller needs to
 * exchange the runs with ONE all-gather over NVLink instead of a host round trip. */
int vspe_sparse_device(vspe_ctx* ctx, uint64_t* n_entries, uint64_t** d_keys, uint64_t** d_counts);
int vspe_sparse_merge_device(vspe_ctx* ctx, const uint64_t* d_keys, const uint64_t* d_counts, uint64_t n_entries);
/* Forget the context's runs (the counters stay): a rank that hands its runs to their key-range owners in an
 * all-to-all exchange clears its list and merges what it received. */
int vspe_sparse_clear(vspe_ctx* ctx);

/* The CUDA stream (cudaStream_t) every kernel and copy of this context is issued on.  Work a caller
 * enqueues on it (e.g. the NCCL allreduce of vspe_matrices_device) is ordered after the counting calls
 * without a host synchronisation. */
void* vspe_stream(vspe_ctx* ctx);

/* Input files of vspe_run: the bytes of a plain file, or of a gzip file (RFC 1952, concatenated
 * members included) inflated in memory -- detected by the magic bytes.  The reference reads plain
 * text only (PE_Inference.py:105,147-152) and raises on a gzip file, so this only extends the set
 * of accepted inputs.  *data is malloc'ed; release it with
 * vspe_free_input. */
int vspe_read_input(const char* path, uint8_t** data, uint64_t* n_bytes);
void vspe_free_input(uint8_t* data);

/* Pinned host memory for callers that want zero-copy streaming in vspe_count_host. */
void* vspe_alloc_pinned(size_t bytes);
void vspe_free_pinned(void* p);

/* Tunables.  None of them changes a result (every kernel path is exact; the parity tests run the
 * fixtures through each of them); they select kernel paths for tests, experiments and profiling.
 *   "chunk_mb"       host-input streaming: bytes per staged chunk in MiB (default 256)
 *   "sparse"         1: keep sorted (key, count) runs instead of N*N matrices
 *   "scan_mode"      0 (default): k_scan_rows (one pass: TMA tile -> record split -> 2-bit rows) + k_walk +
 *                    list-driven tiers; 1: look-back record scan + raw-byte map kernels (also the fallback
 *                    of mode 0 for chunks with records of a few bytes); 2: two-pass record scan
 *   "scan_two_pass"  1: same as scan_mode 2
 *   "force_generic"  1: every read through the exhaustive ASCII tier (the reference's loop as is)
 *   "subst"          0: do not build / use the substitution-hit bitmap
 *   "link_split"     k_bucket_count: keys of one matrix bucket per CTA before the bucket is shared by several CTAs (default 65536)
 *   "memo"           0: do not use the read memo (every read is walked, as if no read repeated)
 *   "tier_overlap"   0: vspe_count_device maps the two mates strictly one after the other (default 1: the list-driven
 *                    tiers of one mate run on a second stream beside the scan of the other)
 *   "pair_cap_log2"  log2 of the first size of the pair table (default 21 = 32 MB, grown on demand); 0: size it by
 *                    the pairs of a batch
 *   "dbg_counters"   profiling aid (tools/dbg_map.py) */
int vspe_set_option(vspe_ctx* ctx, const char* name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* VSPE_H */
