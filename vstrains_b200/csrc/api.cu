// api.cu -- the C ABI of libvspe.so (declared in include/vspe.h) and the host-side plumbing:
// context, chunked pinned streaming, GFA S-line parsing, dense text writer, whole-run driver.
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "link.cuh"

namespace vspe {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

int map_reads_generic_list(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                           const uint32_t* d_worklist, uint64_t n_items, ReadSlot* d_slots);
int map_reads_fast(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                   uint64_t n_reads, ReadSlot* d_slots);
int map_prepare_lists(Ctx* c, uint64_t n_reads);
int map_reads_deferred(Ctx* c, const uint8_t* d_buf, const uint64_t* d_seq_start, const uint64_t* d_seq_end,
                       const uint32_t* d_rows, const uint32_t* d_hdr, uint32_t row_words, uint32_t cap,
                       uint64_t n_reads_cap, ReadSlot* d_slots, const unsigned long long* d_count);
uint32_t map_fast_cap(uint32_t hint);
int run_multi_gpu(const uint8_t* seqs, const uint64_t* seq_off, uint32_t n_nodes, uint32_t split_len,
                  const uint8_t* fwd, uint64_t n_fwd, const uint8_t* rve, uint64_t n_rve, int n_gpus,
                  std::vector<uint64_t>& node_mat, std::vector<uint64_t>& short_mat,
                  std::vector<uint64_t>* sparse_keys, std::vector<uint64_t>* sparse_counts, vspe_stats* stats);

// ---------------------------------------------------------------------------------------
// per-mate streaming state
// ---------------------------------------------------------------------------------------
struct MateStream {
    uint64_t line_base = 0;     // terminators seen so far
    uint64_t lines = 0;         // total lines once the stream ended
    uint64_t n_slots = 0;       // records with a result slot
};

static inline uint64_t seq_lines_before(uint64_t x) { return (x + 2) / 4; }   // #{l < x : l % 4 == 1}

static int report_error_flags(unsigned long long e);
static int check_kernel_errors(Ctx* c) {
    unsigned long long e = 0;
    VSPE_CUDA(cudaMemcpyAsync(&e, c->counters.p + CNT_ERR, 8, cudaMemcpyDeviceToHost, c->stream));
    VSPE_CUDA(cudaStreamSynchronize(c->stream));
    return report_error_flags(e);
}

static int scan_mode_of(Ctx* c) {
    // 0 (default): k_scan_rows + k_walk + list-driven tiers; 1: look-back record scan + raw-byte map kernels
    // (also the fallback of mode 0); 2: two-pass record scan + raw-byte map kernels (cross-check)
    int mode = (int)c->opt_scan_mode;
    if (c->opt_scan_two_pass) mode = 2;
    if (mode < 0 || mode > 2) mode = 0;
    if ((c->opt_force_generic || c->index.split_len > 320) && mode == 0) mode = 1;
    return mode;
}

// A map / intern launch ran out of private list records or spill words: put the cursors back to
// where the chunk started, clear the flags, grow the pools.  The caller repeats the launch.
static int retry_after_pool_overflow(Ctx* c, unsigned long long flags, const unsigned long long* cur0) {
    const unsigned long long cleared = flags & ~(unsigned long long)(ERRF_LISTS_FULL | ERRF_SPILL_FULL);
    VSPE_CUDA(cudaMemcpyAsync(c->counters.p + CNT_ERR, &cleared, 8, cudaMemcpyHostToDevice, c->stream));
    VSPE_CUDA(cudaMemcpyAsync(c->counters.p + CNT_SPILL_CURSOR, &cur0[0], 8, cudaMemcpyHostToDevice, c->stream));
    VSPE_CUDA(cudaMemcpyAsync(c->counters.p + CNT_OVF, &cur0[1], 8, cudaMemcpyHostToDevice, c->stream));
    VSPE_CUDA(cudaStreamSynchronize(c->stream));
    return link_grow_overflow(c);
}

static inline uint64_t guess_reads(Ctx* c, uint64_t n) {
    return n / (2ull * std::max<uint32_t>(c->read_len_hint, 20) + 8) * 21 / 20 + 4096;
}

// Queue the default path (scan_mode 0) for one chunk: k_scan_rows -> k_walk on the context's stream, then the
// list-driven tiers on what the walk left unresolved and the interning of their slots -- on the same stream, or
// (overlap) on the tier stream, so that they run beside the scan of the other mate.  Nothing is synchronised here.
static int fused_launch(Ctx* c, int m, const uint8_t* d_buf, uint64_t n, uint64_t lb, uint64_t guess, bool overlap) {
    MateBuf& mb = c->mate[m];
    const uint64_t rec_first = seq_lines_before(lb);
    const uint32_t cap = map_fast_cap(c->read_len_hint);
    const uint32_t row_words = cap <= 160 ? 12 : cap <= 256 ? 16 : 20;
    unsigned long long* d_count = c->counters.p + (m == 0 ? CNT_DEFER : CNT_DEFER1);
    VSPE_TRY(mb.handles.reserve(rec_first + guess + 2, true, c->stream));
    VSPE_TRY(mb.slots.reserve(guess + 2));                      // slots of the unresolved reads (compact index)
    VSPE_CUDA(cudaMemsetAsync(d_count, 0, 8, c->stream));
    VSPE_TRY(scan_map(c, m, d_buf, n, lb, rec_first, guess, mb.handles.p + rec_first, d_count, row_words, cap));
    VSPE_CUDA(cudaEventRecord(c->ev_m[m][1], c->stream));
    cudaStream_t main_stream = c->stream;
    if (overlap) {
        VSPE_CUDA(cudaEventRecord(c->ev_walk[m], main_stream));
        VSPE_CUDA(cudaStreamWaitEvent(c->tier_stream, c->ev_walk[m], 0));
        c->stream = c->tier_stream;                             // the tier launchers issue on the context's stream
    }
    c->cur_buf_n = n;
    int rc = map_prepare_lists(c, guess + 2);
    if (rc == VSPE_OK)
        rc = map_reads_deferred(c, d_buf, mb.rec.seq_start.p, mb.rec.seq_end.p, mb.d_rows.p, mb.d_hdr.p, row_words, cap, guess,
                                mb.slots.p, d_count);
    if (rc == VSPE_OK) rc = intern_slots(c, mb.slots.p, guess, c->defer_m[m].p, d_count, mb.handles.p + rec_first);
    if (rc == VSPE_OK && cudaEventRecord(c->ev_m[m][2], c->stream) != cudaSuccess) { set_error("cudaEventRecord failed"); rc = VSPE_ERR_CUDA; }
    if (overlap) {
        if (rc == VSPE_OK && cudaEventRecord(c->ev_tier[m], c->tier_stream) != cudaSuccess) { set_error("cudaEventRecord failed"); rc = VSPE_ERR_CUDA; }
        c->stream = main_stream;
    }
    return rc;
}

// One chunk through the default path (scan_mode 0): k_scan_rows -> k_walk -> list-driven tiers on what the walk left
// unresolved -> their slots interned -> one host sync (terminator count + error flags).  A launch whose
// guessed table size was too small, or that ran out of list records / spill words, is repeated
// (every step is idempotent).  *fell_back: a tile owned more reads than the kernel's table holds
// (records of a few bytes): the caller takes the plain path for this chunk.
static int feed_chunk_fused(Ctx* c, int m, MateStream& ms, const uint8_t* d_buf, uint64_t n, bool* fell_back, uint64_t* n_terms_out,
                            uint64_t* n_seq_out) {
    *fell_back = false;
    const uint64_t lb = ms.line_base, rec_first = seq_lines_before(lb);
    uint64_t guess = guess_reads(c, n);
    VSPE_CUDA(cudaEventRecord(c->ev_m[m][0], c->stream));
    for (int attempt = 0;; attempt++) {
        unsigned long long cur0[2] = {0, 0}, h_rt[2] = {0, 0}, h_err = 0;       // h_rt: {redo tiles, terminators}
        VSPE_CUDA(cudaMemcpyAsync(&cur0[0], c->counters.p + CNT_SPILL_CURSOR, 8, cudaMemcpyDeviceToHost, c->stream));
        VSPE_CUDA(cudaMemcpyAsync(&cur0[1], c->counters.p + CNT_OVF, 8, cudaMemcpyDeviceToHost, c->stream));
        VSPE_TRY(fused_launch(c, m, d_buf, n, lb, guess, false));
        VSPE_CUDA(cudaMemcpyAsync(h_rt, scan_map_total_ptr(c, m, n, d_buf), 16, cudaMemcpyDeviceToHost, c->stream));
        VSPE_CUDA(cudaMemcpyAsync(&h_err, c->counters.p + CNT_ERR, 8, cudaMemcpyDeviceToHost, c->stream));
        VSPE_CUDA(cudaStreamSynchronize(c->stream));
        scan_map_account(c);
        const unsigned long long h_total = h_rt[1];
        if (!(h_err & ERRF_TILE_FULL)) c->stats.scan_redo_tiles += (uint32_t)h_rt[0];
        const uint64_t n_seq = seq_lines_before(lb + h_total) - rec_first;
        const unsigned long long transient = ERRF_SLOTS_FULL | ERRF_TILE_FULL;
        if (h_err & transient) {
            const unsigned long long cleared = h_err & ~transient;
            VSPE_CUDA(cudaMemcpyAsync(c->counters.p + CNT_ERR, &cleared, 8, cudaMemcpyHostToDevice, c->stream));
            VSPE_CUDA(cudaStreamSynchronize(c->stream));
        }
        *n_terms_out = h_total;
        *n_seq_out = n_seq;
        if (h_err & ERRF_TILE_FULL) { *fell_back = true; return VSPE_OK; }
        if (attempt < 8 && (h_err & (ERRF_LISTS_FULL | ERRF_SPILL_FULL))) {
            VSPE_TRY(retry_after_pool_overflow(c, h_err, cur0));
            continue;
        }
        if (attempt < 8 && n_seq > guess) { guess = n_seq; continue; }          // the guessed table was too small
        break;
    }
    return VSPE_OK;
}

// One chunk of one mate's byte stream, resident on the device.  The chunk must start at a line
// start and (unless it is the last chunk) end right after a terminator.
static int feed_chunk(Ctx* c, int m, MateStream& ms, const uint8_t* d_buf, uint64_t n, bool is_last, int last_byte,
                      bool sync_after = true) {
    if (n == 0) {
        if (is_last) ms.lines = ms.line_base;
        return VSPE_OK;
    }
    MateBuf& mb = c->mate[m];
    uint64_t n_terms = 0;
    for (auto& e : c->ev_m[m]) if (!e) VSPE_CUDA(cudaEventCreate(&e));
    cudaEvent_t e0 = c->ev_m[m][0], e1 = c->ev_m[m][1], e2 = c->ev_m[m][2];
    const uint64_t lb = ms.line_base;
    const uint64_t rec_first = seq_lines_before(lb);
    uint64_t n_seq = 0;
    c->cur_buf_n = n;
    int mode = scan_mode_of(c);
    auto finish = [&]() {
        ms.n_slots = rec_first + n_seq;
        ms.line_base = lb + n_terms;
        if (is_last) {
            bool term = last_byte == '\n' || last_byte == '\r';
            ms.lines = ms.line_base + (term ? 0 : 1);
        }
    };
    if (mode == 0) {
        bool fell_back = false;
        VSPE_TRY(feed_chunk_fused(c, m, ms, d_buf, n, &fell_back, &n_terms, &n_seq));
        if (!fell_back) {
            if (sync_after) {                             // (streaming callers: account the stage times now)
                VSPE_CUDA(cudaStreamSynchronize(c->stream));
                float a = 0, b = 0;
                cudaEventElapsedTime(&a, e0, e1);
                cudaEventElapsedTime(&b, e1, e2);
                c->stats.ms_scan += a;
                c->stats.ms_map += b;
            }
            finish();
            return VSPE_OK;
        }
        mode = 1;                                         // a tile owned more reads than its table holds: plain path
    }
    VSPE_CUDA(cudaEventRecord(e0, c->stream));
    if (mode == 2) {
        VSPE_TRY(scan_count_lines(c, d_buf, n, &n_terms));
        n_seq = seq_lines_before(lb + n_terms) - rec_first;
        VSPE_TRY(mb.rec.seq_start.reserve(n_seq + 2));
        VSPE_TRY(mb.rec.seq_end.reserve(n_seq + 2));
        VSPE_TRY(scan_index_records(c, d_buf, n, lb, rec_first, n_seq, mb.rec.seq_start.p, mb.rec.seq_end.p));
    } else {
        // one pass with a guessed table size (a FASTQ record is rarely under 48 bytes); if the
        // guess was too small the pass is repeated once with the exact size
        uint64_t guess = n / 48 + 1024;
        bool overflow = false;
        VSPE_TRY(mb.rec.seq_start.reserve(guess + 2));
        VSPE_TRY(mb.rec.seq_end.reserve(guess + 2));
        VSPE_TRY(scan_records_single_pass(c, d_buf, n, lb, rec_first, mb.rec.seq_start.cap - 2, mb.rec.seq_start.p, mb.rec.seq_end.p,
                                          &n_terms, &overflow));
        n_seq = seq_lines_before(lb + n_terms) - rec_first;
        if (overflow) {
            VSPE_TRY(mb.rec.seq_start.reserve(n_seq + 2));
            VSPE_TRY(mb.rec.seq_end.reserve(n_seq + 2));
            VSPE_TRY(scan_records_single_pass(c, d_buf, n, lb, rec_first, n_seq, mb.rec.seq_start.p, mb.rec.seq_end.p, &n_terms, &overflow));
        }
    }
    // the raw-byte map kernels write one slot per read (chunk-local index), which is then interned
    VSPE_TRY(mb.slots.reserve(n_seq + 1));
    VSPE_CUDA(cudaEventRecord(e1, c->stream));
    unsigned long long cur0[2] = {0, 0};                  // spill / private-record cursors before this chunk's map stage
    VSPE_CUDA(cudaMemcpyAsync(&cur0[0], c->counters.p + CNT_SPILL_CURSOR, 8, cudaMemcpyDeviceToHost, c->stream));
    VSPE_CUDA(cudaMemcpyAsync(&cur0[1], c->counters.p + CNT_OVF, 8, cudaMemcpyDeviceToHost, c->stream));
    auto map_all = [&]() -> int {
        if (c->opt_force_generic) return map_reads_generic_list(c, d_buf, mb.rec.seq_start.p, mb.rec.seq_end.p, nullptr, n_seq, mb.slots.p);
        return map_reads_fast(c, d_buf, mb.rec.seq_start.p, mb.rec.seq_end.p, n_seq, mb.slots.p);
    };
    if (n_seq) VSPE_TRY(map_all());
    // slots -> list handles (link.cuh); a launch that ran out of private list records or spill words
    // is repeated after growing the pools (interning is idempotent)
    VSPE_TRY(mb.handles.reserve(rec_first + n_seq + 1, true, c->stream));
    for (int attempt = 0; n_seq; attempt++) {
        // the slots of this chunk keep their spill ranges: interning alone is repeated from the cursors as the map stage left them
        unsigned long long cur1[2] = {0, 0}, h_err = 0;
        VSPE_CUDA(cudaMemcpyAsync(&cur1[0], c->counters.p + CNT_SPILL_CURSOR, 8, cudaMemcpyDeviceToHost, c->stream));
        VSPE_CUDA(cudaMemcpyAsync(&cur1[1], c->counters.p + CNT_OVF, 8, cudaMemcpyDeviceToHost, c->stream));
        VSPE_TRY(intern_slots(c, mb.slots.p, n_seq, nullptr, nullptr, mb.handles.p + rec_first));
        VSPE_CUDA(cudaMemcpyAsync(&h_err, c->counters.p + CNT_ERR, 8, cudaMemcpyDeviceToHost, c->stream));
        VSPE_CUDA(cudaStreamSynchronize(c->stream));
        if (!(h_err & (ERRF_LISTS_FULL | ERRF_SPILL_FULL))) break;
        if (attempt == 8) break;                          // reported by the caller's error check
        const bool remap = (h_err & ERRF_SPILL_FULL) != 0;
        VSPE_TRY(retry_after_pool_overflow(c, h_err, remap ? cur0 : cur1));
        if (remap) VSPE_TRY(map_all());                   // spill words ran out (possibly inside the map tiers): map again
    }
    VSPE_CUDA(cudaEventRecord(e2, c->stream));
    if (sync_after) {
        // streaming callers reuse the chunk buffer next: wait, and account the stage times now
        VSPE_CUDA(cudaStreamSynchronize(c->stream));
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, e0, e1);
        cudaEventElapsedTime(&b, e1, e2);
        c->stats.ms_scan += a;
        c->stats.ms_map += b;
    }
    finish();
    return VSPE_OK;
}

static int report_error_flags(unsigned long long e) {
    if (e & ERRF_NON_ASCII) { set_error("input contains a byte >= 0x80 (non-ASCII FASTQ is outside the reference's contract)"); return VSPE_ERR_NON_ASCII; }
    if (e & (ERRF_SPILL_FULL | ERRF_LISTS_FULL)) { set_error("node-list pools exhausted after repeated growth"); return VSPE_ERR_LIMIT; }
    if (e & ERRF_INTERNAL) { set_error("internal: a read reached the count stage without a node list"); return VSPE_ERR_LIMIT; }
    if (e & ERRF_KEYS_FULL) { set_error("key buffer exhausted"); return VSPE_ERR_LIMIT; }
    if (e & ERRF_TILE_FULL) { set_error("internal: a scan tile overflowed after its count pass accepted it"); return VSPE_ERR_LIMIT; }
    return VSPE_OK;
}

static int finish_pairs(Ctx* c, const MateStream& f, const MateStream& r) {
    uint64_t total = std::min(f.lines / 4, r.lines / 4);       // PE_Inference.py:154
    cudaEvent_t e0 = c->ev[5], e1 = c->ev[6];
    VSPE_CUDA(cudaEventRecord(e0, c->stream));
    c->err_flags_fresh = false;
    VSPE_TRY(count_links(c, c->mate[0].handles.p, c->mate[1].handles.p, total));
    VSPE_CUDA(cudaEventRecord(e1, c->stream));
    unsigned long long h_fast = 0, h_hit = 0;
    VSPE_CUDA(cudaMemcpyAsync(&h_fast, c->counters.p + CNT_FAST, 8, cudaMemcpyDeviceToHost, c->stream));
    VSPE_CUDA(cudaMemcpyAsync(&h_hit, c->counters.p + CNT_MEMO_HIT, 8, cudaMemcpyDeviceToHost, c->stream));
    VSPE_CUDA(cudaStreamSynchronize(c->stream));
    // an input that does not repeat reads (shallow coverage of a large graph) only pays for the memo: stop asking it
    if (c->opt_memo && !c->memo_off && h_fast >= (4ull << 20) && h_hit * 20 < h_fast) c->memo_off = true;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    c->stats.ms_count += ms;
    // count_pairs reads the kernels' error flags together with its key count; without pairs
    // (empty inputs) they are fetched here
    if (c->err_flags_fresh) return report_error_flags(c->last_err_flags);
    return check_kernel_errors(c);
}

static int require_index(Ctx* c) {
    if (!c || !c->index.built) { set_error("vspe_index_build must succeed before this call"); return VSPE_ERR_ARG; }
    VSPE_CUDA(cudaSetDevice(c->device));
    return VSPE_OK;
}

// ---------------------------------------------------------------------------------------
// host-side helpers
// ---------------------------------------------------------------------------------------
// Largest cut <= hi such that [lo, cut) ends right after a line terminator and does not
// split a "\r\n"; returns lo if the range holds no usable terminator.
static uint64_t cut_at_line(const uint8_t* p, uint64_t lo, uint64_t hi, uint64_t n) {
    uint64_t i = hi;
    while (i > lo) {
        uint8_t c = p[i - 1];
        if (c == '\n') return i;
        if (c == '\r' && i < n && p[i] != '\n') return i;
        i--;
    }
    return lo;
}

// Longest 2nd-line length among the records of a FASTQ sample (sizes the packed rows of the walk and
// seed-and-extend tiers; a wrong hint only sends longer reads to the exhaustive tier).  `aligned`: the
// sample starts at a line start, so line numbers are known; otherwise (a window from the middle of a
// buffer) every line is a candidate except those that start with '@' or '+' -- an upper bound is all
// that is needed.
static uint32_t seq_len_hint(const uint8_t* p, uint64_t n, bool aligned = true) {
    uint64_t line = 0, start = 0, best = 0;
    bool first = true;
    for (uint64_t i = 0; i < n && line < 64; i++) {
        uint8_t c = p[i];
        bool term = c == '\n' || (c == '\r' && !(i + 1 < n && p[i + 1] == '\n'));
        if (!term) continue;
        const bool count = aligned ? (line & 3) == 1 : (!first && p[start] != '@' && p[start] != '+');
        if (count) {
            uint64_t e = (c == '\n' && i > start && p[i - 1] == '\r') ? i - 1 : i;
            best = std::max(best, e - start);
        }
        first = false;
        line++;
        start = i + 1;
    }
    return (uint32_t)std::min<uint64_t>(best, 1u << 20);
}

// the hint over a whole host buffer: its head plus windows at 1/4, 1/2, 3/4 and the tail, so that inputs whose
// first records are short (adapter-trimmed reads) still get rows long enough for the rest
static uint32_t seq_len_hint_sampled(const uint8_t* p, uint64_t n) {
    const uint64_t W = 16384;
    uint32_t best = seq_len_hint(p, std::min(n, W));
    for (int q = 1; q <= 4 && n > 2 * W; q++) {
        const uint64_t at = q == 4 ? n - W : n / 4 * q;
        best = std::max(best, seq_len_hint(p + at, std::min(W, n - at), false));
    }
    return best;
}

static void parallel_memcpy(uint8_t* dst, const uint8_t* src, size_t n, unsigned max_threads) {
    unsigned nt = std::min(std::max(1u, max_threads), std::max(1u, std::thread::hardware_concurrency()));
    if (n < (8u << 20) || nt == 1) { memcpy(dst, src, n); return; }
    std::vector<std::thread> th;
    size_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        size_t a = (size_t)t * per, b = std::min(n, a + per);
        if (a >= b) break;
        th.emplace_back([=] { memcpy(dst + a, src + a, b - a); });
    }
    for (auto& t : th) t.join();
}

static bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// Stream both mates' host buffers through the device in line-aligned chunks with two staging
// buffers: the copy of chunk i+1 overlaps the kernels of chunk i, and the first chunk of the
// second mate is already on its way while the last chunk of the first mate is processed.
static int stream_mates_host(Ctx* c, MateStream* ms, const uint8_t* const* srcs, const uint64_t* ns) {
    const uint64_t chunk = (uint64_t)std::max<int64_t>(1, c->opt_chunk_mb) << 20;
    struct Piece { int m; uint64_t lo, hi; bool last; };
    std::vector<Piece> pieces;
    bool pinned_src[2] = {true, true};
    for (int m = 0; m < 2; m++) {
        const uint8_t* src = srcs[m];
        const uint64_t n = ns[m];
        if (n == 0) { VSPE_TRY(feed_chunk(c, m, ms[m], nullptr, 0, true, -1)); continue; }
        pinned_src[m] = is_pinned(src);
        for (uint64_t lo = 0; lo < n;) {
            uint64_t hi = std::min(n, lo + chunk);
            if (hi < n) {
                uint64_t cut = cut_at_line(src, lo, hi, n);
                while (cut == lo && hi < n) {                 // a line longer than the chunk: extend
                    hi = std::min(n, hi + chunk);
                    cut = hi == n ? n : cut_at_line(src, lo, hi, n);
                }
                hi = cut;
            }
            pieces.push_back({m, lo, hi, hi == n});
            lo = hi;
        }
    }
    if (pieces.empty()) return VSPE_OK;
    uint64_t max_piece = 0;
    for (auto& p : pieces) max_piece = std::max(max_piece, p.hi - p.lo);
    for (int b = 0; b < 2; b++) VSPE_TRY(c->dev_in[b].reserve(max_piece + 64));
    if (!(pinned_src[0] && pinned_src[1]) && c->pinned_bytes < max_piece) {
        for (int b = 0; b < 2; b++) {
            if (c->pinned[b]) cudaFreeHost(c->pinned[b]);
            c->pinned[b] = nullptr;
            VSPE_CUDA(cudaMallocHost(&c->pinned[b], max_piece + 64));
        }
        c->pinned_bytes = max_piece;
    }
    cudaEvent_t copied[2];
    for (int b = 0; b < 2; b++) VSPE_CUDA(cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming));
    auto issue_copy = [&](size_t i) -> int {
        int b = (int)(i & 1);
        const Piece& p = pieces[i];
        const uint8_t* from = srcs[p.m] + p.lo;
        if (!pinned_src[p.m]) { parallel_memcpy(c->pinned[b], from, p.hi - p.lo, (unsigned)c->opt_stage_threads); from = c->pinned[b]; }
        VSPE_CUDA(cudaMemcpyAsync(c->dev_in[b].p, from, p.hi - p.lo, cudaMemcpyHostToDevice, c->copy_stream[b]));
        VSPE_CUDA(cudaEventRecord(copied[b], c->copy_stream[b]));
        return VSPE_OK;
    };
    int rc = VSPE_OK;
    rc = issue_copy(0);
    for (size_t i = 0; rc == VSPE_OK && i < pieces.size(); i++) {
        int b = (int)(i & 1);
        // chunk i-1 (other buffer) was fully processed (feed_chunk syncs), so buffer b^1 is free
        if (i + 1 < pieces.size()) rc = issue_copy(i + 1);
        if (rc != VSPE_OK) break;
        if (cudaStreamWaitEvent(c->stream, copied[b], 0) != cudaSuccess) { set_error("cudaStreamWaitEvent failed"); rc = VSPE_ERR_CUDA; break; }
        const Piece& p = pieces[i];
        rc = feed_chunk(c, p.m, ms[p.m], c->dev_in[b].p, p.hi - p.lo, p.last, p.last ? srcs[p.m][ns[p.m] - 1] : -1);
    }
    cudaStreamSynchronize(c->copy_stream[0]);
    cudaStreamSynchronize(c->copy_stream[1]);
    cudaStreamSynchronize(c->stream);
    for (int b = 0; b < 2; b++) cudaEventDestroy(copied[b]);
    return rc;
}

// Python text-mode line iteration + `Line[:-1]` (PE_Inference.py:105-106) over raw bytes.
struct GfaNodes {
    std::vector<std::string> ids;
    std::vector<uint8_t> seqs;
    std::vector<uint64_t> off{0};
};

static int parse_gfa_bytes(const uint8_t* g, uint64_t n, GfaNodes& out) {
    uint64_t s = 0, i = 0;
    auto handle = [&](uint64_t a, uint64_t b) -> int {     // content [a, b)
        if (b <= a || g[a] != 'S') return VSPE_OK;
        if (b - a > 1 && g[a + 1] != '\t') return VSPE_OK;  // first field is not exactly "S"
        uint64_t p = a + 1;
        if (p >= b) { set_error("GFA S line without an id field"); return VSPE_ERR_GFA; }
        uint64_t f1s = p + 1;
        p = f1s;
        while (p < b && g[p] != '\t') p++;
        uint64_t f1e = p;
        if (p >= b) { set_error("GFA S line without a sequence field"); return VSPE_ERR_GFA; }
        uint64_t f2s = p + 1;
        p = f2s;
        while (p < b && g[p] != '\t') p++;
        out.ids.emplace_back(reinterpret_cast<const char*>(g + f1s), f1e - f1s);
        out.seqs.insert(out.seqs.end(), g + f2s, g + p);
        out.off.push_back(out.seqs.size());
        return VSPE_OK;
    };
    while (i < n) {
        uint8_t ch = g[i];
        if (ch & 0x80) { set_error("GFA contains a non-ASCII byte"); return VSPE_ERR_NON_ASCII; }
        if (ch == '\n' || ch == '\r') {
            VSPE_TRY(handle(s, i));
            if (ch == '\r' && i + 1 < n && g[i + 1] == '\n') i++;
            i++;
            s = i;
        } else {
            i++;
        }
    }
    if (s < n) VSPE_TRY(handle(s, n - 1));      // unterminated last line: [:-1] eats a real char
    return VSPE_OK;
}

struct MappedFile {
    const uint8_t* p = nullptr;
    uint64_t n = 0;
    int fd = -1;
    ~MappedFile() {
        if (p && n) munmap(const_cast<uint8_t*>(p), n);
        if (fd >= 0) close(fd);
    }
    int open_ro(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) { set_error("cannot open %s: %s", path, strerror(errno)); return VSPE_ERR_IO; }
        struct stat st;
        if (fstat(fd, &st) != 0) { set_error("cannot stat %s", path); return VSPE_ERR_IO; }
        n = (uint64_t)st.st_size;
        if (n) {
            void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { set_error("cannot mmap %s: %s", path, strerror(errno)); p = nullptr; return VSPE_ERR_IO; }
            madvise(m, n, MADV_SEQUENTIAL);
            p = static_cast<const uint8_t*>(m);
        }
        return VSPE_OK;
    }
};

// gzip (RFC 1952) -> bytes; concatenated members (bgzip, `cat a.gz b.gz`) are one stream
static int inflate_gzip(const uint8_t* src, uint64_t n, std::vector<uint8_t>& out, const char* path) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 15 + 16) != Z_OK) { set_error("zlib inflateInit2 failed"); return VSPE_ERR_IO; }
    out.clear();
    out.resize((size_t)std::max<uint64_t>(n * 4, 1u << 16));
    uint64_t in_pos = 0, out_pos = 0;
    int rc = VSPE_OK;
    while (true) {
        if (zs.avail_in == 0 && in_pos < n) {
            const uint64_t part = std::min<uint64_t>(n - in_pos, 1u << 30);
            zs.next_in = const_cast<Bytef*>(src + in_pos);
            zs.avail_in = (uInt)part;
            in_pos += part;
        }
        if (out_pos == out.size()) out.resize(out.size() * 2);
        const uint64_t room = std::min<uint64_t>(out.size() - out_pos, 1u << 30);
        zs.next_out = out.data() + out_pos;
        zs.avail_out = (uInt)room;
        const int z = inflate(&zs, Z_NO_FLUSH);
        out_pos += room - zs.avail_out;
        if (z == Z_STREAM_END) {
            const uint64_t left = (uint64_t)zs.avail_in + (n - in_pos);
            if (left == 0) break;
            // another member follows (anything else after the trailer is an error, as for gzip -d)
            if (inflateReset(&zs) != Z_OK) { set_error("zlib inflateReset failed on %s", path); rc = VSPE_ERR_IO; break; }
            continue;
        }
        if (z == Z_OK) continue;
        if (z == Z_BUF_ERROR && zs.avail_in == 0 && in_pos == n) { set_error("%s: truncated gzip stream", path); rc = VSPE_ERR_IO; break; }
        if (z == Z_BUF_ERROR) continue;
        set_error("%s: corrupt gzip stream (%s)", path, zs.msg ? zs.msg : "zlib error");
        rc = VSPE_ERR_IO;
        break;
    }
    inflateEnd(&zs);
    if (rc == VSPE_OK) out.resize(out_pos);
    return rc;
}

// An input file of the CLI: the bytes of a plain file (mapped) or of a gzip file (inflated into
// memory).  The reference opens plain text only (PE_Inference.py:105,147-152); a file that starts
// with the gzip magic would make it raise UnicodeDecodeError, so accepting it is a pure extension
//  and never changes the result for an input the reference accepts.
struct InputFile {
    MappedFile map;
    std::vector<uint8_t> mem;
    const uint8_t* p = nullptr;
    uint64_t n = 0;
    bool gz = false;          // map holds a gzip stream (p / n are only valid after inflate_all)
    std::string path;
    // inflate_now = false keeps a gzip file compressed: the single-GPU run streams it (GzipStream below)
    int open_ro(const char* path_, bool inflate_now = true) {
        path = path_;
        VSPE_TRY(map.open_ro(path_));
        gz = map.n >= 18 && map.p[0] == 0x1f && map.p[1] == 0x8b && map.p[2] == 8;
        if (gz && inflate_now) return inflate_all();
        if (!gz) { p = map.p; n = map.n; }
        return VSPE_OK;
    }
    int inflate_all() {
        VSPE_TRY(inflate_gzip(map.p, map.n, mem, path.c_str()));
        p = mem.data();
        n = mem.size();
        return VSPE_OK;
    }
};

// ---------------------------------------------------------------------------------------
// Streaming ingest of gzip inputs: one inflate thread per read file fills line-aligned
// chunks in pinned host memory while the GPU works on the previous ones; the two files inflate side by
// side and are fed alternately, so a C4-size .fastq.gz pair needs two chunks of host memory per file
// instead of both files inflated in RAM.
// ---------------------------------------------------------------------------------------
struct GzipStream {
    z_stream zs;
    const uint8_t* src = nullptr;
    uint64_t n = 0, in_pos = 0;
    bool open = false, finished = false;
    std::string path;
    int begin(const uint8_t* p, uint64_t len, const char* path_) {
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, 15 + 16) != Z_OK) { set_error("zlib inflateInit2 failed"); return VSPE_ERR_IO; }
        open = true; src = p; n = len; path = path_;
        return VSPE_OK;
    }
    ~GzipStream() { if (open) inflateEnd(&zs); }
    // up to cap bytes into dst; *got = 0 only at the end of the stream
    int read(uint8_t* dst, uint64_t cap, uint64_t* got) {
        *got = 0;
        while (!finished && *got < cap) {
            if (zs.avail_in == 0 && in_pos < n) {
                const uint64_t part = std::min<uint64_t>(n - in_pos, 1u << 30);
                zs.next_in = const_cast<Bytef*>(src + in_pos);
                zs.avail_in = (uInt)part;
                in_pos += part;
            }
            const uint64_t room = std::min<uint64_t>(cap - *got, 1u << 30);
            zs.next_out = dst + *got;
            zs.avail_out = (uInt)room;
            const int z = inflate(&zs, Z_NO_FLUSH);
            *got += room - zs.avail_out;
            if (z == Z_STREAM_END) {
                if ((uint64_t)zs.avail_in + (n - in_pos) == 0) { finished = true; break; }
                if (inflateReset(&zs) != Z_OK) { set_error("zlib inflateReset failed on %s", path.c_str()); return VSPE_ERR_IO; }   // next member
                continue;
            }
            if (z == Z_OK) continue;
            if (z == Z_BUF_ERROR && zs.avail_in == 0 && in_pos == n) { set_error("%s: truncated gzip stream", path.c_str()); return VSPE_ERR_IO; }
            if (z == Z_BUF_ERROR) continue;
            set_error("%s: corrupt gzip stream (%s)", path.c_str(), zs.msg ? zs.msg : "zlib error");
            return VSPE_ERR_IO;
        }
        return VSPE_OK;
    }
};

// Producer of line-aligned chunks of one inflating file: two pinned buffers, filled by its own thread.
struct ChunkProducer {
    GzipStream gz;
    uint8_t* buf[2] = {nullptr, nullptr};
    uint64_t cap = 0, len[2] = {0, 0};
    bool last[2] = {false, false};
    int state[2] = {0, 0};                 // 0 free, 1 filled
    std::vector<uint8_t> carry;            // bytes after the last terminator of the previous chunk
    std::mutex mu;
    std::condition_variable cv;
    std::thread th;
    int rc = VSPE_OK;
    std::string err;
    bool stop = false;
    ~ChunkProducer() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        if (th.joinable()) th.join();
        for (auto& b : buf) if (b) cudaFreeHost(b);
    }
    int start(const InputFile& in, uint64_t chunk) {
        cap = chunk;
        for (auto& b : buf) if (cudaMallocHost(&b, cap + 64) != cudaSuccess) { set_error("cudaMallocHost(%llu) failed", (unsigned long long)cap); return VSPE_ERR_CUDA; }
        VSPE_TRY(gz.begin(in.map.p, in.map.n, in.path.c_str()));
        th = std::thread([this] { run(); });
        return VSPE_OK;
    }
    void run() {
        int b = 0;
        bool eof = false;
        while (!eof) {
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return state[b] == 0 || stop; }); if (stop) return; }
            uint64_t have = carry.size();
            if (have > cap) { fail(VSPE_ERR_LIMIT, "a line of the gzip input is longer than the streaming chunk (raise chunk_mb)"); return; }
            if (have) memcpy(buf[b], carry.data(), have);
            carry.clear();
            uint64_t got = 0;
            const int r = gz.read(buf[b] + have, cap - have, &got);
            if (r != VSPE_OK) { fail(r, get_error()); return; }
            have += got;
            eof = gz.finished || got == 0;
            uint64_t cut = have;
            if (!eof) {
                // largest cut right after a terminator that cannot be the '\r' of a "\r\n" split by the chunk end
                cut = have;
                while (cut > 0) {
                    const uint8_t ch = buf[b][cut - 1];
                    if (ch == '\n') break;
                    if (ch == '\r' && cut < have && buf[b][cut] != '\n') break;
                    cut--;
                }
                if (cut == 0) { fail(VSPE_ERR_LIMIT, "a line of the gzip input is longer than the streaming chunk (raise chunk_mb)"); return; }
                carry.assign(buf[b] + cut, buf[b] + have);
            }
            { std::lock_guard<std::mutex> lk(mu); len[b] = cut; last[b] = eof; state[b] = 1; }
            cv.notify_all();
            b ^= 1;
        }
    }
    void fail(int code, const char* msg) {
        { std::lock_guard<std::mutex> lk(mu); rc = code; err = msg; state[0] = state[1] = 1; len[0] = len[1] = 0; last[0] = last[1] = true; }
        cv.notify_all();
    }
    // blocks until buffer b is filled
    void wait_filled(int b) { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return state[b] == 1; }); }
    void release(int b) { { std::lock_guard<std::mutex> lk(mu); state[b] = 0; } cv.notify_all(); }
};

static inline char* put_u64(char* p, uint64_t v) {
    char tmp[24];
    int k = 0;
    do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (k) *p++ = tmp[--k];
    return p;
}

}  // namespace vspe

using namespace vspe;

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

const char* vspe_last_error(void) { return get_error(); }
const char* vspe_version(void) { return "vspe-b200 0.1 (sm_100a)"; }

int vspe_create(int device, vspe_ctx** out) {
    if (!out) { set_error("null out pointer"); return VSPE_ERR_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s); libvspe has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return VSPE_ERR_CUDA;
    }
    if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return VSPE_ERR_ARG; }
    VSPE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VSPE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libvspe is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return VSPE_ERR_CUDA;
    }
    vspe_ctx* c = new vspe_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    const int rc = [&]() -> int {
        VSPE_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) VSPE_CUDA(cudaStreamCreateWithFlags(&c->copy_stream[b], cudaStreamNonBlocking));
        {   // the list-driven tiers are short kernels that run beside a device-filling scan: with the higher
            // priority their blocks take the SM slots the scan's blocks free, instead of queueing behind its whole grid
            int prio_lo = 0, prio_hi = 0;
            VSPE_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            const char* e = getenv("VSPE_TIER_PRIO");
            VSPE_CUDA(cudaStreamCreateWithPriority(&c->tier_stream, cudaStreamNonBlocking, (e && e[0] == '0') ? prio_lo : prio_hi));
        }
        for (int b = 0; b < 2; b++) {
            VSPE_CUDA(cudaEventCreateWithFlags(&c->ev_walk[b], cudaEventDisableTiming));
            VSPE_CUDA(cudaEventCreateWithFlags(&c->ev_tier[b], cudaEventDisableTiming));
        }
        for (auto& ev : c->ev) VSPE_CUDA(cudaEventCreate(&ev));
        VSPE_TRY(c->counters.reserve(CNT_COUNT_));
        VSPE_CUDA(cudaMemset(c->counters.p, 0, c->counters.cap * 8));
        return VSPE_OK;
    }();
    if (rc != VSPE_OK) { vspe_destroy(c); return rc; }         // nothing of a half-built context is leaked
    *out = c;
    return VSPE_OK;
}

void vspe_destroy(vspe_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int b = 0; b < 2; b++) {
        if (c->pinned[b]) cudaFreeHost(c->pinned[b]);
        if (c->copy_stream[b]) cudaStreamDestroy(c->copy_stream[b]);
    }
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    for (auto& evs : c->ev_scan) for (auto& ev : evs) if (ev) cudaEventDestroy(ev);
    for (auto& evs : c->ev_m) for (auto& ev : evs) if (ev) cudaEventDestroy(ev);
    for (int b = 0; b < 2; b++) { if (c->ev_walk[b]) cudaEventDestroy(c->ev_walk[b]); if (c->ev_tier[b]) cudaEventDestroy(c->ev_tier[b]); }
    if (c->tier_stream) cudaStreamDestroy(c->tier_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int vspe_index_build(vspe_ctx* c, const uint8_t* seqs, const uint64_t* seq_off, uint32_t n_nodes, uint32_t split_len) {
    if (!c || !seq_off || (n_nodes && !seqs && seq_off[n_nodes])) { set_error("bad arguments"); return VSPE_ERR_ARG; }
    VSPE_CUDA(cudaSetDevice(c->device));
    c->scratch_valid = false;
    VSPE_TRY(index_build_device(c, seqs, seq_off, n_nodes, split_len));
    VSPE_TRY(link_setup(c));
    return vspe_reset(c);
}

int vspe_reset(vspe_ctx* c) {
    VSPE_TRY(require_index(c));
    uint64_t nn = 2ull * c->index.n_nodes * c->index.n_nodes;
    if (nn && !c->sparse.enabled) VSPE_CUDA(cudaMemsetAsync(c->mats.p, 0, nn * 8, c->stream));
    c->sparse.n_runs = 0;
    VSPE_TRY(link_reset(c));
    VSPE_CUDA(cudaMemsetAsync(c->counters.p, 0, CNT_COUNT_ * 8, c->stream));   // stream-ordered: no host sync needed
    vspe_stats keep = c->stats;
    c->stats = {};
    c->stats_overridden = false;
    c->stats.ms_index = keep.ms_index;
    c->stats.n_nodes = keep.n_nodes;
    c->stats.n_kmers = keep.n_kmers;
    c->stats.table_slots = keep.table_slots;
    c->launches = 0;
    return VSPE_OK;
}

static void begin_call(vspe_ctx* c) {
    // the spill pool is per call: slots of earlier calls are no longer referenced
    cudaMemsetAsync(c->counters.p + CNT_SPILL_CURSOR, 0, 8, c->stream);
    cudaMemsetAsync(c->counters.p + CNT_OVF, 0, 8, c->stream);      // ... and so are the private list records
}

int vspe_count_device(vspe_ctx* c, const uint8_t* d_fwd, uint64_t n_fwd, const uint8_t* d_rve, uint64_t n_rve) {
    VSPE_TRY(require_index(c));
    begin_call(c);
    cudaEvent_t t0 = c->ev[0], t1 = c->ev[1];
    VSPE_CUDA(cudaEventRecord(t0, c->stream));
    MateStream f, r;
    const uint8_t* bufs[2] = {d_fwd, d_rve};
    uint64_t ns[2] = {n_fwd, n_rve};
    MateStream* ms[2] = {&f, &r};
    int last[2] = {-1, -1};
    uint32_t hint = 0;
    {   // one small D2H round for both mates: last byte (does the file end with a terminator?)
        // and a prefix to size the packed rows
        static thread_local std::vector<uint8_t> head(2 * 16384);
        uint8_t lastb[2] = {0, 0};
        uint64_t hn[2] = {0, 0};
        for (int m = 0; m < 2; m++) {
            if (!ns[m]) continue;
            hn[m] = std::min<uint64_t>(ns[m], 16384);
            VSPE_CUDA(cudaMemcpyAsync(&lastb[m], bufs[m] + ns[m] - 1, 1, cudaMemcpyDeviceToHost, c->stream));
            VSPE_CUDA(cudaMemcpyAsync(head.data() + 16384 * m, bufs[m], hn[m], cudaMemcpyDeviceToHost, c->stream));
        }
        VSPE_CUDA(cudaStreamSynchronize(c->stream));
        for (int m = 0; m < 2; m++) {
            if (!ns[m]) continue;
            last[m] = lastb[m];
            hint = std::max(hint, seq_len_hint(head.data() + 16384 * m, hn[m]));
        }
    }
    c->read_len_hint = hint;
    // Default path, both mates non-empty: the two mates are queued back to back -- the list-driven tiers of the
    // first one run on the tier stream beside the scan of the second one -- and collected with ONE host sync.
    // Anything out of the ordinary (a guess that was too small, pools to grow, tiles with too many records)
    // repeats the attempt or takes the serial per-mate path below.
    bool done = false;
    if (scan_mode_of(c) == 0 && ns[0] && ns[1] && c->opt_tier_overlap) {
        uint64_t guess[2] = {guess_reads(c, ns[0]), guess_reads(c, ns[1])};
        for (auto& e : c->ev_m) for (auto& ev : e) if (!ev) VSPE_CUDA(cudaEventCreate(&ev));
        bool serial = false;
        for (int attempt = 0; attempt < 8 && !done && !serial; attempt++) {
            unsigned long long h_total[2] = {0, 0}, h_rt[2][2] = {{0, 0}, {0, 0}}, h_err = 0;   // h_rt: {redo tiles, terminators}
            for (int m = 0; m < 2; m++) {
                VSPE_CUDA(cudaEventRecord(c->ev_m[m][0], c->stream));
                VSPE_TRY(fused_launch(c, m, bufs[m], ns[m], 0, guess[m], true));
            }
            for (int m = 0; m < 2; m++) {
                VSPE_CUDA(cudaStreamWaitEvent(c->stream, c->ev_tier[m], 0));
                VSPE_CUDA(cudaMemcpyAsync(h_rt[m], scan_map_total_ptr(c, m, ns[m], bufs[m]), 16, cudaMemcpyDeviceToHost, c->stream));
            }
            VSPE_CUDA(cudaMemcpyAsync(&h_err, c->counters.p + CNT_ERR, 8, cudaMemcpyDeviceToHost, c->stream));
            VSPE_CUDA(cudaStreamSynchronize(c->stream));
            scan_map_account(c);
            for (int m = 0; m < 2; m++) { h_total[m] = h_rt[m][1]; if (!(h_err & ERRF_TILE_FULL)) c->stats.scan_redo_tiles += (uint32_t)h_rt[m][0]; }
            const unsigned long long redo = ERRF_SLOTS_FULL | ERRF_TILE_FULL | ERRF_LISTS_FULL | ERRF_SPILL_FULL;
            if (h_err & redo) {
                const unsigned long long cleared = h_err & ~redo;
                const unsigned long long zero2[2] = {0, 0};
                VSPE_CUDA(cudaMemcpyAsync(c->counters.p + CNT_ERR, &cleared, 8, cudaMemcpyHostToDevice, c->stream));
                // the call started with empty pools (begin_call): a repeated attempt starts there again
                VSPE_CUDA(cudaMemcpyAsync(c->counters.p + CNT_SPILL_CURSOR, &zero2[0], 8, cudaMemcpyHostToDevice, c->stream));
                VSPE_CUDA(cudaMemcpyAsync(c->counters.p + CNT_OVF, &zero2[1], 8, cudaMemcpyHostToDevice, c->stream));
                VSPE_CUDA(cudaStreamSynchronize(c->stream));
            }
            if (h_err & ERRF_TILE_FULL) { serial = true; break; }
            bool again = false;
            if (h_err & (ERRF_LISTS_FULL | ERRF_SPILL_FULL)) { VSPE_TRY(link_grow_overflow(c)); again = true; }
            for (int m = 0; m < 2; m++) {
                const uint64_t n_seq = seq_lines_before(h_total[m]);
                if (n_seq > guess[m]) { guess[m] = n_seq; again = true; }
            }
            if (again) continue;
            for (int m = 0; m < 2; m++) {
                ms[m]->n_slots = seq_lines_before(h_total[m]);
                ms[m]->line_base = h_total[m];
                const bool term = last[m] == '\n' || last[m] == '\r';
                ms[m]->lines = h_total[m] + (term ? 0 : 1);
            }
            done = true;
        }
    }
    if (!done)
        for (int m = 0; m < 2; m++) VSPE_TRY(feed_chunk(c, m, *ms[m], bufs[m], ns[m], true, last[m], false));
    VSPE_TRY(finish_pairs(c, f, r));
    VSPE_CUDA(cudaEventRecord(t1, c->stream));
    VSPE_CUDA(cudaStreamSynchronize(c->stream));
    float ms_total = 0;
    cudaEventElapsedTime(&ms_total, t0, t1);
    c->stats.ms_total += ms_total;
    scan_map_account(c);
    for (int m = 0; m < 2; m++) {
        if (!ns[m]) continue;
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, c->ev_m[m][0], c->ev_m[m][1]);
        cudaEventElapsedTime(&b, c->ev_m[m][1], c->ev_m[m][2]);
        c->stats.ms_scan += a;
        c->stats.ms_map += b;
    }
    c->stats.bytes_fwd += n_fwd;
    c->stats.bytes_rve += n_rve;
    return VSPE_OK;
}

int vspe_count_host(vspe_ctx* c, const uint8_t* fwd, uint64_t n_fwd, const uint8_t* rve, uint64_t n_rve) {
    VSPE_TRY(require_index(c));
    if ((n_fwd && !fwd) || (n_rve && !rve)) { set_error("null input buffer"); return VSPE_ERR_ARG; }
    begin_call(c);
    cudaEvent_t t0 = c->ev[7];
    VSPE_CUDA(cudaEventRecord(t0, c->stream));
    MateStream f, r;
    c->read_len_hint = std::max(seq_len_hint_sampled(fwd, n_fwd), seq_len_hint_sampled(rve, n_rve));
    MateStream ms[2];
    const uint8_t* srcs[2] = {fwd, rve};
    const uint64_t ns[2] = {n_fwd, n_rve};
    VSPE_TRY(stream_mates_host(c, ms, srcs, ns));
    f = ms[0];
    r = ms[1];
    VSPE_TRY(finish_pairs(c, f, r));
    cudaEvent_t t1 = c->ev[1];
    VSPE_CUDA(cudaEventRecord(t1, c->stream));
    VSPE_CUDA(cudaStreamSynchronize(c->stream));
    float ms_total = 0;
    cudaEventElapsedTime(&ms_total, t0, t1);
    c->stats.ms_total += ms_total;
    c->stats.bytes_fwd += n_fwd;
    c->stats.bytes_rve += n_rve;
    return VSPE_OK;
}

static int require_dense(Ctx* c) {
    if (c->sparse.enabled) { set_error("this context counts sparsely (graph too large for N*N matrices or sparse forced): use vspe_sparse_host"); return VSPE_ERR_ARG; }
    return VSPE_OK;
}

int vspe_sparse_host(vspe_ctx* c, uint64_t* n_entries, const uint64_t** keys, const uint64_t** counts) {
    VSPE_TRY(require_index(c));
    if (!c->sparse.enabled) { set_error("this context counts densely: use vspe_matrices_host"); return VSPE_ERR_ARG; }
    Sparse& sp = c->sparse;
    sp.h_keys.resize(sp.n_runs ? sp.n_runs : 1);
    sp.h_counts.resize(sp.n_runs ? sp.n_runs : 1);
    if (sp.n_runs) {
        VSPE_CUDA(cudaMemcpyAsync(sp.h_keys.data(), sp.k[0].p, sp.n_runs * 8, cudaMemcpyDeviceToHost, c->stream));
        VSPE_CUDA(cudaMemcpyAsync(sp.h_counts.data(), sp.v[0].p, sp.n_runs * 8, cudaMemcpyDeviceToHost, c->stream));
        VSPE_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (n_entries) *n_entries = sp.n_runs;
    if (keys) *keys = sp.h_keys.data();
    if (counts) *counts = sp.h_counts.data();
    return VSPE_OK;
}

int vspe_sparse_merge(vspe_ctx* c, const uint64_t* keys, const uint64_t* counts, uint64_t n_entries) {
    VSPE_TRY(require_index(c));
    if (!c->sparse.enabled) { set_error("this context counts densely"); return VSPE_ERR_ARG; }
    return sparse_merge_host(c, keys, counts, n_entries);
}

int vspe_is_sparse(vspe_ctx* c) { return c && c->sparse.enabled ? 1 : 0; }

int vspe_sparse_device(vspe_ctx* c, uint64_t* n_entries, uint64_t** d_keys, uint64_t** d_counts) {
    VSPE_TRY(require_index(c));
    if (!c->sparse.enabled) { set_error("this context counts densely: use vspe_matrices_device"); return VSPE_ERR_ARG; }
    if (n_entries) *n_entries = c->sparse.n_runs;
    if (d_keys) *d_keys = reinterpret_cast<uint64_t*>(c->sparse.k[0].p);
    if (d_counts) *d_counts = reinterpret_cast<uint64_t*>(c->sparse.v[0].p);
    return VSPE_OK;
}

int vspe_sparse_merge_device(vspe_ctx* c, const uint64_t* d_keys, const uint64_t* d_counts, uint64_t n_entries) {
    VSPE_TRY(require_index(c));
    if (!c->sparse.enabled) { set_error("this context counts densely"); return VSPE_ERR_ARG; }
    return sparse_merge_device(c, d_keys, d_counts, n_entries);
}

int vspe_sparse_clear(vspe_ctx* c) {
    VSPE_TRY(require_index(c));
    if (!c->sparse.enabled) { set_error("this context counts densely"); return VSPE_ERR_ARG; }
    c->sparse.n_runs = 0;
    return VSPE_OK;
}

void* vspe_stream(vspe_ctx* c) { return c ? reinterpret_cast<void*>(c->stream) : nullptr; }

int vspe_write_info_sparse(const char* path, const char* const* ids, uint32_t n, const uint64_t* keys, const uint64_t* counts,
                           uint64_t n_entries, int mat) {
    if (!path || (n && !ids) || (n_entries && (!keys || !counts)) || (mat != 0 && mat != 1)) { set_error("bad arguments"); return VSPE_ERR_ARG; }
    FILE* fh = fopen(path, "wb");
    if (!fh) { set_error("cannot create %s: %s", path, strerror(errno)); return VSPE_ERR_IO; }
    const uint64_t NN = (uint64_t)n * n, lo = (uint64_t)mat * NN, hi = lo + NN;
    std::vector<char> buf;
    buf.reserve(1 << 20);
    for (uint64_t e = 0; e < n_entries; e++) {
        if (keys[e] < lo || keys[e] >= hi || counts[e] == 0) continue;
        const uint64_t cell = keys[e] - lo, i = cell / n, j = cell % n;
        char tmp[32];
        char* q = put_u64(tmp, counts[e]);
        const char *a = ids[i], *b = ids[j];
        buf.insert(buf.end(), a, a + strlen(a)); buf.push_back(':');
        buf.insert(buf.end(), b, b + strlen(b)); buf.push_back(':');
        buf.insert(buf.end(), tmp, q); buf.push_back('\n');
        if (buf.size() > (1u << 20) - 256) { fwrite(buf.data(), 1, buf.size(), fh); buf.clear(); }
    }
    if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), fh);
    if (fclose(fh) != 0) { set_error("write to %s failed", path); return VSPE_ERR_IO; }
    return VSPE_OK;
}

int vspe_matrices_device(vspe_ctx* c, uint64_t** d_mats, uint64_t* n_elems) {
    VSPE_TRY(require_index(c));
    VSPE_TRY(require_dense(c));
    if (d_mats) *d_mats = c->mats.p;
    if (n_elems) *n_elems = 2ull * c->index.n_nodes * c->index.n_nodes;
    return VSPE_OK;
}

int vspe_matrices_host(vspe_ctx* c, uint64_t* node_mat, uint64_t* short_mat) {
    VSPE_TRY(require_index(c));
    VSPE_TRY(require_dense(c));
    uint64_t nn = (uint64_t)c->index.n_nodes * c->index.n_nodes;
    if (nn == 0) return VSPE_OK;
    if (node_mat) VSPE_CUDA(cudaMemcpyAsync(node_mat, c->mats.p, nn * 8, cudaMemcpyDeviceToHost, c->stream));
    if (short_mat) VSPE_CUDA(cudaMemcpyAsync(short_mat, c->mats.p + nn, nn * 8, cudaMemcpyDeviceToHost, c->stream));
    VSPE_CUDA(cudaStreamSynchronize(c->stream));
    return VSPE_OK;
}

int vspe_get_stats(vspe_ctx* c, vspe_stats* out) {
    if (!c || !out) { set_error("null argument"); return VSPE_ERR_ARG; }
    VSPE_CUDA(cudaSetDevice(c->device));
    unsigned long long h[CNT_COUNT_];
    VSPE_CUDA(cudaStreamSynchronize(c->stream));           // (vspe_reset zeroes the counters stream-ordered)
    VSPE_CUDA(cudaMemcpy(h, c->counters.p, sizeof(h), cudaMemcpyDeviceToHost));
    if (!c->stats_overridden) {
        c->stats.n_pairs = h[CNT_N];
        c->stats.short_pairs = h[CNT_SHORT];
        c->stats.used_pairs = h[CNT_USED];
    }
    c->stats.n_keys = h[CNT_KEYS];
    c->stats.reads_fast = h[CNT_FAST] - h[CNT_BAILED];
    c->stats.reads_generic = h[CNT_GENERIC];
    c->stats.reads_memo = h[CNT_MEMO_HIT];
    c->stats.kernel_launches = c->launches;
    *out = c->stats;
    return VSPE_OK;
}

int vspe_set_pair_counters(vspe_ctx* c, uint64_t total, uint64_t n, uint64_t shrt, uint64_t used) {
    if (!c) { set_error("null context"); return VSPE_ERR_ARG; }
    c->stats.total_pairs = total;
    c->stats.n_pairs = n;
    c->stats.short_pairs = shrt;
    c->stats.used_pairs = used;
    c->stats_overridden = true;
    return VSPE_OK;
}

int vspe_split_records(vspe_ctx* c, const uint8_t* fq, uint64_t n_bytes, uint64_t* n_lines, uint64_t* n_records,
                       const uint64_t** seq_start, const uint32_t** seq_len) {
    if (!c) { set_error("null context"); return VSPE_ERR_ARG; }
    VSPE_CUDA(cudaSetDevice(c->device));
    uint64_t lines = 0, recs = 0;
    c->h_seq_start.clear();
    c->h_seq_len.clear();
    if (n_bytes) {
        VSPE_TRY(c->dev_in[0].reserve(n_bytes + 64));
        VSPE_CUDA(cudaMemcpyAsync(c->dev_in[0].p, fq, n_bytes, cudaMemcpyHostToDevice, c->stream));
        uint64_t n_terms = 0, n_seq = 0;
        Records& rec = c->mate[0].rec;
        if (c->opt_scan_two_pass) {
            VSPE_TRY(scan_count_lines(c, c->dev_in[0].p, n_bytes, &n_terms));
            n_seq = seq_lines_before(n_terms);
            VSPE_TRY(rec.seq_start.reserve(n_seq + 2));
            VSPE_TRY(rec.seq_end.reserve(n_seq + 2));
            VSPE_TRY(scan_index_records(c, c->dev_in[0].p, n_bytes, 0, 0, n_seq, rec.seq_start.p, rec.seq_end.p));
        } else {
            bool overflow = false;
            VSPE_TRY(rec.seq_start.reserve(n_bytes / 48 + 1026));
            VSPE_TRY(rec.seq_end.reserve(n_bytes / 48 + 1026));
            VSPE_TRY(scan_records_single_pass(c, c->dev_in[0].p, n_bytes, 0, 0, rec.seq_start.cap - 2, rec.seq_start.p, rec.seq_end.p, &n_terms, &overflow));
            n_seq = seq_lines_before(n_terms);
            if (overflow) {
                VSPE_TRY(rec.seq_start.reserve(n_seq + 2));
                VSPE_TRY(rec.seq_end.reserve(n_seq + 2));
                VSPE_TRY(scan_records_single_pass(c, c->dev_in[0].p, n_bytes, 0, 0, n_seq, rec.seq_start.p, rec.seq_end.p, &n_terms, &overflow));
            }
        }
        bool term = fq[n_bytes - 1] == '\n' || fq[n_bytes - 1] == '\r';
        lines = n_terms + (term ? 0 : 1);
        recs = lines / 4;
        std::vector<uint64_t> e(n_seq);
        c->h_seq_start.resize(n_seq);
        if (n_seq) {
            VSPE_CUDA(cudaMemcpyAsync(c->h_seq_start.data(), rec.seq_start.p, n_seq * 8, cudaMemcpyDeviceToHost, c->stream));
            VSPE_CUDA(cudaMemcpyAsync(e.data(), rec.seq_end.p, n_seq * 8, cudaMemcpyDeviceToHost, c->stream));
        }
        VSPE_CUDA(cudaStreamSynchronize(c->stream));
        c->h_seq_start.resize(recs);
        c->h_seq_len.resize(recs);
        for (uint64_t r = 0; r < recs; r++) c->h_seq_len[r] = (uint32_t)(e[r] - c->h_seq_start[r]);
        VSPE_TRY(check_kernel_errors(c));
    }
    if (n_lines) *n_lines = lines;
    if (n_records) *n_records = recs;
    if (seq_start) *seq_start = c->h_seq_start.data();
    if (seq_len) *seq_len = c->h_seq_len.data();
    return VSPE_OK;
}

int vspe_map_reads(vspe_ctx* c, const uint8_t* fq, uint64_t n_bytes, uint64_t* n_reads, const uint64_t** offsets,
                   const uint32_t** nodes, const uint8_t** status) {
    VSPE_TRY(require_index(c));
    begin_call(c);
    MateStream both[2];
    c->read_len_hint = seq_len_hint_sampled(fq, n_bytes);
    {
        const uint8_t* srcs[2] = {fq, nullptr};
        const uint64_t ns[2] = {n_bytes, 0};
        VSPE_TRY(stream_mates_host(c, both, srcs, ns));
    }
    const MateStream& ms = both[0];
    VSPE_TRY(check_kernel_errors(c));
    uint64_t recs = ms.lines / 4;
    std::vector<ReadSlot> slots(recs);
    if (recs) {
        VSPE_TRY(c->mate[0].slots.reserve(recs + 1));
        VSPE_TRY(export_slots(c, c->mate[0].handles.p, recs, c->mate[0].slots.p));
        VSPE_CUDA(cudaStreamSynchronize(c->stream));
        VSPE_CUDA(cudaMemcpy(slots.data(), c->mate[0].slots.p, recs * sizeof(ReadSlot), cudaMemcpyDeviceToHost));
    }
    unsigned long long spill_n = 0;
    VSPE_CUDA(cudaMemcpy(&spill_n, c->counters.p + CNT_SPILL_CURSOR, 8, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> spill(spill_n);
    if (spill_n) VSPE_CUDA(cudaMemcpy(spill.data(), c->spill.p, spill_n * 4, cudaMemcpyDeviceToHost));
    c->h_offsets.assign(recs + 1, 0);
    c->h_nodes.clear();
    c->h_status.assign(recs, 0);
    for (uint64_t r = 0; r < recs; r++) {
        c->h_offsets[r] = c->h_nodes.size();
        uint32_t st = slots[r].hdr & 0xFF, n = slots[r].hdr >> 8;
        c->h_status[r] = (uint8_t)st;
        if (st != ST_OK) continue;
        const uint32_t* ids = n <= (uint32_t)SLOT_IDS ? slots[r].ids : spill.data() + slots[r].ids[0];
        c->h_nodes.insert(c->h_nodes.end(), ids, ids + n);
    }
    c->h_offsets[recs] = c->h_nodes.size();
    if (c->h_nodes.empty()) c->h_nodes.push_back(0);
    if (n_reads) *n_reads = recs;
    if (offsets) *offsets = c->h_offsets.data();
    if (nodes) *nodes = c->h_nodes.data();
    if (status) *status = c->h_status.data();
    return VSPE_OK;
}

int vspe_write_info(const char* path, const char* const* ids, uint32_t n, const uint64_t* mat) {
    if (!path || (n && (!ids || !mat))) { set_error("bad arguments"); return VSPE_ERR_ARG; }
    int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) { set_error("cannot create %s: %s", path, strerror(errno)); return VSPE_ERR_IO; }
    std::vector<size_t> idlen(n);
    size_t maxid = 0;
    for (uint32_t i = 0; i < n; i++) { idlen[i] = strlen(ids[i]); maxid = std::max(maxid, idlen[i]); }
    unsigned nt = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (n < 64) nt = 1;
    // rows are formatted in parallel into per-thread buffers (bounded slabs), written in order
    const uint32_t SLAB = std::max<uint32_t>(nt, 64);           // rows per round
    std::vector<std::vector<char>> buf(nt);
    int rc = VSPE_OK;
    for (uint32_t r0 = 0; r0 < n && rc == VSPE_OK; r0 += SLAB) {
        uint32_t r1 = std::min(n, r0 + SLAB);
        uint32_t per = (r1 - r0 + nt - 1) / nt;
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++) {
            uint32_t a = r0 + t * per, b = std::min(r1, a + per);
            buf[t].clear();
            if (a >= b) continue;
            th.emplace_back([&, t, a, b] {
                std::vector<char>& out = buf[t];
                out.resize((size_t)(b - a) * n * (2 * maxid + 24));
                char* p = out.data();
                for (uint32_t i = a; i < b; i++) {
                    const uint64_t* row = mat + (size_t)i * n;
                    for (uint32_t j = 0; j < n; j++) {
                        memcpy(p, ids[i], idlen[i]); p += idlen[i]; *p++ = ':';
                        memcpy(p, ids[j], idlen[j]); p += idlen[j]; *p++ = ':';
                        p = put_u64(p, row[j]);
                        *p++ = '\n';
                    }
                }
                out.resize(p - out.data());
            });
        }
        for (auto& t : th) t.join();
        for (unsigned t = 0; t < nt && rc == VSPE_OK; t++) {
            const char* p = buf[t].data();
            size_t left = buf[t].size();
            while (left) {
                ssize_t w = ::write(fd, p, left);
                if (w < 0) { if (errno == EINTR) continue; set_error("write to %s failed: %s", path, strerror(errno)); rc = VSPE_ERR_IO; break; }
                p += w; left -= (size_t)w;
            }
        }
    }
    if (close(fd) != 0 && rc == VSPE_OK) { set_error("close of %s failed", path); rc = VSPE_ERR_IO; }
    return rc;
}

// Both read files are gzip streams: count them chunk by chunk as they inflate (see ChunkProducer).
static int count_gzip_streams(vspe_ctx* c, const InputFile& f, const InputFile& r) {
    VSPE_TRY(require_index(c));
    begin_call(c);
    cudaEvent_t t0 = c->ev[7], t1 = c->ev[1];
    VSPE_CUDA(cudaEventRecord(t0, c->stream));
    const uint64_t chunk = (uint64_t)std::max<int64_t>(1, c->opt_chunk_mb) << 20;
    ChunkProducer prod[2];
    VSPE_TRY(prod[0].start(f, chunk));
    VSPE_TRY(prod[1].start(r, chunk));
    for (int m = 0; m < 2; m++) VSPE_TRY(c->dev_in[m].reserve(chunk + 64));
    MateStream ms[2];
    uint32_t hint = 0;
    for (int m = 0; m < 2; m++) {                               // the first chunks size the packed rows
        prod[m].wait_filled(0);
        if (prod[m].rc != VSPE_OK) { set_error("%s", prod[m].err.c_str()); return prod[m].rc; }
        hint = std::max(hint, seq_len_hint(prod[m].buf[0], std::min<uint64_t>(prod[m].len[0], 16384)));
    }
    c->read_len_hint = hint;
    int b[2] = {0, 0};
    bool done[2] = {false, false};
    uint64_t bytes[2] = {0, 0};
    while (!done[0] || !done[1]) {
        for (int m = 0; m < 2; m++) {
            if (done[m]) continue;
            ChunkProducer& pr = prod[m];
            pr.wait_filled(b[m]);
            if (pr.rc != VSPE_OK) { set_error("%s", pr.err.c_str()); return pr.rc; }
            const uint64_t n = pr.len[b[m]];
            const bool last = pr.last[b[m]];
            if (n) VSPE_CUDA(cudaMemcpyAsync(c->dev_in[m].p, pr.buf[b[m]], n, cudaMemcpyHostToDevice, c->stream));
            VSPE_TRY(feed_chunk(c, m, ms[m], n ? c->dev_in[m].p : nullptr, n, last, n ? pr.buf[b[m]][n - 1] : -1));
            bytes[m] += n;
            pr.release(b[m]);
            b[m] ^= 1;
            done[m] = last;
        }
    }
    VSPE_TRY(finish_pairs(c, ms[0], ms[1]));
    VSPE_CUDA(cudaEventRecord(t1, c->stream));
    VSPE_CUDA(cudaStreamSynchronize(c->stream));
    float ms_total = 0;
    cudaEventElapsedTime(&ms_total, t0, t1);
    c->stats.ms_total += ms_total;
    c->stats.bytes_fwd += bytes[0];
    c->stats.bytes_rve += bytes[1];
    return VSPE_OK;
}

static int rm_rf(const std::string& dir) {
    // PE_Inference.py:95 `rm -rf DIR`: the script owns its output directory
    std::string cmd = "rm -rf -- '";
    for (char ch : dir) { if (ch == '\'') cmd += "'\\''"; else cmd += ch; }
    cmd += "'";
    return system(cmd.c_str());
}

int vspe_run(const char* gfa_path, const char* fwd_path, const char* rve_path, int kmer_size, const char* out_dir,
             int n_gpus, vspe_stats* stats) {
    if (!gfa_path || !fwd_path || !rve_path || !out_dir || kmer_size < 1) { set_error("bad arguments"); return VSPE_ERR_ARG; }
    if (n_gpus < 1 || n_gpus > 16) { set_error("n_gpus must be in 1..16 (got %d)", n_gpus); return VSPE_ERR_ARG; }
    std::string dir(out_dir);
    if (!dir.empty() && dir.back() == '/') dir.pop_back();      // :93-94
    if (dir.empty()) { set_error("empty output directory"); return VSPE_ERR_ARG; }
    if (rm_rf(dir) != 0 || mkdir(dir.c_str(), 0777) != 0) {
        // os.makedirs creates parents too: one mkdir(2) per path component (no shell involved)
        bool ok = true;
        for (size_t i = 1; i <= dir.size() && ok; i++) {
            if (i != dir.size() && dir[i] != '/') continue;
            const std::string part = dir.substr(0, i);
            if (mkdir(part.c_str(), 0777) != 0 && errno != EEXIST) ok = false;
        }
        struct stat sb;
        if (!ok || stat(dir.c_str(), &sb) != 0 || !S_ISDIR(sb.st_mode)) { set_error("cannot create output directory %s", dir.c_str()); return VSPE_ERR_IO; }
    }
    InputFile g, f, r;
    VSPE_TRY(g.open_ro(gfa_path));
    GfaNodes nodes;
    VSPE_TRY(parse_gfa_bytes(g.p, g.n, nodes));
    {   // the two read files are opened (and, if gzipped, inflated) side by side
        int rc_r = VSPE_OK;
        std::string err_r;
        // (one GPU: gzip read files stay compressed here and are streamed; shards of several GPUs need the bytes)
        const bool lazy = n_gpus == 1;
        std::thread tr([&] { rc_r = r.open_ro(rve_path, !lazy); if (rc_r != VSPE_OK) err_r = get_error(); });
        int rc_f = f.open_ro(fwd_path, !lazy);
        tr.join();
        if (rc_f == VSPE_OK && rc_r == VSPE_OK && lazy && f.gz != r.gz) {     // only one of them is gzip: inflate it whole
            if (f.gz) rc_f = f.inflate_all(); else { rc_r = r.inflate_all(); if (rc_r != VSPE_OK) err_r = get_error(); }
        }
        if (rc_f != VSPE_OK) return rc_f;
        if (rc_r != VSPE_OK) { set_error("%s", err_r.c_str()); return rc_r; }
    }
    const bool stream_gz = n_gpus == 1 && f.gz && r.gz && f.p == nullptr && r.p == nullptr;
    uint32_t N = (uint32_t)nodes.ids.size();
    std::vector<uint64_t> nm, sm, sk, sc;
    bool sparse_out = !dense_possible(N) || (getenv("VSPE_SPARSE") && atoi(getenv("VSPE_SPARSE")) != 0);
    if (n_gpus > 1) {
        VSPE_TRY(run_multi_gpu(nodes.seqs.data(), nodes.off.data(), N, (uint32_t)kmer_size + 1, f.p, f.n, r.p, r.n, n_gpus, nm, sm,
                               sparse_out ? &sk : nullptr, sparse_out ? &sc : nullptr, stats));
    } else {
        vspe_ctx* c = nullptr;
        VSPE_TRY(vspe_create(0, &c));
        if (sparse_out) c->opt_sparse = 1;
        int rc = vspe_index_build(c, nodes.seqs.data(), nodes.off.data(), N, (uint32_t)kmer_size + 1);
        if (rc == VSPE_OK) rc = stream_gz ? count_gzip_streams(c, f, r) : vspe_count_host(c, f.p, f.n, r.p, r.n);
        if (rc == VSPE_OK && !sparse_out) {
            nm.assign((size_t)N * N, 0);
            sm.assign((size_t)N * N, 0);
            rc = vspe_matrices_host(c, nm.data(), sm.data());
        }
        if (rc == VSPE_OK && sparse_out) {
            uint64_t ne = 0;
            const uint64_t *pk = nullptr, *pc = nullptr;
            rc = vspe_sparse_host(c, &ne, &pk, &pc);
            if (rc == VSPE_OK) { sk.assign(pk, pk + ne); sc.assign(pc, pc + ne); }
        }
        if (rc == VSPE_OK && stats) rc = vspe_get_stats(c, stats);
        vspe_destroy(c);
        if (rc != VSPE_OK) return rc;
    }
    if (sparse_out) {
        // N*N lines cannot be written for such graphs; only the non-zero lines are.  The consumer
        // (reference utils/VStrains_IO.py:598-612) zero-initialises every key, so the parsed result
        // is the same dict a dense file would give.
        std::vector<const char*> idq(N);
        for (uint32_t i = 0; i < N; i++) idq[i] = nodes.ids[i].c_str();
        VSPE_TRY(vspe_write_info_sparse((dir + "/pe_info").c_str(), idq.data(), N, sk.data(), sc.data(), sk.size(), 0));
        VSPE_TRY(vspe_write_info_sparse((dir + "/st_info").c_str(), idq.data(), N, sk.data(), sc.data(), sk.size(), 1));
        return VSPE_OK;
    }
    std::vector<const char*> idp(N);
    for (uint32_t i = 0; i < N; i++) idp[i] = nodes.ids[i].c_str();
    VSPE_TRY(vspe_write_info((dir + "/pe_info").c_str(), idp.data(), N, nm.data()));
    VSPE_TRY(vspe_write_info((dir + "/st_info").c_str(), idp.data(), N, sm.data()));
    return VSPE_OK;
}

int vspe_read_input(const char* path, uint8_t** data, uint64_t* n_bytes) {
    if (!path || !data || !n_bytes) { set_error("bad arguments"); return VSPE_ERR_ARG; }
    *data = nullptr;
    *n_bytes = 0;
    InputFile in;
    VSPE_TRY(in.open_ro(path));
    uint8_t* out = static_cast<uint8_t*>(malloc(in.n ? in.n : 1));
    if (!out) { set_error("out of memory reading %s", path); return VSPE_ERR_IO; }
    if (in.n) memcpy(out, in.p, in.n);
    *data = out;
    *n_bytes = in.n;
    return VSPE_OK;
}
void vspe_free_input(uint8_t* data) { free(data); }

void* vspe_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { set_error("cudaMallocHost(%zu) failed", bytes); cudaGetLastError(); return nullptr; }
    return p;
}
void vspe_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int vspe_set_option(vspe_ctx* c, const char* name, int64_t value) {
    if (!c || !name) { set_error("bad arguments"); return VSPE_ERR_ARG; }
    if (!strcmp(name, "force_generic")) c->opt_force_generic = value;
    else if (!strcmp(name, "chunk_mb")) c->opt_chunk_mb = value;
    else if (!strcmp(name, "scan_two_pass")) c->opt_scan_two_pass = value;
    else if (!strcmp(name, "scan_mode")) c->opt_scan_mode = value;
    else if (!strcmp(name, "sparse")) {
        // effective for the next index build; on a built index only switching a dense-capable graph is allowed
        c->opt_sparse = value;
        if (c->index.built && dense_possible(c->index.n_nodes)) {
            c->sparse.enabled = value != 0;
            c->sparse.n_runs = 0;
            if (!c->sparse.enabled) {
                uint64_t nn = 2ull * c->index.n_nodes * c->index.n_nodes;
                VSPE_TRY(c->mats.reserve(nn ? nn : 1));
                VSPE_CUDA(cudaMemset(c->mats.p, 0, (nn ? nn : 1) * 8));
            }
        }
    }
    else if (!strcmp(name, "tier_overlap")) c->opt_tier_overlap = value;
    else if (!strcmp(name, "memo")) c->opt_memo = value;
    else if (!strcmp(name, "link_split")) c->opt_link_split = value < 256 ? 256 : value;
    else if (!strcmp(name, "pair_cap_log2")) c->opt_pair_cap_log2 = value;
    else if (!strcmp(name, "subst")) { c->opt_subst = value; if (!value) c->index.has_subst = false; }
    else if (!strcmp(name, "dbg_counters")) {
        // profiling aid: copy the device counters (enum Counter order, CNT_COUNT_ words) to the host pointer `value`
        if (value) cudaMemcpy(reinterpret_cast<void*>(value), c->counters.p, CNT_COUNT_ * 8, cudaMemcpyDeviceToHost);
    }
    else { set_error("unknown option %s", name); return VSPE_ERR_ARG; }
    return VSPE_OK;
}

}  // extern "C"
