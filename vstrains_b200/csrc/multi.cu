// multi.cu -- multi-GPU driver of the whole path inside ONE process (the reference's caller
// spawns exactly one process, reference utils/VStrains_SPAdes.py:118-132).
//
// Read pairs shard naturally (PE_Inference.py:154-188 has no cross-pair state): both FASTQ files
// are cut at the SAME record numbers, every device gets the replicated index and one contiguous
// record range, and the per-device [node_mat | short_mat] are summed with ONE ncclAllReduce
// (uint64 sum) over NVLink.  Integer addition makes the result independent of the device count.
// NCCL is resolved with dlopen so that processes which already carry another NCCL (PyTorch)
// never see two copies unless they ask for this entry point.
#include <dlfcn.h>

#include <thread>

#include "ctx.cuh"

namespace vspe {

typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
static constexpr int kNcclUint64 = 5, kNcclSum = 0;     // ncclDataType_t / ncclRedOp_t values (nccl.h)

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    int load() {
        if (h) return VSPE_OK;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (h) break; }
        if (!h) { set_error("cannot load NCCL: %s", dlerror()); return VSPE_ERR_NCCL; }
#define VSPE_SYM(field, name) *(void**)(&field) = dlsym(h, name); if (!field) { set_error("NCCL symbol %s missing", name); return VSPE_ERR_NCCL; }
        VSPE_SYM(CommInitAll, "ncclCommInitAll")
        VSPE_SYM(CommDestroy, "ncclCommDestroy")
        VSPE_SYM(AllReduce, "ncclAllReduce")
        VSPE_SYM(AllGather, "ncclAllGather")
        VSPE_SYM(GroupStart, "ncclGroupStart")
        VSPE_SYM(GroupEnd, "ncclGroupEnd")
        VSPE_SYM(GetErrorString, "ncclGetErrorString")
#undef VSPE_SYM
        return VSPE_OK;
    }
};

// Universal-newline terminator test at byte i of p[0..n)  ("\r\n" counts once, at the '\n').
static inline bool is_term(const uint8_t* p, uint64_t n, uint64_t i) {
    const uint8_t c = p[i];
    return c == '\n' || (c == '\r' && !(i + 1 < n && p[i + 1] == '\n'));
}

// Lines of a buffer and the byte offsets one past the terminator that ends line 4*rec - 1 for every rec
// in `recs` (ascending).  The buffer is cut into slices whose terminators are counted in parallel; a cut
// then only rescans the one slice that holds it.
struct LineIndex {
    const uint8_t* p = nullptr;
    uint64_t n = 0, per = 0;
    std::vector<uint64_t> before;                  // terminators before slice t
    uint64_t lines = 0;
    void build(const uint8_t* buf, uint64_t len) {
        p = buf;
        n = len;
        unsigned nt = std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
        if (n < (1u << 20)) nt = 1;
        per = (n + nt - 1) / nt;
        std::vector<uint64_t> part(nt, 0);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++) {
            const uint64_t a = t * per, b = std::min(n, a + per);
            if (a >= b) break;
            th.emplace_back([=, &part] {
                uint64_t c = 0;
                for (uint64_t i = a; i < b; i++) c += is_term(p, n, i);
                part[t] = c;
            });
        }
        for (auto& t : th) t.join();
        before.assign(nt + 1, 0);
        for (unsigned t = 0; t < nt; t++) before[t + 1] = before[t] + part[t];
        lines = before[nt];
        if (n && !(p[n - 1] == '\n' || p[n - 1] == '\r')) lines++;
    }
    // byte offset one past the terminator number `line` (1-based); n if there are fewer
    uint64_t end_of_line(uint64_t line) const {
        if (line == 0) return 0;
        if (line > before.back()) return n;
        size_t t = 0;
        while (before[t + 1] < line) t++;           // slice holding terminator number `line`
        uint64_t seen = before[t];
        for (uint64_t i = t * per; i < n; i++) {
            if (!is_term(p, n, i)) continue;
            if (++seen == line) return i + 1;
        }
        return n;
    }
};

extern "C" int vspe_count_host(vspe_ctx* c, const uint8_t* fwd, uint64_t n_fwd, const uint8_t* rve, uint64_t n_rve);
extern "C" int vspe_sparse_host(vspe_ctx* c, uint64_t* n_entries, const uint64_t** keys, const uint64_t** counts);

int run_multi_gpu(const uint8_t* seqs, const uint64_t* seq_off, uint32_t n_nodes, uint32_t split_len,
                  const uint8_t* fwd, uint64_t n_fwd, const uint8_t* rve, uint64_t n_rve, int n_gpus,
                  std::vector<uint64_t>& node_mat, std::vector<uint64_t>& short_mat,
                  std::vector<uint64_t>* sparse_keys, std::vector<uint64_t>* sparse_counts, vspe_stats* stats) {
    const bool sparse = sparse_keys != nullptr;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < n_gpus) {
        set_error("asked for %d GPUs but %d are visible", n_gpus, have);
        return VSPE_ERR_CUDA;
    }
    static NcclApi nccl;
    VSPE_TRY(nccl.load());
    // record-aligned shards (pairing is by record index: PE_Inference.py:154-159); both files are
    // indexed side by side, every cut then rescans one slice only
    LineIndex li_f, li_r;
    {
        std::thread tr([&] { li_r.build(rve, n_rve); });
        li_f.build(fwd, n_fwd);
        tr.join();
    }
    const uint64_t total = std::min(li_f.lines / 4, li_r.lines / 4);
    std::vector<uint64_t> cf(n_gpus + 1), cr(n_gpus + 1);
    for (int g = 0; g <= n_gpus; g++) {
        const uint64_t rec = total * g / n_gpus;
        cf[g] = g == n_gpus ? n_fwd : li_f.end_of_line(4 * rec);
        cr[g] = g == n_gpus ? n_rve : li_r.end_of_line(4 * rec);
    }
    // the communicator is created while the devices count (it takes longer than a small run)
    const uint64_t nn = (uint64_t)n_nodes * n_nodes;
    std::vector<ncclComm_t> comms(n_gpus, nullptr);
    ncclResult_t comm_rc = 0;
    std::thread comm_thread([&] {
        std::vector<int> devs(n_gpus);
        for (int g = 0; g < n_gpus; g++) devs[g] = g;
        comm_rc = nccl.CommInitAll(comms.data(), n_gpus, devs.data());
    });
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());

    std::vector<vspe_ctx*> ctx(n_gpus, nullptr);
    std::vector<int> rc(n_gpus, VSPE_OK);
    std::vector<std::string> err(n_gpus);
    std::vector<std::thread> th;
    for (int g = 0; g < n_gpus; g++) {
        th.emplace_back([&, g] {
            int r = vspe_create(g, &ctx[g]);
            if (r == VSPE_OK && sparse) ctx[g]->opt_sparse = 1;
            if (r == VSPE_OK) ctx[g]->opt_stage_threads = std::max(1u, std::min(8u, hw / (unsigned)n_gpus));
            if (r == VSPE_OK) r = vspe_index_build(ctx[g], seqs, seq_off, n_nodes, split_len);
            if (r == VSPE_OK) r = vspe_count_host(ctx[g], fwd + cf[g], cf[g + 1] - cf[g], rve + cr[g], cr[g + 1] - cr[g]);
            rc[g] = r;
            if (r != VSPE_OK) err[g] = get_error();
        });
    }
    for (auto& t : th) t.join();
    int result = VSPE_OK;
    for (int g = 0; g < n_gpus; g++)
        if (rc[g] != VSPE_OK) { set_error("GPU %d: %s", g, err[g].c_str()); result = rc[g]; break; }

    comm_thread.join();
    if (result == VSPE_OK && comm_rc != 0) { set_error("ncclCommInitAll: %s", nccl.GetErrorString(comm_rc)); result = VSPE_ERR_NCCL; }
    if (result == VSPE_OK && sparse) {
        // sparse runs: ONE exchange step over NVLink -- all-gather of the run counts, all-gather of the
        // (padded) runs -- then device 0 merges the other devices' runs locally (sort + run-length reduce)
        std::vector<DevBuf<unsigned long long>> cnt(n_gpus), gath_n(n_gpus), send(n_gpus), gath(n_gpus);
        for (int g = 0; g < n_gpus && result == VSPE_OK; g++) {
            cudaSetDevice(g);
            if (cnt[g].reserve(1) != VSPE_OK || gath_n[g].reserve(n_gpus) != VSPE_OK) { result = VSPE_ERR_CUDA; break; }
            const unsigned long long nr = ctx[g]->sparse.n_runs;
            cudaMemcpyAsync(cnt[g].p, &nr, 8, cudaMemcpyHostToDevice, ctx[g]->stream);
            cudaStreamSynchronize(ctx[g]->stream);
        }
        ncclResult_t nr = 0;
        if (result == VSPE_OK) {
            nccl.GroupStart();
            for (int g = 0; g < n_gpus; g++) {
                cudaSetDevice(g);
                nr = nccl.AllGather(cnt[g].p, gath_n[g].p, 1, kNcclUint64, comms[g], ctx[g]->stream);
                if (nr != 0) break;
            }
            const ncclResult_t ge = nccl.GroupEnd();
            if (nr != 0 || ge != 0) { set_error("ncclAllGather: %s", nccl.GetErrorString(nr ? nr : ge)); result = VSPE_ERR_NCCL; }
        }
        std::vector<unsigned long long> sizes(n_gpus, 0);
        if (result == VSPE_OK) {
            cudaSetDevice(0);
            cudaMemcpyAsync(sizes.data(), gath_n[0].p, 8ull * n_gpus, cudaMemcpyDeviceToHost, ctx[0]->stream);
            cudaStreamSynchronize(ctx[0]->stream);
            unsigned long long cap = 1;
            for (auto v : sizes) cap = std::max(cap, v);
            for (int g = 0; g < n_gpus && result == VSPE_OK; g++) {
                cudaSetDevice(g);
                if (send[g].reserve(2 * cap) != VSPE_OK || gath[g].reserve(2 * cap * n_gpus) != VSPE_OK) { result = VSPE_ERR_CUDA; break; }
                cudaMemsetAsync(send[g].p, 0, 16 * cap, ctx[g]->stream);
                if (ctx[g]->sparse.n_runs) {
                    cudaMemcpyAsync(send[g].p, ctx[g]->sparse.k[0].p, 8 * ctx[g]->sparse.n_runs, cudaMemcpyDeviceToDevice, ctx[g]->stream);
                    cudaMemcpyAsync(send[g].p + cap, ctx[g]->sparse.v[0].p, 8 * ctx[g]->sparse.n_runs, cudaMemcpyDeviceToDevice, ctx[g]->stream);
                }
            }
            if (result == VSPE_OK) {
                nccl.GroupStart();
                for (int g = 0; g < n_gpus; g++) {
                    cudaSetDevice(g);
                    nr = nccl.AllGather(send[g].p, gath[g].p, 2 * cap, kNcclUint64, comms[g], ctx[g]->stream);
                    if (nr != 0) break;
                }
                const ncclResult_t ge = nccl.GroupEnd();
                if (nr != 0 || ge != 0) { set_error("ncclAllGather: %s", nccl.GetErrorString(nr ? nr : ge)); result = VSPE_ERR_NCCL; }
                for (int g = 0; g < n_gpus; g++) { cudaSetDevice(g); cudaStreamSynchronize(ctx[g]->stream); }
            }
            cudaSetDevice(0);
            for (int g = 1; g < n_gpus && result == VSPE_OK; g++) {
                const uint64_t* base = reinterpret_cast<const uint64_t*>(gath[0].p) + 2 * cap * g;
                result = sparse_merge_device(ctx[0], base, base + cap, sizes[g]);
            }
        }
        if (result == VSPE_OK) {
            uint64_t ne = 0;
            const uint64_t *pk = nullptr, *pc = nullptr;
            cudaSetDevice(0);
            result = vspe_sparse_host(ctx[0], &ne, &pk, &pc);
            if (result == VSPE_OK) { sparse_keys->assign(pk, pk + ne); sparse_counts->assign(pc, pc + ne); }
        }
        for (int g = 0; g < n_gpus; g++) {          // the scratch buffers belong to their devices
            cudaSetDevice(g);
            cnt[g].release(); gath_n[g].release(); send[g].release(); gath[g].release();
        }
    }
    if (result == VSPE_OK && nn && !sparse) {
        ncclResult_t nr = 0;
        nccl.GroupStart();
        for (int g = 0; g < n_gpus; g++) {
            cudaSetDevice(g);
            nr = nccl.AllReduce(ctx[g]->mats.p, ctx[g]->mats.p, 2 * nn, kNcclUint64, kNcclSum, comms[g], ctx[g]->stream);
            if (nr != 0) break;
        }
        const ncclResult_t ge = nccl.GroupEnd();
        if (nr != 0 || ge != 0) { set_error("ncclAllReduce: %s", nccl.GetErrorString(nr ? nr : ge)); result = VSPE_ERR_NCCL; }
        for (int g = 0; g < n_gpus; g++) { cudaSetDevice(g); cudaStreamSynchronize(ctx[g]->stream); }
    }
    if (comm_rc == 0) for (int g = 0; g < n_gpus; g++) if (comms[g]) nccl.CommDestroy(comms[g]);
    if (result == VSPE_OK && !sparse) {
        node_mat.assign(nn, 0);
        short_mat.assign(nn, 0);
        result = vspe_matrices_host(ctx[0], node_mat.data(), short_mat.data());
    }
    if (result == VSPE_OK && stats) {
        vspe_stats sum = {};
        for (int g = 0; g < n_gpus; g++) {
            vspe_stats s;
            if (vspe_get_stats(ctx[g], &s) != VSPE_OK) continue;
            if (g == 0) sum = s;
            else {
                sum.total_pairs += s.total_pairs; sum.n_pairs += s.n_pairs; sum.short_pairs += s.short_pairs;
                sum.used_pairs += s.used_pairs; sum.bytes_fwd += s.bytes_fwd; sum.bytes_rve += s.bytes_rve;
                sum.reads_fast += s.reads_fast; sum.reads_generic += s.reads_generic; sum.n_keys += s.n_keys;
                sum.kernel_launches += s.kernel_launches;
                sum.ms_total = std::max(sum.ms_total, s.ms_total);
            }
        }
        *stats = sum;
    }
    for (int g = 0; g < n_gpus; g++) if (ctx[g]) vspe_destroy(ctx[g]);
    return result;
}

}  // namespace vspe
