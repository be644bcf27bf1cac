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
        VSPE_SYM(GroupStart, "ncclGroupStart")
        VSPE_SYM(GroupEnd, "ncclGroupEnd")
        VSPE_SYM(GetErrorString, "ncclGetErrorString")
#undef VSPE_SYM
        return VSPE_OK;
    }
};

// byte offsets one past the terminator that ends line `4*rec - 1` for rec in cuts (host scan)
static void record_cuts(const uint8_t* p, uint64_t n, const std::vector<uint64_t>& recs, std::vector<uint64_t>& out) {
    out.assign(recs.size(), n);
    uint64_t line = 0;
    size_t k = 0;
    while (k < recs.size() && recs[k] == 0) out[k++] = 0;
    for (uint64_t i = 0; i < n && k < recs.size(); i++) {
        uint8_t c = p[i];
        bool term = c == '\n' || (c == '\r' && !(i + 1 < n && p[i + 1] == '\n'));
        if (!term) continue;
        line++;
        while (k < recs.size() && line == 4 * recs[k]) out[k++] = i + 1;
    }
}

static uint64_t count_lines_host(const uint8_t* p, uint64_t n) {
    // parallel terminator count; a "\r\n" split across two slices is counted once because the
    // '\r' looks at its successor byte wherever that lives
    unsigned nt = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    std::vector<uint64_t> part(nt, 0);
    std::vector<std::thread> th;
    uint64_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        uint64_t a = t * per, b = std::min(n, a + per);
        if (a >= b) break;
        th.emplace_back([=, &part] {
            uint64_t c = 0;
            for (uint64_t i = a; i < b; i++) {
                uint8_t ch = p[i];
                c += ch == '\n' || (ch == '\r' && !(i + 1 < n && p[i + 1] == '\n'));
            }
            part[t] = c;
        });
    }
    for (auto& t : th) t.join();
    uint64_t lines = 0;
    for (auto v : part) lines += v;
    if (n && !(p[n - 1] == '\n' || p[n - 1] == '\r')) lines++;
    return lines;
}

extern "C" int vspe_count_host(vspe_ctx* c, const uint8_t* fwd, uint64_t n_fwd, const uint8_t* rve, uint64_t n_rve);
extern "C" int vspe_sparse_host(vspe_ctx* c, uint64_t* n_entries, const uint64_t** keys, const uint64_t** counts);
extern "C" int vspe_sparse_merge(vspe_ctx* c, const uint64_t* keys, const uint64_t* counts, uint64_t n_entries);

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
    // record-aligned shards (pairing is by record index: PE_Inference.py:154-159)
    const uint64_t total = std::min(count_lines_host(fwd, n_fwd) / 4, count_lines_host(rve, n_rve) / 4);
    std::vector<uint64_t> recs(n_gpus + 1), cf, cr;
    for (int g = 0; g <= n_gpus; g++) recs[g] = total * g / n_gpus;
    record_cuts(fwd, n_fwd, recs, cf);
    record_cuts(rve, n_rve, recs, cr);

    std::vector<vspe_ctx*> ctx(n_gpus, nullptr);
    std::vector<int> rc(n_gpus, VSPE_OK);
    std::vector<std::string> err(n_gpus);
    std::vector<std::thread> th;
    for (int g = 0; g < n_gpus; g++) {
        th.emplace_back([&, g] {
            int r = vspe_create(g, &ctx[g]);
            if (r == VSPE_OK && sparse) ctx[g]->opt_sparse = 1;
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

    const uint64_t nn = (uint64_t)n_nodes * n_nodes;
    if (result == VSPE_OK && sparse) {
        // sparse runs: host-mediated merge into device 0 (append + radix sort + run-length reduce)
        for (int g = 1; g < n_gpus && result == VSPE_OK; g++) {
            uint64_t ne = 0;
            const uint64_t *pk = nullptr, *pc = nullptr;
            cudaSetDevice(g);
            result = vspe_sparse_host(ctx[g], &ne, &pk, &pc);
            if (result == VSPE_OK) { cudaSetDevice(0); result = vspe_sparse_merge(ctx[0], pk, pc, ne); }
        }
        if (result == VSPE_OK) {
            uint64_t ne = 0;
            const uint64_t *pk = nullptr, *pc = nullptr;
            cudaSetDevice(0);
            result = vspe_sparse_host(ctx[0], &ne, &pk, &pc);
            if (result == VSPE_OK) { sparse_keys->assign(pk, pk + ne); sparse_counts->assign(pc, pc + ne); }
        }
    }
    if (result == VSPE_OK && nn && !sparse) {
        std::vector<ncclComm_t> comms(n_gpus);
        std::vector<int> devs(n_gpus);
        for (int g = 0; g < n_gpus; g++) devs[g] = g;
        ncclResult_t nr = nccl.CommInitAll(comms.data(), n_gpus, devs.data());
        if (nr != 0) { set_error("ncclCommInitAll: %s", nccl.GetErrorString(nr)); result = VSPE_ERR_NCCL; }
        if (result == VSPE_OK) {
            nccl.GroupStart();
            for (int g = 0; g < n_gpus; g++) {
                cudaSetDevice(g);
                nr = nccl.AllReduce(ctx[g]->mats.p, ctx[g]->mats.p, 2 * nn, kNcclUint64, kNcclSum, comms[g], ctx[g]->stream);
                if (nr != 0) break;
            }
            ncclResult_t ge = nccl.GroupEnd();
            if (nr != 0 || ge != 0) { set_error("ncclAllReduce: %s", nccl.GetErrorString(nr ? nr : ge)); result = VSPE_ERR_NCCL; }
            for (int g = 0; g < n_gpus; g++) { cudaSetDevice(g); cudaStreamSynchronize(ctx[g]->stream); }
            for (int g = 0; g < n_gpus; g++) nccl.CommDestroy(comms[g]);
        }
    }
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
