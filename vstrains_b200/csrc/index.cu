// index.cu -- K3: (k+1)-mer hash index over the node sequences and their reverse complements.
//
// Replaces the Python dict build of reference utils/VStrains_PE_Inference.py:117-135
// (kmer_htable[kmer].append((i, sub_i)) for the forward k-mer and for reverse_seq(kmer)).
// Layout and semantics are described in vspe_internal.cuh (IndexView).
#include "ctx.cuh"

namespace vspe {

// --- pack both strands of every node into the 2-bit text ---------------------------------
__global__ void __launch_bounds__(256)
k_pack_text(const uint8_t* __restrict__ seqs, const uint64_t* __restrict__ seq_off,
            const uint32_t* __restrict__ strand_start, const uint32_t* __restrict__ node_len,
            uint32_t n2, uint32_t text_len, uint64_t* __restrict__ text, uint32_t n_words) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint64_t b0 = (uint64_t)w * 32;
    uint64_t word = 0;
    if (b0 < text_len) {
        uint32_t q = strand_of(strand_start, n2, (uint32_t)b0);
        for (uint32_t j = 0; j < 32; j++) {
            uint64_t b = b0 + j;
            if (b >= text_len) break;
            while (b >= strand_start[q + 1]) q++;
            uint32_t i = q >> 1, s = q & 1;
            uint32_t o = (uint32_t)(b - strand_start[q]);
            uint32_t len = node_len[i];
            uint32_t c = s ? seqs[seq_off[i] + (len - 1 - o)] : seqs[seq_off[i] + o];
            uint64_t code = base_code(c) ^ (s ? 2u : 0u);
            word |= code << (2 * j);
        }
    }
    text[w] = word;
}

// --- one table entry per (k+1)-mer occurrence -----------------------------------------------
__global__ void __launch_bounds__(256)
k_index_insert(IndexView ix, unsigned long long* __restrict__ slots64, uint32_t* __restrict__ bloom) {
    uint32_t tp = blockIdx.x * blockDim.x + threadIdx.x;
    if (tp >= ix.text_len) return;
    uint32_t q = strand_of(ix.strand_start, 2 * ix.n_nodes, tp);
    if ((uint64_t)tp + ix.split_len > ix.strand_start[q + 1]) return;
    uint64_t h = hash_packed(ix.text, tp, ix.split_len);
    uint32_t meta = ((uint32_t)h & ~ix.node_mask) | (q >> 1);
    unsigned long long entry = ((unsigned long long)meta << 32) | tp;
    if (bloom) {
        uint32_t block, b0, b1, b2;
        bloom_bits(h, ix.bloom_mask, block, b0, b1, b2);
        atomicOr(bloom + 8 * (size_t)block + (b0 >> 5), 1u << (b0 & 31));
        atomicOr(bloom + 8 * (size_t)block + (b1 >> 5), 1u << (b1 & 31));
        atomicOr(bloom + 8 * (size_t)block + (b2 >> 5), 1u << (b2 & 31));
    }
    uint32_t j = slot_of(h, ix.slot_mask);
    while (true) {
        unsigned long long old = atomicCAS(&slots64[j], EMPTY_SLOT, entry);
        if (old == EMPTY_SLOT) break;
        j = (j + 1) & ix.slot_mask;
    }
}

// number of postings whose (k+1)-mer equals the window at text position tp
__device__ __forceinline__ uint32_t count_postings_text(const IndexView& ix, uint32_t tp, uint64_t h) {
    uint32_t cnt = 0;
    uint32_t j = slot_of(h, ix.slot_mask);
    while (true) {
        uint2 e = __ldg(ix.slots + j);
        if (e.x == EMPTY_TP) break;
        if (fp_match(e.y, h, ix.node_mask) && (e.x == tp || text_equal(ix.text, e.x, tp, ix.split_len))) cnt++;
        j = (j + 1) & ix.slot_mask;
    }
    return cnt;
}

// --- uniq bitmap: exactly one posting in total for the k-mer starting at tp -------------------
__global__ void __launch_bounds__(256)
k_index_unique(IndexView ix, uint32_t* __restrict__ uniq, uint32_t n_words) {
    uint32_t tp = blockIdx.x * blockDim.x + threadIdx.x;   // blockDim multiple of 32
    bool flag = false;
    if (tp < ix.text_len) {
        uint32_t q = strand_of(ix.strand_start, 2 * ix.n_nodes, tp);
        if ((uint64_t)tp + ix.split_len <= ix.strand_start[q + 1]) {
            uint64_t h = hash_packed(ix.text, tp, ix.split_len);
            flag = count_postings_text(ix, tp, h) == 1;
        }
    }
    uint32_t m = __ballot_sync(0xFFFFFFFFu, flag);
    if ((threadIdx.x & 31) == 0 && (tp >> 5) < n_words) uniq[tp >> 5] = m;
}

// --- successor table: for the last window of every strand and every next base b, the text
// position and node {tp, node} of the k-mer (window[1:] + b) if it has exactly one posting,
// else {NONE32, 0}.
// This is a precomputed table lookup, so it is exact for any graph, not only de Bruijn ones.
__global__ void __launch_bounds__(128)
k_index_succ(IndexView ix, uint32_t* __restrict__ succ) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t q = t >> 2, b = t & 3;
    if (q >= 2 * ix.n_nodes) return;
    uint32_t s0 = ix.strand_start[q], s1 = ix.strand_start[q + 1];
    uint32_t L = ix.split_len;
    uint32_t res = NONE32, res_node = 0;
    if (s1 - s0 >= L) {
        uint64_t from = (uint64_t)s1 - L + 1;        // last L-1 bases of the strand
        KmerHash hs;
        for (uint32_t m = 0; m < L; m += 32) {
            uint64_t w = extract64(ix.text, from + m);
            uint32_t rem = L - m;                     // bases of the query in this word
            // position L-1 of the query is the appended base b
            if (rem <= 32) {
                uint64_t keep = rem - 1 == 0 ? 0ull : ((1ull << (2 * (rem - 1))) - 1);
                w = (w & keep) | ((uint64_t)b << (2 * (rem - 1)));
            }
            hs.add64(w, rem);
        }
        const uint64_t h = hs.finish();
        uint32_t cnt = 0, found = NONE32, found_node = 0;
        uint32_t j = slot_of(h, ix.slot_mask);
        while (true) {
            uint2 e = __ldg(ix.slots + j);
            if (e.x == EMPTY_TP) break;
            if (fp_match(e.y, h, ix.node_mask)) {
                // compare L-1 overlap bases, then the appended base
                bool eq = text_equal(ix.text, e.x, from, L - 1) && text_base(ix.text, (uint64_t)e.x + L - 1) == b;
                if (eq) { cnt++; found = e.x; found_node = e.y & ix.node_mask; }
            }
            j = (j + 1) & ix.slot_mask;
        }
        if (cnt == 1) { res = found; res_node = found_node; }
    }
    succ[2 * t] = res;
    succ[2 * t + 1] = res_node;
}

// the successor table again, widened to everything a walk needs to enter the successor strand
__global__ void __launch_bounds__(128)
k_index_succ16(IndexView ix, const uint32_t* __restrict__ succ, uint4* __restrict__ succ16) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 8 * ix.n_nodes) return;
    const uint32_t tp = succ[2 * t], node = succ[2 * t + 1];
    uint4 out = make_uint4(NONE32, 0, 0, 0);
    if (tp != NONE32) {
        const uint4 nr = ix.node_rec[node];
        const bool rcs = tp >= nr.y;
        out = make_uint4(tp, 2 * node + (rcs ? 1u : 0u), rcs ? nr.z : nr.y, nr.w);
    }
    succ16[t] = out;
}

// --- substitution-hit bitmap -----------------------------------------------------------------
// For every window w of every strand, every offset o and every other base b: does the k-mer
// "window w with base o replaced by b" have a posting?  If yes, bit b of text base w+o is set.
// The seed-and-extend tier uses a CLEAR bit as a proof that all windows covering a mismatching
// read base miss (they equal text windows except for that one base), so it can skip probing them.
// The polynomial hash makes each substituted hash an O(1) update of the window's accumulators.
__global__ void __launch_bounds__(128)
k_index_subst(IndexView ix, uint32_t* __restrict__ subst, uint32_t inv1, uint32_t inv2) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= ix.text_len) return;
    const uint32_t L = ix.split_len;
    const uint32_t q = strand_of(ix.strand_start, 2 * ix.n_nodes, w);
    if ((uint64_t)w + L > ix.strand_start[q + 1]) return;
    const uint32_t n = (L + 15) >> 4;
    KmerHash base;
    for (uint32_t m = 0; m < L; m += 32) base.add64(extract64(ix.text, (uint64_t)w + m), L - m);
    // power of the multiplier that weighs 16-base word k: M^(n-1-k)
    uint32_t p1 = 1, p2 = 1;
    for (uint32_t k = 1; k < n; k++) { p1 *= 0x9E3779B1u; p2 *= 0x85EBCA77u; }
    for (uint32_t o = 0; o < L; o++) {
        if (o && (o & 15) == 0) { p1 *= inv1; p2 *= inv2; }
        const uint32_t cur = text_base(ix.text, (uint64_t)w + o);
        for (uint32_t b = 0; b < 4; b++) {
            if (b == cur) continue;
            const uint32_t delta = (b - cur) << (2 * (o & 15));
            KmerHash hs = base;
            hs.h1 += delta * p1;
            hs.h2 += delta * p2;
            const uint64_t h = hs.finish();
            if (ix.bloom != nullptr && !bloom_maybe(ix.bloom, ix.bloom_mask, h)) continue;      // proven miss (HBM-sized tables)
            uint32_t j = slot_of(h, ix.slot_mask);
            bool found = false;
            while (!found) {
                const uint2 e = __ldg(ix.slots + j);
                if (e.x == EMPTY_TP) break;
                if (fp_match(e.y, h, ix.node_mask)) {
                    bool eq = true;
                    for (uint32_t m = 0; m < L && eq; m += 32) {
                        uint64_t x = extract64(ix.text, (uint64_t)e.x + m) ^ extract64(ix.text, (uint64_t)w + m);
                        const uint32_t rem = L - m;
                        if (rem < 32) x &= (1ull << (2 * rem)) - 1;
                        const uint64_t want = (o >= m && o < m + 32) ? ((uint64_t)(cur ^ b) << (2 * (o - m))) : 0ull;
                        eq = x == want;
                    }
                    found = eq;
                }
                j = (j + 1) & ix.slot_mask;
            }
            if (found) {
                const uint32_t pos = w + o;
                atomicOr(&subst[pos >> 3], 1u << (4 * (pos & 7) + b));
            }
        }
    }
}

static uint32_t inv32(uint32_t a) {        // inverse of an odd number mod 2^32 (Newton)
    uint32_t x = a;
    for (int i = 0; i < 5; i++) x *= 2u - a * x;
    return x;
}

int index_build_device(Ctx* c, const uint8_t* seqs, const uint64_t* seq_off, uint32_t n_nodes, uint32_t split_len) {
    Index& ix = c->index;
    ix.built = false;
    if (split_len < 2) { set_error("split_len must be >= 2 (kmer_size >= 1)"); return VSPE_ERR_ARG; }
    // host validation == the reference's failure modes
    std::vector<uint32_t> h_ss(2 * (size_t)n_nodes + 1), h_len(n_nodes ? n_nodes : 1);
    uint64_t text_len = 0, n_kmers = 0;
    for (uint32_t i = 0; i < n_nodes; i++) {
        uint64_t len = seq_off[i + 1] - seq_off[i];
        if (len > 0x7FFFFFFFull) { set_error("node %u longer than 2^31", i); return VSPE_ERR_LIMIT; }
        h_len[i] = (uint32_t)len;
        const uint8_t* s = seqs + seq_off[i];
        bool indexed = len >= split_len;
        if (indexed) {
            for (uint64_t j = 0; j < len; j++)
                if (!is_acgt(s[j])) {
                    set_error("node #%u: character 0x%02x at offset %llu is not one of ACGT "
                              "(the reference raises KeyError in reverse_seq)", i, s[j], (unsigned long long)j);
                    return VSPE_ERR_NODE_SEQ;
                }
            n_kmers += 2 * (len - split_len + 1);
        }
        h_ss[2 * i] = (uint32_t)text_len;
        if (indexed) text_len += len;
        h_ss[2 * i + 1] = (uint32_t)text_len;
        if (indexed) text_len += len;
        if (text_len >= 0xFFFFFF00ull) { set_error("graph too large: packed text exceeds 2^32 bases"); return VSPE_ERR_LIMIT; }
    }
    h_ss[2 * (size_t)n_nodes] = (uint32_t)text_len;
    uint32_t nbits = 1;
    while (nbits < 31 && (1ull << nbits) < n_nodes) nbits++;
    // load <= 1/4 while the table stays L2 sized (a miss -- the usual answer for a window with sequencing errors -- then
    // ends after 1.4 slots on average instead of 2.5), <= 1/2 beyond
    uint64_t slots = 1024;
    while (slots < 2 * n_kmers + 2) slots <<= 1;
    if (slots * sizeof(uint2) <= (32ull << 20)) slots <<= 1;
    if (slots > 0x80000000ull) { set_error("graph too large: hash table exceeds 2^31 slots"); return VSPE_ERR_LIMIT; }

    cudaStream_t st = c->stream;
    uint32_t n_words = (uint32_t)((text_len + 31) / 32) + 4;
    VSPE_TRY(ix.text.reserve(n_words));
    VSPE_TRY(ix.strand_start.reserve(h_ss.size()));
    VSPE_TRY(ix.node_len.reserve(h_len.size()));
    VSPE_TRY(ix.node_rec.reserve(h_len.size()));
    std::vector<uint4> h_rec(h_len.size());
    for (uint32_t i = 0; i < n_nodes; i++) h_rec[i] = make_uint4(h_ss[2 * i], h_ss[2 * i + 1], h_ss[2 * i + 2], h_len[i]);
    VSPE_TRY(ix.slots.reserve(slots));
    VSPE_TRY(ix.uniq.reserve(n_words));
    VSPE_TRY(ix.succ.reserve(16 * (size_t)n_nodes + 16));
    VSPE_TRY(ix.succ16.reserve(8 * (size_t)n_nodes + 8));
    DevBuf<uint8_t> d_seqs;
    DevBuf<uint64_t> d_off;
    uint64_t seq_bytes = n_nodes ? seq_off[n_nodes] : 0;
    VSPE_TRY(d_seqs.reserve(seq_bytes + 1));
    VSPE_TRY(d_off.reserve((size_t)n_nodes + 1));
    cudaEvent_t e0 = c->ev[0], e1 = c->ev[1];
    VSPE_CUDA(cudaEventRecord(e0, st));
    if (seq_bytes) VSPE_CUDA(cudaMemcpyAsync(d_seqs.p, seqs, seq_bytes, cudaMemcpyHostToDevice, st));
    VSPE_CUDA(cudaMemcpyAsync(d_off.p, seq_off, ((size_t)n_nodes + 1) * 8, cudaMemcpyHostToDevice, st));
    VSPE_CUDA(cudaMemcpyAsync(ix.strand_start.p, h_ss.data(), h_ss.size() * 4, cudaMemcpyHostToDevice, st));
    VSPE_CUDA(cudaMemcpyAsync(ix.node_len.p, h_len.data(), (size_t)n_nodes * 4, cudaMemcpyHostToDevice, st));
    VSPE_CUDA(cudaMemcpyAsync(ix.node_rec.p, h_rec.data(), (size_t)n_nodes * sizeof(uint4), cudaMemcpyHostToDevice, st));
    VSPE_CUDA(cudaMemsetAsync(ix.slots.p, 0xFF, slots * sizeof(uint2), st));
    VSPE_CUDA(cudaMemsetAsync(ix.uniq.p, 0, (size_t)n_words * 4, st));

    // a slot table beyond L2 size (the 200 000-node stress graph: 2 GB) gets an L2-sized Bloom filter in front:
    // about 8 bits per (k+1)-mer occurrence, three of them set
    ix.has_bloom = false;
    if (slots * sizeof(uint2) > (96ull << 20)) {
        uint64_t blocks = 1024;
        while (blocks * 256 < n_kmers * 6) blocks <<= 1;       // 6..12 bits per k-mer
        VSPE_TRY(ix.bloom.reserve(blocks * 8));
        VSPE_CUDA(cudaMemsetAsync(ix.bloom.p, 0, blocks * 32, st));
        ix.bloom_mask = (uint32_t)(blocks - 1);
        ix.has_bloom = true;
    }
    ix.text_len = (uint32_t)text_len;
    ix.slot_mask = (uint32_t)(slots - 1);
    ix.node_mask = (1u << nbits) - 1;
    ix.split_len = split_len;
    ix.n_nodes = n_nodes;
    ix.n_kmers = n_kmers;
    IndexView v = ix.view();

    k_pack_text<<<(n_words + 255) / 256, 256, 0, st>>>(d_seqs.p, d_off.p, ix.strand_start.p, ix.node_len.p,
                                                       2 * n_nodes, ix.text_len, ix.text.p, n_words);
    VSPE_LAUNCH_CHECK(c);
    if (text_len) {
        uint32_t nb = (uint32_t)((text_len + 255) / 256);
        k_index_insert<<<nb, 256, 0, st>>>(v, (unsigned long long*)ix.slots.p, ix.has_bloom ? ix.bloom.p : nullptr);
        VSPE_LAUNCH_CHECK(c);
        k_index_unique<<<nb, 256, 0, st>>>(v, ix.uniq.p, n_words);
        VSPE_LAUNCH_CHECK(c);
    }
    if (n_nodes) {
        k_index_succ<<<(8 * n_nodes + 127) / 128, 128, 0, st>>>(v, ix.succ.p);
        VSPE_LAUNCH_CHECK(c);
        k_index_succ16<<<(8 * n_nodes + 127) / 128, 128, 0, st>>>(ix.view(), ix.succ.p, ix.succ16.p);
        VSPE_LAUNCH_CHECK(c);
    }
    // the substitution-hit bitmap costs 3*L probes per window: build it for viral-scale graphs only (the map kernels probe
    // the windows themselves when it is absent).  Measured on the 200 000-node stress graph (C5): 213 ms to build (Bloom
    // filter in front of the probes) for 10 ms less map time per 12.5 M pairs -- not worth it at that size.
    ix.has_subst = false;
    if (text_len && c->opt_subst && (double)n_kmers * split_len * 3.0 <= 6e9) {
        VSPE_TRY(ix.subst.reserve(n_words * 4 + 8));
        VSPE_CUDA(cudaMemsetAsync(ix.subst.p, 0, ((size_t)n_words * 4 + 8) * 4, st));
        k_index_subst<<<(uint32_t)((text_len + 127) / 128), 128, 0, st>>>(v, ix.subst.p, inv32(0x9E3779B1u), inv32(0x85EBCA77u));
        VSPE_LAUNCH_CHECK(c);
        ix.has_subst = true;
    }
    VSPE_CUDA(cudaEventRecord(e1, st));
    VSPE_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    c->stats.ms_index = ms;
    c->stats.n_nodes = n_nodes;
    c->stats.n_kmers = n_kmers;
    c->stats.table_slots = slots;
    // count matrices: dense [2][N][N] u64, unless the graph is too large (or sparse mode is forced)
    c->sparse.enabled = c->opt_sparse != 0 || !dense_possible(n_nodes);
    c->sparse.n_runs = 0;
    if (!c->sparse.enabled) {
        uint64_t nn = 2ull * n_nodes * n_nodes;
        VSPE_TRY(c->mats.reserve(nn ? nn : 1));
        VSPE_CUDA(cudaMemsetAsync(c->mats.p, 0, (nn ? nn : 1) * 8, st));
    }
    VSPE_CUDA(cudaStreamSynchronize(st));
    ix.built = true;
    return VSPE_OK;
}

}  // namespace vspe
