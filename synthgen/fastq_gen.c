/*
 * fastq_gen.c -- fast, deterministic paired-end read simulator for the bench workloads
 * (SURVEY.md section 8d read model; test / bench tooling, not part of the product path).
 *
 * Every pair is generated from a counter-based random stream keyed by (seed, pair index), so any
 * sub-range of pairs can be produced independently (OpenMP over pairs; the CPU baselines take a
 * prefix of exactly the bytes the GPU arm sees) and the output does not depend on the thread count.
 *
 *   strain ~ abundance; insert ~ N(2*rl + 100, 30^2) clipped to [rl, G]; start uniform;
 *   mate1 = frag[:rl], mate2 = revcomp(frag)[:rl]; mates swapped with p = 0.5;
 *   per-base substitutions at sub_rate; n_rate of the pairs get one 'N'; short_rate of the mates are
 *   truncated to 1..k bases; records "@p%09llu/<mate>\n<seq>\n+\n<I*len>\n" (2*len + 18 bytes).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const uint8_t* genomes;   /* [n_genomes][strains][G] base codes 0..3 (A C G T) */
    uint64_t n_genomes, strains, G;
    const double* ab_cdf;     /* [strains] cumulative abundances, last = 1 */
    uint32_t read_len, k;
    double sub_rate, n_rate, short_rate;
    uint64_t seed;
} fq_params;

static inline uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
typedef struct { uint64_t s; } rng_t;
static inline uint64_t next_u64(rng_t* r) { r->s += 0x9E3779B97F4A7C15ull; return mix(r->s); }
static inline double next_unit(rng_t* r) { return (double)(next_u64(r) >> 11) * (1.0 / 9007199254740992.0); }

typedef struct { uint64_t off; uint32_t start, ins, len[2]; uint16_t ncol; uint8_t genome_hi, swap, nmate; uint32_t genome, strain; } pair_plan;

static void plan_pair(const fq_params* p, uint64_t idx, pair_plan* q) {
    rng_t r = { mix(p->seed ^ mix(idx)) };
    const uint32_t rl = p->read_len;
    q->genome = (uint32_t)(next_u64(&r) % p->n_genomes);
    const double u = next_unit(&r);
    uint32_t s = 0;
    while (s + 1 < p->strains && u >= p->ab_cdf[s]) s++;
    q->strain = s;
    /* Box-Muller */
    double u1 = next_unit(&r), u2 = next_unit(&r);
    if (u1 < 1e-300) u1 = 1e-300;
    const double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    double ins = rint(2.0 * rl + 100.0 + 30.0 * z);
    if (ins < rl) ins = rl;
    if (ins > (double)p->G) ins = (double)p->G;
    q->ins = (uint32_t)ins;
    q->start = (uint32_t)(next_unit(&r) * (double)(p->G - q->ins + 1));
    q->swap = (uint8_t)(next_u64(&r) & 1);
    q->nmate = 0;
    if (next_unit(&r) < p->n_rate) { q->nmate = (uint8_t)(1 + (next_u64(&r) & 1)); q->ncol = (uint16_t)(next_u64(&r) % rl); }
    for (int m = 0; m < 2; m++) {
        q->len[m] = rl;
        if (next_unit(&r) < p->short_rate) q->len[m] = 1 + (uint32_t)(next_u64(&r) % p->k);
    }
}

static const char ACGT[4] = {'A', 'C', 'G', 'T'};

static void write_mate(const fq_params* p, uint64_t idx, int mate, const pair_plan* q, uint8_t* out) {
    const uint32_t rl = p->read_len, len = q->len[mate];
    const uint8_t* g = p->genomes + ((uint64_t)q->genome * p->strains + q->strain) * p->G;
    uint8_t seq[1024];
    /* file `mate` (0 = forward file) holds fragment end (mate ^ swap): 0 = frag[:rl], 1 = revcomp(frag)[:rl] */
    const int end = mate ^ q->swap;
    if (end == 0) for (uint32_t i = 0; i < rl; i++) seq[i] = g[q->start + i];
    else for (uint32_t i = 0; i < rl; i++) seq[i] = (uint8_t)(3 - g[q->start + q->ins - 1 - i]);
    /* substitutions: geometric skipping */
    rng_t r = { mix(p->seed ^ mix(idx * 2 + 1 + (uint64_t)mate) ^ 0xA5A5A5A5ull) };
    if (p->sub_rate > 0) {
        const double lq = log(1.0 - p->sub_rate);
        double pos = -1.0;
        while (1) {
            double u = next_unit(&r);
            if (u < 1e-300) u = 1e-300;
            pos += floor(log(u) / lq) + 1.0;
            if (pos >= rl) break;
            const uint32_t i = (uint32_t)pos;
            seq[i] = (uint8_t)((seq[i] + 1 + next_u64(&r) % 3) & 3);
        }
    }
    /* header */
    uint8_t* o = out;
    *o++ = '@'; *o++ = 'p';
    uint64_t v = idx;
    for (int d = 8; d >= 0; d--) { o[d] = (uint8_t)('0' + v % 10); v /= 10; }
    o += 9;
    *o++ = '/'; *o++ = (uint8_t)('1' + mate); *o++ = '\n';
    for (uint32_t i = 0; i < len; i++) o[i] = (uint8_t)ACGT[seq[i]];
    if (q->nmate == 1 + mate && q->ncol < len) o[q->ncol] = 'N';
    o += len;
    *o++ = '\n'; *o++ = '+'; *o++ = '\n';
    memset(o, 'I', len);
    o += len;
    *o++ = '\n';
}

/* Generates pairs [first, first + n).  out_f / out_r need n * (2 * read_len + 18) bytes at most;
 * the exact sizes are returned.  Returns 0, or -1 for bad parameters. */
int fq_generate(const fq_params* p, uint64_t first, uint64_t n, uint8_t* out_f, uint8_t* out_r, uint64_t* size_f, uint64_t* size_r) {
    if (!p || p->read_len == 0 || p->read_len > 1000 || p->G < p->read_len || p->strains == 0 || p->n_genomes == 0) return -1;
    uint64_t* off_f = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
    uint64_t* off_r = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
    if (!off_f || !off_r) { free(off_f); free(off_r); return -1; }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        pair_plan q;
        plan_pair(p, first + (uint64_t)i, &q);
        off_f[i + 1] = 2ull * q.len[0] + 18;
        off_r[i + 1] = 2ull * q.len[1] + 18;
    }
    off_f[0] = off_r[0] = 0;
    for (uint64_t i = 0; i < n; i++) { off_f[i + 1] += off_f[i]; off_r[i + 1] += off_r[i]; }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        pair_plan q;
        plan_pair(p, first + (uint64_t)i, &q);
        write_mate(p, first + (uint64_t)i, 0, &q, out_f + off_f[i]);
        write_mate(p, first + (uint64_t)i, 1, &q, out_r + off_r[i]);
    }
    *size_f = off_f[n];
    *size_r = off_r[n];
    free(off_f);
    free(off_r);
    return 0;
}
