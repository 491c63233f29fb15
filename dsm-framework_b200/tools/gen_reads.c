/*
 * gen_reads.c -- seeded synthetic metagenome read generator (input preparation only).
 *
 * The reference's toy data is download-only (README.md:62-73), so tests and bench.py
 * draw reads from random genomes as SURVEY.md section 8(d) specifies: genomes i.i.d.
 * uniform ACGT; each read picks a genome, a start and a strand uniformly (reverse
 * strand = reverse complement), then every base is substituted with probability `sub`
 * (uniform over the other three bases) and replaced by 'N' with probability `pn`.
 * All randomness is SplitMix64; genome g of the pool uses stream (pool_seed, g) and
 * read i uses stream (seed, i), so output does not depend on the thread count.
 *
 * Outputs: FASTA text (">r<i>\n<bases>\n"), or the documents exactly as the builder
 * CLI hands them to InsertText (builder.cpp:183-201: complement(read) + '-' +
 * reverse(read)), each followed by '\0'.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define GEN_API __attribute__((visibility("default")))

typedef struct dsmgen_params {
    uint64_t seed;        /* read stream seed                                   */
    uint64_t pool_seed;   /* genome pool seed                                   */
    uint32_t pool_size;   /* genomes in the shared pool                         */
    uint32_t n_genomes;   /* genomes this sample draws from the pool (<= pool)  */
    uint64_t genome_len;
    uint64_t n_reads;
    uint32_t read_len;
    uint32_t reserved;
    double sub;           /* per-base substitution probability                  */
    double pn;            /* per-base N probability                             */
} dsmgen_params;

static inline uint64_t sm64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t mix2(uint64_t a, uint64_t b)
{
    uint64_t s = a * 0xD6E8FEB86659FD93ull + b;
    sm64(&s);
    return sm64(&s);
}
static inline double u01(uint64_t *s) { return (double)(sm64(s) >> 11) * (1.0 / 9007199254740992.0); }

static const char kBases[4] = {'A', 'C', 'G', 'T'};

static uint8_t *make_genomes(const dsmgen_params *p, uint32_t **pick_out)
{
    /* which pool genomes this sample uses: a seeded partial shuffle of 0..pool-1 */
    uint32_t *idx = (uint32_t *)malloc(sizeof(uint32_t) * p->pool_size);
    uint64_t s = mix2(p->seed, 0x5eedull);
    for (uint32_t i = 0; i < p->pool_size; ++i) idx[i] = i;
    for (uint32_t i = 0; i < p->n_genomes; ++i) {
        uint32_t j = i + (uint32_t)(sm64(&s) % (p->pool_size - i));
        uint32_t t = idx[i]; idx[i] = idx[j]; idx[j] = t;
    }
    uint8_t *g = (uint8_t *)malloc((size_t)p->n_genomes * p->genome_len);
    if (!g) { free(idx); return NULL; }
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)p->n_genomes; ++k) {
        uint64_t st = mix2(p->pool_seed, idx[k]);
        uint8_t *dst = g + (size_t)k * p->genome_len;
        uint64_t i = 0;
        while (i < p->genome_len) {
            uint64_t r = sm64(&st);
            for (int b = 0; b < 32 && i < p->genome_len; ++b, ++i) dst[i] = kBases[(r >> (2 * b)) & 3];
        }
    }
    *pick_out = idx;
    return g;
}

static inline uint8_t comp(uint8_t c)
{
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; }
}

static void make_read(const dsmgen_params *p, const uint8_t *genomes, uint64_t i, uint8_t *out)
{
    uint64_t st = mix2(p->seed, i + 1);
    const uint32_t L = p->read_len;
    const uint8_t *g = genomes + (size_t)(sm64(&st) % p->n_genomes) * p->genome_len;
    const uint64_t start = sm64(&st) % (p->genome_len - L + 1);
    const int rev = (int)(sm64(&st) & 1);
    for (uint32_t k = 0; k < L; ++k) out[k] = rev ? comp(g[start + L - 1 - k]) : g[start + k];
    if (p->sub > 0.0 || p->pn > 0.0) {
        for (uint32_t k = 0; k < L; ++k) {
            if (p->sub > 0.0 && u01(&st) < p->sub) {
                uint8_t c;
                do { c = (uint8_t)kBases[sm64(&st) & 3]; } while (c == out[k]);
                out[k] = c;
            }
            if (p->pn > 0.0 && u01(&st) < p->pn) out[k] = 'N';
        }
    }
}

static int check(const dsmgen_params *p)
{
    return p && p->pool_size && p->n_genomes && p->n_genomes <= p->pool_size && p->read_len &&
           p->genome_len >= p->read_len;
}

static int digits(uint64_t v) { int d = 1; while (v >= 10) { v /= 10; ++d; } return d; }

/* bytes of the FASTA text for these parameters */
GEN_API uint64_t dsmgen_fasta_size(const dsmgen_params *p)
{
    if (!check(p)) return 0;
    uint64_t total = 0, lo = 0, width = 1, hi = 10;
    while (lo < p->n_reads) { /* ">r" + digits + "\n" + read + "\n" */
        uint64_t cnt = (p->n_reads < hi ? p->n_reads : hi) - lo;
        total += cnt * (2 + width + 1 + p->read_len + 1);
        lo = hi; hi *= 10; ++width;
    }
    return total;
}

GEN_API int dsmgen_fasta(const dsmgen_params *p, uint8_t *out, uint64_t cap)
{
    if (!check(p) || cap < dsmgen_fasta_size(p)) return -1;
    uint32_t *pick = NULL;
    uint8_t *g = make_genomes(p, &pick);
    if (!g) return -2;
    /* record offsets are a closed form of i, so records can be written in parallel */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)p->n_reads; ++i) {
        uint64_t off = 0, lo = 0, width = 1, hi = 10;
        while (hi <= (uint64_t)i) { off += (hi - lo) * (2 + width + 1 + p->read_len + 1); lo = hi; hi *= 10; ++width; }
        off += ((uint64_t)i - lo) * (2 + width + 1 + p->read_len + 1);
        uint8_t *d = out + off;
        *d++ = '>'; *d++ = 'r';
        int nd = digits((uint64_t)i);
        uint64_t v = (uint64_t)i;
        for (int k = nd - 1; k >= 0; --k) { d[k] = (uint8_t)('0' + v % 10); v /= 10; }
        d += nd;
        *d++ = '\n';
        make_read(p, g, (uint64_t)i, d);
        d[p->read_len] = '\n';
    }
    free(g); free(pick);
    return 0;
}

/* reads as an n_reads x read_len byte matrix */
GEN_API int dsmgen_reads(const dsmgen_params *p, uint8_t *out)
{
    if (!check(p)) return -1;
    uint32_t *pick = NULL;
    uint8_t *g = make_genomes(p, &pick);
    if (!g) return -2;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)p->n_reads; ++i) make_read(p, g, (uint64_t)i, out + (size_t)i * p->read_len);
    free(g); free(pick);
    return 0;
}

/* documents as InsertText receives them, '\0'-terminated: n_reads * (2*read_len + 2) bytes */
GEN_API int dsmgen_docs(const dsmgen_params *p, uint8_t *out)
{
    if (!check(p)) return -1;
    uint32_t *pick = NULL;
    uint8_t *g = make_genomes(p, &pick);
    if (!g) return -2;
    const uint32_t L = p->read_len;
    const size_t D = 2 * (size_t)L + 2;
#pragma omp parallel
    {
        uint8_t *r = (uint8_t *)malloc(L);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)p->n_reads; ++i) {
            make_read(p, g, (uint64_t)i, r);
            uint8_t *d = out + (size_t)i * D;
            for (uint32_t k = 0; k < L; ++k) d[k] = comp(r[k]);
            d[L] = '-';
            for (uint32_t k = 0; k < L; ++k) d[L + 1 + k] = r[L - 1 - k];
            d[2 * L + 1] = 0;
        }
        free(r);
    }
    free(g); free(pick);
    return 0;
}

#ifdef DSMGEN_MAIN
/* gen_fasta seed pool_seed pool_size n_genomes genome_len n_reads read_len sub pn > out.fasta */
int main(int argc, char **argv)
{
    if (argc != 10) { fprintf(stderr, "usage: %s seed pool_seed pool_size n_genomes genome_len n_reads read_len sub pn\n", argv[0]); return 2; }
    dsmgen_params p;
    memset(&p, 0, sizeof p);
    p.seed = strtoull(argv[1], 0, 10); p.pool_seed = strtoull(argv[2], 0, 10);
    p.pool_size = (uint32_t)strtoul(argv[3], 0, 10); p.n_genomes = (uint32_t)strtoul(argv[4], 0, 10);
    p.genome_len = strtoull(argv[5], 0, 10); p.n_reads = strtoull(argv[6], 0, 10);
    p.read_len = (uint32_t)strtoul(argv[7], 0, 10); p.sub = atof(argv[8]); p.pn = atof(argv[9]);
    uint64_t sz = dsmgen_fasta_size(&p);
    uint8_t *buf = (uint8_t *)malloc(sz ? sz : 1);
    if (!sz || !buf || dsmgen_fasta(&p, buf, sz)) { fprintf(stderr, "bad parameters\n"); return 1; }
    fwrite(buf, 1, sz, stdout);
    free(buf);
    return 0;
}
#endif
