/* host/gen.c — seeded synthetic boolean matrices for the benchmark configurations and a
 * Matrix Market writer whose output readCOO() maps back to the same in-memory CSR.
 * The reference's generator is Matlab (`sprand(n,n,d/n)>0` + mmwrite, Matlab/write_spm.m:5-8); these are
 * the C stand-ins used by bench.py, the drivers' self-test and the parity tests. */
#include "../../include/bspgemm_host.h"
#include "mmio_compat.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* counter-based: value for (seed, a, b) */
static inline uint64_t rng3(uint64_t seed, uint64_t a, uint64_t b) { return mix64(mix64(seed ^ mix64(a)) + b); }

static void sort_u32(int32_t *v, int n)
{
    for (int i = 1; i < n; ++i) { int32_t x = v[i]; int j = i - 1; while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; } v[j + 1] = x; }
}
static int cmp_i32(const void *a, const void *b) { int32_t x = *(const int32_t *)a, y = *(const int32_t *)b; return (x > y) - (x < y); }

int bs_gen_uniform(uint32_t n, uint32_t d, uint64_t seed, int32_t **row_out, int32_t **col_out, int64_t *nnz_out)
{
    int32_t *row = (int32_t *)malloc(((size_t)n + 1) * sizeof(int32_t));
    int32_t *tmp = (int32_t *)malloc((size_t)n * d * sizeof(int32_t) + 4);
    int32_t *len = (int32_t *)malloc(((size_t)n + 1) * sizeof(int32_t));
    if (!row || !tmp || !len) { free(row); free(tmp); free(len); return 1; }
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n; ++r) {
        int32_t *v = tmp + (size_t)r * d;
        for (uint32_t s = 0; s < d; ++s) v[s] = (int32_t)(((rng3(seed, (uint64_t)r, s) >> 32) * (uint64_t)n) >> 32);
        if (d <= 64) sort_u32(v, (int)d); else qsort(v, d, sizeof(int32_t), cmp_i32);
        int m = 0;
        for (uint32_t s = 0; s < d; ++s) if (s == 0 || v[s] != v[s - 1]) v[m++] = v[s];
        len[r] = m;
    }
    row[0] = 0;
    for (uint32_t r = 0; r < n; ++r) row[r + 1] = row[r] + len[r];
    int32_t *col = (int32_t *)malloc(((size_t)row[n] + 1) * sizeof(int32_t));
    if (!col) { free(row); free(tmp); free(len); return 1; }
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n; ++r) memcpy(col + row[r], tmp + (size_t)r * d, (size_t)len[r] * sizeof(int32_t));
    free(tmp); free(len);
    *row_out = row; *col_out = col; *nnz_out = row[n];
    return 0;
}

/* The reference's own distribution: Matlab `sprand(n,n,d/n) > 0` (Matlab/write_spm.m:5) — about d*n entries at uniformly random
 * positions, i.e. row lengths are Binomial(n, d/n) ~ Poisson(d): ragged rows, unlike bs_gen_uniform's fixed d per row.  The
 * length of row r is drawn by inversion of the Poisson distribution (Knuth's product of uniforms, d is small), its columns are
 * uniform in [0,n), sorted, repeats removed — all from the counter-based generator, so the matrix is a function of (n, d, seed). */
int bs_gen_sprand(uint32_t n, double d, uint64_t seed, int32_t **row_out, int32_t **col_out, int64_t *nnz_out)
{
    if (d < 0 || d > 64) return 1;
    const int cap = (int)(d * 4 + 32);                       /* P(len > cap) is astronomically small; lengths are clamped */
    int32_t *row = (int32_t *)malloc(((size_t)n + 1) * sizeof(int32_t));
    int32_t *tmp = (int32_t *)malloc((size_t)n * cap * sizeof(int32_t) + 4);
    int32_t *len = (int32_t *)malloc(((size_t)n + 1) * sizeof(int32_t));
    if (!row || !tmp || !len) { free(row); free(tmp); free(len); return 1; }
    const double limit = exp(-d);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n; ++r) {
        int k = 0;
        double prod = 1.0;
        for (;;) {                                           /* Poisson(d): number of uniforms whose product stays above e^-d */
            prod *= (double)((rng3(seed ^ 0x5eedull, (uint64_t)r, (uint64_t)(1000 + k)) >> 11) + 1) * (1.0 / 9007199254740993.0);
            if (prod <= limit || k >= cap) break;
            ++k;
        }
        int32_t *v = tmp + (size_t)r * cap;
        for (int s = 0; s < k; ++s) v[s] = (int32_t)(((rng3(seed, (uint64_t)r, (uint64_t)s) >> 32) * (uint64_t)n) >> 32);
        sort_u32(v, k);
        int m = 0;
        for (int s = 0; s < k; ++s) if (s == 0 || v[s] != v[s - 1]) v[m++] = v[s];
        len[r] = m;
    }
    row[0] = 0;
    for (uint32_t r = 0; r < n; ++r) row[r + 1] = row[r] + len[r];
    int32_t *col = (int32_t *)malloc(((size_t)row[n] + 1) * sizeof(int32_t));
    if (!col) { free(row); free(tmp); free(len); return 1; }
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n; ++r) memcpy(col + row[r], tmp + (size_t)r * cap, (size_t)len[r] * sizeof(int32_t));
    free(tmp); free(len);
    *row_out = row; *col_out = col; *nnz_out = row[n];
    return 0;
}

/* rows sorted + unique from an edge list (r[e], c[e]) */
static int edges_to_csr(uint32_t n, int64_t m, const int32_t *er, const int32_t *ec, int32_t **row_out, int32_t **col_out, int64_t *nnz_out)
{
    int64_t *ptr = (int64_t *)calloc((size_t)n + 2, sizeof(int64_t));
    int32_t *buf = (int32_t *)malloc(((size_t)m + 1) * sizeof(int32_t));
    int32_t *row = (int32_t *)malloc(((size_t)n + 1) * sizeof(int32_t));
    if (!ptr || !buf || !row) { free(ptr); free(buf); free(row); return 1; }
    for (int64_t e = 0; e < m; ++e) ptr[er[e] + 2]++;
    for (uint32_t r = 0; r < n; ++r) ptr[r + 2] += ptr[r + 1];
    for (int64_t e = 0; e < m; ++e) buf[ptr[er[e] + 1]++] = ec[e];       /* now ptr[r+1] = end of row r, ptr[r] = start */
    int64_t w = 0;
    row[0] = 0;
    for (uint32_t r = 0; r < n; ++r) {
        int64_t s = ptr[r], t = ptr[r + 1];
        qsort(buf + s, (size_t)(t - s), sizeof(int32_t), cmp_i32);
        for (int64_t p = s; p < t; ++p) if (p == s || buf[p] != buf[p - 1]) buf[w++] = buf[p];
        row[r + 1] = (int32_t)w;
    }
    free(ptr);
    *row_out = row; *col_out = buf; *nnz_out = w;
    return 0;
}

int bs_gen_rmat(uint32_t scale, uint32_t edge_factor, double a, double b, double c, uint64_t seed,
                int32_t **row_out, int32_t **col_out, int64_t *nnz_out)
{
    if (scale < 1 || scale > 30) return 1;
    const uint32_t n = 1u << scale;
    const int64_t m = (int64_t)edge_factor * n;
    int32_t *er = (int32_t *)malloc((size_t)m * sizeof(int32_t) + 4), *ec = (int32_t *)malloc((size_t)m * sizeof(int32_t) + 4);
    if (!er || !ec) { free(er); free(ec); return 1; }
    const uint64_t ta = (uint64_t)(a * 4294967296.0), tb = (uint64_t)((a + b) * 4294967296.0), tc = (uint64_t)((a + b + c) * 4294967296.0);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < m; ++e) {
        uint32_t r = 0, cc = 0;
        for (uint32_t l = 0; l < scale; ++l) {
            const uint64_t u = rng3(seed, (uint64_t)e, l) >> 32;
            const int q = u < ta ? 0 : u < tb ? 1 : u < tc ? 2 : 3;       /* quadrant: 0=a (0,0) 1=b (0,1) 2=c (1,0) 3=d (1,1) */
            r = (r << 1) | (uint32_t)(q >> 1);
            cc = (cc << 1) | (uint32_t)(q & 1);
        }
        er[e] = (int32_t)r; ec[e] = (int32_t)cc;
    }
    int rc = edges_to_csr(n, m, er, ec, row_out, col_out, nnz_out);
    free(er); free(ec);
    return rc;
}

int bs_gen_banded(uint32_t n, uint32_t d, int32_t **row_out, int32_t **col_out, int64_t *nnz_out)
{
    int32_t *row = (int32_t *)malloc(((size_t)n + 1) * sizeof(int32_t));
    if (!row) return 1;
    const int64_t half = d / 2;
    int64_t w = 0;
    row[0] = 0;
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        int64_t lo = i - half, hi = i - half + (int64_t)d;     /* [i-d/2, i+d/2) */
        if (lo < 0) lo = 0;
        if (hi > (int64_t)n) hi = n;
        w += hi - lo;
        row[i + 1] = (int32_t)w;
    }
    int32_t *col = (int32_t *)malloc(((size_t)w + 1) * sizeof(int32_t));
    if (!col) { free(row); return 1; }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        int64_t lo = i - half; if (lo < 0) lo = 0;
        for (int32_t p = row[i]; p < row[i + 1]; ++p) col[p] = (int32_t)(lo + (p - row[i]));
    }
    *row_out = row; *col_out = col; *nnz_out = w;
    return 0;
}

int bs_gen_blockdiag(uint32_t n, uint32_t d, int32_t **row_out, int32_t **col_out, int64_t *nnz_out)
{
    int32_t *row = (int32_t *)malloc(((size_t)n + 1) * sizeof(int32_t));
    if (!row || d == 0) { free(row); return 1; }
    int64_t w = 0;
    row[0] = 0;
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        int64_t b0 = (i / d) * d, b1 = b0 + d; if (b1 > (int64_t)n) b1 = n;
        w += b1 - b0;
        row[i + 1] = (int32_t)w;
    }
    int32_t *col = (int32_t *)malloc(((size_t)w + 1) * sizeof(int32_t));
    if (!col) { free(row); return 1; }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        int64_t b0 = (i / d) * d;
        for (int32_t p = row[i]; p < row[i + 1]; ++p) col[p] = (int32_t)(b0 + (p - row[i]));
    }
    *row_out = row; *col_out = col; *nnz_out = w;
    return 0;
}

int bs_write_mtx(const char *path, uint32_t n, const int32_t *row, const int32_t *col)
{
    FILE *f = fopen(path, "w");
    if (!f) return 1;
    MM_typecode t;
    mm_initialize_typecode(&t); mm_set_matrix(&t); mm_set_coordinate(&t); mm_set_pattern(&t); mm_set_general(&t);
    static char big[1 << 20];
    setvbuf(f, big, _IOFBF, sizeof big);
    mm_write_banner(f, t);
    fprintf(f, "%% in-memory CSR row r / column c is stored as the entry (c+1, r+1): readCOO transposes on read\n");
    mm_write_mtx_crd_size(f, (int)n, (int)n, row[n]);
    /* "c+1 r+1\n" per entry, digits produced by hand into a 1 MB chunk (fprintf: 1.9 s for 8.4e6 entries) */
    static char out[(1 << 20) + 64];
    size_t used = 0;
    int ok = 1;
    for (uint32_t r = 0; r < n && ok; ++r)
        for (int32_t p = row[r]; p < row[r + 1]; ++p) {
            uint32_t v[2] = { (uint32_t)col[p] + 1u, r + 1u };
            for (int k = 0; k < 2; ++k) {
                char tmp[12]; int len = 0;
                uint32_t x = v[k];
                do { tmp[len++] = (char)('0' + x % 10u); x /= 10u; } while (x);
                while (len) out[used++] = tmp[--len];
                out[used++] = k ? '\n' : ' ';
            }
            if (used >= (1u << 20)) { if (fwrite(out, 1, used, f) != used) { ok = 0; break; } used = 0; }
        }
    if (ok && used && fwrite(out, 1, used, f) != used) ok = 0;
    return (fclose(f) == 0 && ok) ? 0 : 1;
}
