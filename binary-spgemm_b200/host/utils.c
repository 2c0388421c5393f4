/* host/utils.c — Matrix Market reader, stopwatch, timing statistics for the drivers.
 * Interface and observable behaviour follow final/utils.c:47-113 and final/SpGEMM_mpi_omp.c:330-333;
 * the implementation is new (whole-file read + hand tokenizer, run by all host threads when the file is one entry per line,
 * instead of one fscanf per entry). */
#define _POSIX_C_SOURCE 200809L
#include "utils.h"
#include "mmio_compat.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* parse the next unsigned integer token; returns 0 at end of buffer, 2 on a non-numeric token */
static inline int next_uint(const char **pp, const char *end, uint64_t *out)
{
    const char *p = *pp;
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
    if (p >= end) { *pp = p; return 0; }
    if (*p < '0' || *p > '9') { *pp = p; return 2; }
    uint64_t v = 0;
    while (p < end && *p >= '0' && *p <= '9') v = v * 10 + (uint64_t)(*p++ - '0');
    *out = v; *pp = p;
    return 1;
}
static inline void skip_token(const char **pp, const char *end)
{
    const char *p = *pp;
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
    while (p < end && !(*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
    *pp = p;
}

static inline int is_ws(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r'; }

/* Parallel tokenizer for the regular case — one entry per line, `2 + extra` tokens on it, every index in range: the buffer is cut
 * into one piece per thread at line ends, the pieces' non-blank lines are counted, a prefix sum gives every piece its first entry
 * number, then the pieces are parsed independently.  Returns 1 when the nz entries were read; 0 when anything at all is unusual
 * (fewer lines than entries, a short / long / non-numeric line, an index out of range): the caller then runs the sequential
 * tokenizer, whose verdict (and error code) stays the single definition of the format. */
static int parse_entries_parallel(const char *buf, size_t len, int nz, int m, int n, int extra, uint32_t *I, uint32_t *J)
{
#ifdef _OPENMP
    int T = omp_get_max_threads();
    if (T > 64) T = 64;
    if (T < 4 || len < ((size_t)4 << 20)) return 0;       /* two passes over the text: pays from four threads on */
    size_t cut[65], first[65];
    cut[0] = 0; cut[T] = len;
    for (int t = 1; t < T; ++t) {
        size_t p = len / (size_t)T * (size_t)t;
        while (p < len && buf[p] != '\n') ++p;
        cut[t] = p < len ? p + 1 : len;
    }
    size_t lines[64];
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; ++t) {
        size_t c = 0;
        int blank = 1;
        for (size_t p = cut[t]; p < cut[t + 1]; ++p) {
            if (buf[p] == '\n') { c += !blank; blank = 1; }
            else if (!is_ws(buf[p])) blank = 0;
        }
        c += !blank;                                  /* last line of the buffer without a newline */
        lines[t] = c;
    }
    first[0] = 0;
    for (int t = 0; t < T; ++t) first[t + 1] = first[t] + lines[t];
    if (first[T] < (size_t)nz) return 0;
    int odd = 0;
#pragma omp parallel for num_threads(T) schedule(static, 1) reduction(|:odd)
    for (int t = 0; t < T; ++t) {
        const char *p = buf + cut[t], *end = buf + cut[t + 1];
        size_t e = first[t];
        while (p < end && e < (size_t)nz && !odd) {
            const char *eol = p;
            while (eol < end && *eol != '\n') ++eol;
            const char *q = p;
            while (q < eol && is_ws(*q)) ++q;
            if (q < eol) {                             /* a non-blank line: exactly one entry */
                uint64_t v[2];
                for (int k = 0; k < 2; ++k) {
                    while (q < eol && is_ws(*q)) ++q;
                    if (q >= eol || *q < '0' || *q > '9') { odd = 1; break; }
                    uint64_t x = 0;
                    while (q < eol && *q >= '0' && *q <= '9') x = x * 10 + (uint64_t)(*q++ - '0');
                    if (q < eol && !is_ws(*q)) { odd = 1; break; }
                    v[k] = x;
                }
                if (odd) break;
                for (int k = 0; k < extra; ++k) {      /* value columns: present, skipped */
                    while (q < eol && is_ws(*q)) ++q;
                    if (q >= eol) { odd = 1; break; }
                    while (q < eol && !is_ws(*q)) ++q;
                }
                while (q < eol && is_ws(*q)) ++q;
                if (odd || q != eol || v[0] < 1 || v[0] > (uint64_t)m || v[1] < 1 || v[1] > (uint64_t)n) { odd = 1; break; }
                I[e] = (uint32_t)(v[0] - 1);
                J[e] = (uint32_t)(v[1] - 1);
                ++e;
            }
            p = eol < end ? eol + 1 : end;
        }
    }
    return !odd;
#else
    (void)buf; (void)len; (void)nz; (void)m; (void)n; (void)extra; (void)I; (void)J;
    return 0;
#endif
}

int readCOO_convert(const char *mat, uint32_t **row, uint32_t **col, uint32_t *M, uint32_t *N, uint32_t *nnz, bs_coo2csc_fn convert)
{
    MM_typecode code;
    int m = 0, n = 0, nz = 0, rc = 0;
    FILE *f = fopen(mat, "rb");
    if (!f) return MM_COULD_NOT_READ_FILE;
    if ((rc = mm_read_banner(f, &code)) != 0) { fclose(f); return rc; }
    if (!mm_is_coordinate(code)) { fclose(f); return MM_UNSUPPORTED_TYPE; }
    if ((rc = mm_read_mtx_crd_size(f, &m, &n, &nz)) != 0) { fclose(f); return rc; }
    if (m < 0 || n < 0 || nz < 0) { fclose(f); return MM_UNSUPPORTED_TYPE; }
    *M = (uint32_t)m; *N = (uint32_t)n; *nnz = (uint32_t)nz;

    long here = ftell(f);
    fseek(f, 0, SEEK_END);
    long size = ftell(f) - here;
    fseek(f, here, SEEK_SET);
    char *buf = (char *)malloc((size_t)size + 1);
    uint32_t *I = (uint32_t *)malloc(((size_t)nz + 1) * sizeof(uint32_t));
    uint32_t *J = (uint32_t *)malloc(((size_t)nz + 1) * sizeof(uint32_t));
    if (!buf || !I || !J) { free(buf); free(I); free(J); fclose(f); return MM_COULD_NOT_READ_FILE; }
    size_t got = fread(buf, 1, (size_t)size, f);
    fclose(f);

    /* entries: "i j" (pattern) | "i j v" (real, integer) | "i j re im" (complex); values are ignored —
     * every stored entry is `true` (the reference reads pairs only, final/utils.c:68) */
    const int extra = mm_is_pattern(code) ? 0 : mm_is_complex(code) ? 2 : 1;
    const char *p = buf, *end = buf + got;
    const int fast = parse_entries_parallel(buf, got, nz, m, n, extra, I, J);     /* all threads; 0 = let the loop below decide */
    for (int e = fast ? nz : 0; e < nz; ++e) {
        uint64_t i = 0, j = 0;
        if (next_uint(&p, end, &i) != 1 || next_uint(&p, end, &j) != 1) { rc = MM_PREMATURE_EOF; break; }
        for (int k = 0; k < extra; ++k) skip_token(&p, end);
        if (i < 1 || i > (uint64_t)m || j < 1 || j > (uint64_t)n) { rc = MM_UNSUPPORTED_TYPE; break; }
        I[e] = (uint32_t)(i - 1);                      /* 1-based -> 0-based (final/utils.c:69-70) */
        J[e] = (uint32_t)(j - 1);
    }
    free(buf);
    if (rc) { free(I); free(J); return rc; }

    /* compressed by the file's column J, indices = the file's rows I (final/utils.c:77): transpose-on-read.
     * Pointer array has N+1 entries (N == M for the square matrices the reference supports). */
    *row = (uint32_t *)malloc(((size_t)n + 1) * sizeof(uint32_t));
    *col = (uint32_t *)malloc(((size_t)nz + 1) * sizeof(uint32_t));
    if (!*row || !*col) { free(I); free(J); return MM_COULD_NOT_READ_FILE; }
    if (convert) {                                     /* e.g. bspgemm_coo2csc: the same stable sort on the GPU */
        if (convert(*col, *row, I, J, (uint32_t)nz, (uint32_t)n, 0) != 0) rc = MM_UNSUPPORTED_TYPE;
    } else {
        coo2csc(*col, *row, I, J, (uint32_t)nz, (uint32_t)n, 0);
    }
    free(I); free(J);
    if (rc) { free(*row); free(*col); *row = *col = NULL; }
    return rc;
}

int readCOO_status(const char *mat, uint32_t **row, uint32_t **col, uint32_t *M, uint32_t *N, uint32_t *nnz)
{
    return readCOO_convert(mat, row, col, M, N, nnz, NULL);
}

void readCOO(const char *mat, uint32_t **row, uint32_t **col, uint32_t *M, uint32_t *N, uint32_t *nnz)
{
    int rc = readCOO_status(mat, row, col, M, N, nnz);
    if (rc == 0) return;
    if (rc == MM_NO_HEADER || rc == MM_PREMATURE_EOF || rc == MM_UNSUPPORTED_TYPE || rc == MM_NOT_MTX)
        printf("Could not process Matrix Market banner.\n");      /* same message as final/utils.c:57 */
    exit(1);
}

double tictoc(int mode)
{
    static struct timespec t0;
    struct timespec t1;
    if (mode == 0) { clock_gettime(CLOCK_MONOTONIC, &t0); return 0.0; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-9;
}

static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

void bs_time_stats(double *t, int times, double *mean, double *median, double *fastest)
{
    double sum = 0;
    for (int i = 0; i < times; ++i) sum += t[i];
    qsort(t, (size_t)times, sizeof(double), cmp_double);
    *mean = times > 0 ? sum / times : 0.0;
    *median = times > 0 ? t[(times - 1) / 2] : 0.0;
    *fastest = times > 0 ? t[0] : 0.0;
}
