/*
 * include/bspgemm_host.h — host-side C surface of the drop-in (libbspgemm_host.so): Matrix Market
 * input, COO -> CSC/CSR conversion, timing, synthetic generators.  Mirrors the reference's host
 * interface for the hot path: same names, argument meaning and error behaviour.
 *
 *   readCOO            final/utils.h:13,  final/utils.c:47-81
 *   coo2csc            final/coo2csc.h:5-13, final/coo2csc.c:22-64
 *   tictoc / tic / toc final/utils.h:7-8, final/utils.c:104-113
 *   mm_read_banner, mm_read_mtx_crd_size, MM_typecode, mm_is_* / mm_set_* , MM_* error codes
 *                      final/mmio.h:16-85, final/mmio.c:96-217   (see host/mmio_compat.h)
 */
#ifndef BSPGEMM_HOST_H
#define BSPGEMM_HOST_H

#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Reads a Matrix Market coordinate file into compressed form.  Exactly like the reference:
 * `*row` receives the M+1 pointers, indexed by the file's COLUMN index J; `*col` receives the file's ROW
 * indices I in file order inside each J (stable) — i.e. CSC of the file matrix = CSR of its transpose
 * (SURVEY.md §3.4).  Both arrays are malloc'ed here, the caller frees them (final/SpGEMM_mpi_omp.c:339-340).
 * Open / banner / size failures print a message and exit(1) like the reference (final/utils.c:54-61).
 * Beyond the reference: value columns of real/integer/complex files are skipped instead of being mis-read
 * as indices, and entries outside [1,M]x[1,N] are an error instead of a wild write. */
void readCOO(const char *mat, uint32_t **row, uint32_t **col, uint32_t *M, uint32_t *N, uint32_t *nnz);
/* Same, returning a status instead of exiting (0 = ok). */
int readCOO_status(const char *mat, uint32_t **row, uint32_t **col, uint32_t *M, uint32_t *N, uint32_t *nnz);
/* Same, with the COO -> CSC step done by `convert` (coo2csc's argument list, 0 = ok; NULL = the host coo2csc below).
 * libbspgemm.so's bspgemm_coo2csc (include/bspgemm.h) fits: parse on the host, sort on the GPU. */
typedef int (*bs_coo2csc_fn)(uint32_t *row, uint32_t *col, const uint32_t *row_coo, const uint32_t *col_coo,
                             uint32_t nnz, uint32_t n, uint32_t isOneBased);
int readCOO_convert(const char *mat, uint32_t **row, uint32_t **col, uint32_t *M, uint32_t *N, uint32_t *nnz, bs_coo2csc_fn convert);

/* Stable counting sort of COO entries by col_coo (final/coo2csc.c:22-64).  `row`: nnz indices out,
 * `col`: n+1 pointers out.  Same parameter order and meaning as the reference. */
void coo2csc(uint32_t *const row, uint32_t *const col,
             uint32_t const *const row_coo, uint32_t const *const col_coo,
             uint32_t const nnz, uint32_t const n, uint32_t const isOneBased);

/* CLOCK_MONOTONIC stopwatch (final/utils.c:104-113): tictoc(0) starts, tictoc(1) returns seconds. */
double tictoc(int mode);
#define tic tictoc(0)
#define toc tictoc(1)

/* mean / lower-median / fastest of `times` samples, computed exactly like final/SpGEMM_mpi_omp.c:330-333
 * (sort ascending, median = element (times-1)/2).  Sorts `t` in place. */
void bs_time_stats(double *t, int times, double *mean, double *median, double *fastest);

/* ---- synthetic boolean matrices (the reference's own generator is Matlab/write_spm.m:5-8) ----
 * All return CSR with sorted, duplicate-free rows; arrays malloc'ed here (caller frees); 0 = ok.
 *   uniform : every row draws d columns uniformly in [0,n) from splitmix64(seed,row,slot), sort+unique
 *   rmat    : n = 2^scale, edge_factor*n edge draws with quadrant probabilities (a,b,c), duplicates removed
 *   banded  : row i holds columns i-d/2 .. i+d/2-1 clipped to [0,n)
 *   blockdiag: dense d x d blocks on the diagonal */
int bs_gen_uniform(uint32_t n, uint32_t d, uint64_t seed, int32_t **row, int32_t **col, int64_t *nnz);
/* the reference's own distribution, Matlab sprand(n,n,d/n)>0 (Matlab/write_spm.m:5): Poisson(d) row lengths */
int bs_gen_sprand(uint32_t n, double d, uint64_t seed, int32_t **row, int32_t **col, int64_t *nnz);
int bs_gen_rmat(uint32_t scale, uint32_t edge_factor, double a, double b, double c, uint64_t seed,
                int32_t **row, int32_t **col, int64_t *nnz);
int bs_gen_banded(uint32_t n, uint32_t d, int32_t **row, int32_t **col, int64_t *nnz);
int bs_gen_blockdiag(uint32_t n, uint32_t d, int32_t **row, int32_t **col, int64_t *nnz);

/* Writes in-memory CSR (rows r, columns c) as "%%MatrixMarket matrix coordinate pattern general" such that
 * readCOO() reproduces exactly these arrays: entry (r,c) is written as the line "c+1 r+1", grouped by r
 * (the column-major order Matlab's mmwrite emits, Matlab/write_spm.m:8). */
int bs_write_mtx(const char *path, uint32_t n, const int32_t *row, const int32_t *col);

#ifdef __cplusplus
}
#endif
#endif
