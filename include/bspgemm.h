/*
 * include/bspgemm.h — C ABI of libbspgemm.so: boolean (pattern-only) CSR SpGEMM  C = A·B  on B200 (sm_100a).
 *
 * This is the drop-in boundary for the reference's hot path (pavlidic/Binary-SpGEMM, final/):
 * plain C, plain pointers and sizes, no torch / C++ types.  Every entry point names the reference
 * interface it replaces (file:line under /root/reference).  There is NO CPU fallback behind any of
 * them: without a usable GPU they return BSPGEMM_ERR_NOGPU.
 *
 * Conventions kept from the reference:
 *   - argument order "col before row" (final/SpGEMM_mpi_omp.c:155-158);
 *   - A.row pointers are ABSOLUTE offsets into Acol, so a row block can be passed as a shifted Arow
 *     pointer with the unshifted Acol, exactly like `&Arow[rank*tasksize]` (:171);
 *   - *Ccol is allocated by the callee with malloc() and released by the caller with free() (:115/:200, :327);
 *   - Crow is caller-allocated, An+1 entries (:311);
 *   - result contract: Crow[0]=0; Ccol strictly ascending inside each row; empty rows allowed.
 * Additions the reference lacks: explicit Bn (rows of B — the reference never needs it on the CPU),
 * int status codes instead of void/exit(1), a 64-bit row-pointer variant (nnz(C) >= 2^31), any An and any
 * number of GPUs (the reference silently drops rows unless An % (tasks*tBlock) == 0, :77,:165).
 */
#ifndef BSPGEMM_H
#define BSPGEMM_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (the reference returns void and exit(1)s, final/utils.c:54-61) ---- */
#define BSPGEMM_OK             0
#define BSPGEMM_ERR_CUDA       1   /* a CUDA runtime call failed (message in bspgemm_last_error) */
#define BSPGEMM_ERR_NCCL       2   /* NCCL could not be loaded / a collective failed */
#define BSPGEMM_ERR_OOM        3   /* host or device allocation failed */
#define BSPGEMM_ERR_OVERFLOW32 4   /* nnz(C) >= 2^31 on the 32-bit row-pointer ABI: use the _i64 entry */
#define BSPGEMM_ERR_BADARG     5   /* NULL / negative sizes / column index outside [0,Bn) or [0,Bm) */
#define BSPGEMM_ERR_NOGPU      6   /* no CUDA device (there is no CPU fallback) */
#define BSPGEMM_ERR_CAPACITY   7   /* caller-provided Ccol buffer too small (needed nnz is reported) */
#define BSPGEMM_ERR_STATE      8   /* bspgemm_init not called / called twice */

const char *bspgemm_strerror(int status);
const char *bspgemm_last_error(void);          /* detail of the last failure on this thread */
const char *bspgemm_version(void);

/* ---- process-wide context: stands in for MPI_Init_thread / MPI_Comm_size / MPI_Finalize
 *      (final/SpGEMM_mpi_omp.c:352-355,:364).  ngpus = number of "tasks"; 0 = all visible GPUs.
 *      With ngpus > 1 an NCCL communicator over the GPUs is created (ncclCommInitAll). ---- */
int bspgemm_init(int ngpus);
/* Same with an explicit device list (e.g. one process per GPU under torchrun: {LOCAL_RANK}). */
int bspgemm_init_devices(const int *devices, int ngpus);
int bspgemm_finalize(void);
int bspgemm_num_gpus(void);                    /* 0 before init */

/* ---- host-pointer operators (inputs and outputs in host memory; H2D/D2H inside) ----
 * Replaces SpGEMM_mpi (final/SpGEMM_mpi_omp.c:155-225): A is split into contiguous row blocks, one per
 * GPU (the rank split :165-171); B is uploaded once to GPU 0 and replicated with ncclBroadcast (the
 * reference replicates B by having every rank read the file, :309); shards are gathered to the host at
 * their displacements with the row-pointer offset applied on the device (replaces :178-223). */
int bspgemm_csr(const int *Acol, const int *Arow, int An,
                const int *Bcol, const int *Brow, int Bn, int Bm,
                int **Ccol, int *Crow);
/* same, 64-bit row pointers for C (A and B stay 32-bit) */
int bspgemm_csr_i64(const int *Acol, const int *Arow, int An,
                    const int *Bcol, const int *Brow, int Bn, int Bm,
                    int **Ccol, int64_t *Crow);
/* Caller-allocated output, replaces SpGEMM_mat (Matlab/inc/BSpGEMM.h:2-4; Matlab/inc/BSpGEMM.c:9-47).
 * Ccol_buf may be pinned memory.  *nnz_out is always set; BSPGEMM_ERR_CAPACITY if capacity < nnz(C). */
int bspgemm_csr_into(const int *Acol, const int *Arow, int An,
                     const int *Bcol, const int *Brow, int Bn, int Bm,
                     int *Ccol_buf, int64_t capacity, int *Crow, int64_t *nnz_out);
/* Rows [start_row,end_row) only, slice-relative Crow (Crow[0]=0, end_row-start_row+1 entries):
 * replaces SpGEMM_bigslice (final/SpGEMM_mpi_omp.c:15-58).  Runs on GPU 0. */
int bspgemm_csr_slice(const int *Acol, const int *Arow, int An,
                      const int *Bcol, const int *Brow, int Bn, int Bm,
                      int **Ccol, int *Crow, int start_row, int end_row);
/* Masked product C = F .* (A·B): replaces SpGEMM_masked (final/SpGEMM_mpi_omp.c:232-288).  Row i of the result
 * = the distinct columns of row i of A·B that occur in row i of the mask F (An x Bm, CSR, any column order, repeats allowed —
 * the reference's mask is a flag array), ascending.  Same ownership rules as bspgemm_csr; runs on GPU 0 (the reference's
 * version is serial too). */
int bspgemm_csr_masked(const int *Acol, const int *Arow, int An,
                       const int *Bcol, const int *Brow, int Bn, int Bm,
                       const int *Fcol, const int *Frow,
                       int **Ccol, int *Crow);
/* Σ_{(i,j) in A} len(B_j): the intermediate-product count the throughput metric is quoted in
 * (trip count of final/SpGEMM_mpi_omp.c:33-37), computed by the work-estimation kernel on GPU 0. */
int bspgemm_intermediate_products(const int *Acol, const int *Arow, int An,
                                  const int *Brow, int Bn, int64_t *ip_out);

/* ---- distributed consumer: the product stays sharded on the GPUs instead of being gathered to one place
 *      (the reference gathers to rank 0: MPI_Gatherv + MPI_Gather + serial fix-up, final/SpGEMM_mpi_omp.c:203-223).
 *      bspgemm_csr_sharded runs the product of bspgemm_csr and returns a handle; the shards (device pointers into the contexts'
 *      arenas: valid until the next product / bspgemm_finalize) can be inspected, all-gathered GPU-to-GPU so that every GPU
 *      holds the whole CSR (row pointers offset on the owning GPU; ncclBroadcast per shard over NVLink), or written one file
 *      per shard (format 0: binary, 1: Matrix Market pattern, readable by readCOO). ---- */
typedef struct bspgemm_result bspgemm_result;
int bspgemm_csr_sharded(const int *Acol, const int *Arow, int An,
                        const int *Bcol, const int *Brow, int Bn, int Bm,
                        int crow_is_i64, bspgemm_result **out);
int bspgemm_result_shards(const bspgemm_result *r);
int64_t bspgemm_result_nnz(const bspgemm_result *r);
int bspgemm_result_shard(const bspgemm_result *r, int shard, int *device, int *row0, int *rows, int64_t *nnz, int64_t *disp,
                         const int **dCcol, const void **dCrow /* slice-relative, rows+1 entries */);
int bspgemm_result_allgather(bspgemm_result *r, int **dCcol_per_task, void **dCrow_per_task);
int bspgemm_result_write(const bspgemm_result *r, const char *prefix, int format);
int bspgemm_result_free(bspgemm_result *r);

/* Legacy-signature drop-ins: same names' worth of arguments as the reference, void return, print to
 * stderr and exit(1) on failure like the reference's I/O paths.  Bn is derived as max(Acol)+1 bounded by
 * nothing else, so Brow must have at least that many + 1 entries (true for any valid CSR pair).
 * tBlock is accepted and ignored (it only sizes the CPU thread slices, :77). */
void bspgemm_SpGEMM_mpi(int *Acol, int *Arow, int An, int *Bcol, int *Brow, int Bm,
                        int **Ccol, int *Crow, int tBlock);                       /* :155-158 */
void bspgemm_SpGEMM_masked(int *Acol, int *Arow, int An, int *Bcol, int *Brow, int Bm,
                           int *Fcol, int *Frow, int **Ccol, int *Crow, int *Csize);     /* :232-235 */
void bspgemm_SpGEMM_omp(int *Acol, int *Arow, int An, int *Bcol, int *Brow, int Bm,
                        int **Ccol, int *Crow, int tBlock);                       /* :71-74  (GPU 0 only) */
void bspgemm_SpGEMM_bigslice(int *Acol, int *Arow, int An, int *Bcol, int *Brow, int Bm,
                             int **Ccol, int *Crow, int *Csize,
                             int start_row, int end_row);                         /* :15-18 */

/* ---- device-resident operator (one GPU; what each torch.distributed rank / each shard calls) ----
 * All pointers are device pointers on `device`.  Work is enqueued on `stream` (a cudaStream_t; NULL = the CUDA
 * legacy default stream, i.e. ordered after the caller's earlier default-stream work) and the call returns after
 * the result size is known (it synchronises the stream after the probes and at the end; once, at the end, for a
 * product replayed from the plan of a prepared B).  dCrow: An+1 entries of 32- or 64-bit.
 * *dCcol_out points into an arena owned by the handle, valid until the next multiply / destroy. */
typedef struct bspgemm_dev bspgemm_dev;

#define BSPGEMM_MODE_AUTO     0   /* fused one-pass when the IP bound fits in memory, else two-phase */
#define BSPGEMM_MODE_FUSED    1   /* estimate -> [M/L symbolic] -> fused symbolic+scan+fill -> [M/L numeric] */
#define BSPGEMM_MODE_TWOPHASE 2   /* estimate -> symbolic -> scan -> numeric (north-star steps 1-4 as 4 launches) */

typedef struct bspgemm_stats {
  int64_t ip;                 /* intermediate products of the last multiply */
  int64_t nnz;                /* nnz(C) */
  int64_t rows_s, rows_m, rows_l;   /* rows per bin (warp / CTA shared-memory / CTA global bitmap) */
  int32_t mode;               /* BSPGEMM_MODE_FUSED or _TWOPHASE actually used */
  int32_t cap_s;              /* S-bin capacity (max IP handled by one warp) */
  int32_t group;              /* lanes cooperating on one B row (G) */
  int32_t launches;           /* kernels launched by the last multiply */
  float   ms_total;           /* CUDA-event time, first launch .. last launch */
  float   ms_estimate;        /* work-estimation kernel */
  float   ms_symbolic;        /* symbolic kernels (two-phase: all bins; fused: M/L bins only) */
  float   ms_main;            /* fused kernel (fused mode) or scan + S numeric kernel (two-phase) */
  float   ms_numeric;         /* M/L numeric kernels */
  int64_t algorithmic_bytes;  /* 4(An+1)+12nnzA+4IP+4nnzC+4(An+1) (8-byte terms for _i64 Crow) */
  int32_t variant;            /* 0 = CSR-gather kernels (kernels.cuh); ELL fast paths: 1 = ordered-table kernel (fused_ell.cuh),
                                 2 = register sorting network (fused_sort.cuh); then cap_s = table words per row, group =
                                 ELL width W, ms_symbolic = the CSR->ELL re-layout of B */
  int32_t rows_per_tile;      /* fused kernels: consecutive rows per look-back tile */
  int32_t kernel_flags;       /* bit 0: variant 2 ran k_fused_sort_async (cp.async input) rather than k_fused_sort; bit 1: floating-point network; bit 2: small rows ran as count -> scan -> fill (k_rows_tiny / k_rows_warp) instead of the ordered one-pass kernel */
  int32_t b_prepared;         /* 1: B was the matrix given to bspgemm_dev_prepare_b (its re-layout was not rebuilt) */
  int32_t plan_cached;        /* 1: launched from the cached plan of the previous product (no probes, one kernel) */
} bspgemm_stats;

int bspgemm_dev_create(bspgemm_dev **h, int device);
int bspgemm_dev_destroy(bspgemm_dev *h);
int bspgemm_dev_set_mode(bspgemm_dev *h, int mode);
int bspgemm_dev_multiply(bspgemm_dev *h, void *stream,
                         const int *dAcol, const int *dArow, int An, int64_t Annz,
                         const int *dBcol, const int *dBrow, int Bn, int Bm, int64_t Bnnz,
                         void *dCrow, int crow_is_i64,
                         int **dCcol_out, int64_t *nnz_out);
/* Device-resident masked product (see bspgemm_csr_masked).  dFcol/dFrow: the mask, An+1 row pointers (absolute offsets, like
 * A's).  *dCcol_out points into a second arena of the handle (the unmasked rows occupy the first), valid until the next call. */
int bspgemm_dev_multiply_masked(bspgemm_dev *h, void *stream,
                                const int *dAcol, const int *dArow, int An, int64_t Annz,
                                const int *dBcol, const int *dBrow, int Bn, int Bm, int64_t Bnnz,
                                const int *dFcol, const int *dFrow, int64_t Fnnz,
                                void *dCrow, int crow_is_i64,
                                int **dCcol_out, int64_t *nnz_out);
int bspgemm_dev_get_stats(bspgemm_dev *h, bspgemm_stats *out);
/* B resident once for many products — the reference replicates B once, before its timed loop (every rank parses the file,
 * final/SpGEMM_mpi_omp.c:309 vs :318-328).  Builds B's gather-friendly copy in the handle (ELL re-layout, or run descriptors
 * for banded matrices) and remembers the plan of the next product with it: later bspgemm_dev_multiply calls that pass the SAME
 * dBcol/dBrow/Bn/Bm/Bnnz skip the re-layout and the probe kernels.  Contract: the caller does not modify B until the next
 * bspgemm_dev_prepare_b / bspgemm_dev_forget_b.  A is free to change between products (a product the cached plan does not fit
 * is detected on the device and redone with fresh probes).  Results are identical with and without it. */
int bspgemm_dev_prepare_b(bspgemm_dev *h, void *stream,
                          const int *dBcol, const int *dBrow, int Bn, int Bm, int64_t Bnnz);
int bspgemm_dev_forget_b(bspgemm_dev *h);

/* ---- COO -> CSC/CSR on the device.  Replaces coo2csc (final/coo2csc.c:22-64, final/coo2csc.h:5-13): same
 * argument list and the same result — col[0..n] pointers by `col_coo`, row[] = the `row_coo` values of each column in INPUT
 * ORDER (stable), indices made 0-based — plus a status.  A key outside [0,n) is BSPGEMM_ERR_BADARG (undefined behaviour in the
 * reference).  nnz, n < 2^31.  Runs on the current CUDA device, no bspgemm_init needed; the _dev form takes device pointers
 * (row_coo/col_coo are not modified) and returns after the stream has been synchronised. */
int bspgemm_coo2csc(uint32_t *row, uint32_t *col, const uint32_t *row_coo, const uint32_t *col_coo,
                    uint32_t nnz, uint32_t n, uint32_t isOneBased);
int bspgemm_coo2csc_dev(void *stream, uint32_t *d_row, uint32_t *d_col, const uint32_t *d_row_coo, const uint32_t *d_col_coo,
                        uint32_t nnz, uint32_t n, uint32_t isOneBased);

#ifdef __cplusplus
}
#endif
#endif /* BSPGEMM_H */
