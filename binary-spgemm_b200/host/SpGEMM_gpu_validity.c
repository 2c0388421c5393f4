/* host/SpGEMM_gpu_validity.c — validity driver, mirrors final/SpGEMM_mpi_omp_validity.c:308-375.
 *
 *   SpGEMM_gpu_validity <matrix.mtx> <block_size> <threads>
 *
 * The reference multiplies once with the hybrid MPI+OpenMP path and once with the serial kernel
 * (SpGEMM_bigslice over all rows, :337) and compares the two CSR results element for element
 * (SpGEMM_valid, :290-302).  Here the two independent paths are
 *   (1) the sharded operator bspgemm_csr (all GPUs, fused one-pass pipeline), and
 *   (2) the slice operator bspgemm_csr_slice over rows [0,An) on GPU 0 forced to the two-phase pipeline
 *       (symbolic -> scan -> numeric), i.e. different kernels and a different row-pointer construction.
 * Prints the reference's two fixed strings (:340,:342); unlike the reference the exit status reflects
 * the outcome.  Comparison against the CPU oracle lives in tests/, not here. */
#include <stdio.h>
#include <stdlib.h>
#include <stdbool.h>
#include "../../include/bspgemm.h"
#include "../../include/bspgemm_host.h"

static bool SpGEMM_valid(const int *Acol, const int *Arow, const int *Bcol, const int *Brow, int n)
{
    for (int i = 0; i <= n; i++) if (Arow[i] != Brow[i]) return false;
    for (int i = 0; i < Arow[n]; i++) if (Acol[i] != Bcol[i]) return false;
    return true;
}

int main(int argc, char const *argv[])
{
    if (argc != 4) {
        printf("usage: [BSPGEMM_GPUS=numgpus]  SpGEMM_gpu_validity  path-to-matrix  threadslice_size  number_of_threads\n");
        exit(1);
    }
    const char *eg = getenv("BSPGEMM_GPUS");
    int st = bspgemm_init(eg ? atoi(eg) : 1);
    if (st != BSPGEMM_OK) { fprintf(stderr, "bspgemm_init: %s: %s\n", bspgemm_strerror(st), bspgemm_last_error()); exit(1); }

    uint32_t *Arow, *Acol, M, N, Annz;
    readCOO(argv[1], &Arow, &Acol, &M, &N, &Annz);
    const int An = (int)N;
    if (M != N) { fprintf(stderr, "SpGEMM_gpu_validity: C = A*A needs a square matrix\n"); exit(1); }

    int *nCrow = (int *)calloc((size_t)An + 1, sizeof(int)), *nCcol = NULL;
    int *tCrow = (int *)calloc((size_t)An + 1, sizeof(int)), *tCcol = NULL;

    setenv("BSPGEMM_MODE", "fused", 1);
    st = bspgemm_csr((int *)Acol, (int *)Arow, An, (int *)Acol, (int *)Arow, An, (int)M, &nCcol, nCrow);
    if (st != BSPGEMM_OK) { fprintf(stderr, "bspgemm_csr: %s: %s\n", bspgemm_strerror(st), bspgemm_last_error()); exit(1); }
    bspgemm_finalize();

    setenv("BSPGEMM_MODE", "twophase", 1);      /* read when the per-GPU contexts are (re)created */
    st = bspgemm_init(1);
    if (st == BSPGEMM_OK) st = bspgemm_csr_slice((int *)Acol, (int *)Arow, An, (int *)Acol, (int *)Arow, An, (int)M, &tCcol, tCrow, 0, An);
    if (st != BSPGEMM_OK) { fprintf(stderr, "bspgemm_csr_slice: %s: %s\n", bspgemm_strerror(st), bspgemm_last_error()); exit(1); }

    const bool same = SpGEMM_valid(nCcol, nCrow, tCcol, tCrow, An);
    if (same) printf("Results of serial and multricore are the same!\n");
    else      printf("The results dont match\n");

    free(Acol); free(Arow); free(nCcol); free(nCrow); free(tCcol); free(tCrow);
    bspgemm_finalize();
    return same ? 0 : 2;
}
