/* host/SpGEMM_gpu.c — performance driver with the reference's CLI and result line.
 *
 *   SpGEMM_gpu <matrix.mtx> <block_size> <threads> <times>
 *
 * Mirrors test_mpi + main of final/SpGEMM_mpi_omp.c:294-366: read the matrix with readCOO, compute
 * C = A·A `times` times (the driver passes A twice, :322), time each repetition with tic/toc (the timed
 * region holds everything from host CSR in to host CSR out, like :320-324), print
 *   tasks,threads,tasks*threads,block,path,An,Annz,Cnnz,mean,median,fastest        (:336)
 * "tasks" is the number of GPUs (BSPGEMM_GPUS, default 1 — stands in for `mpirun -n`); block_size and
 * threads are echoed for CSV compatibility (they only shape the CPU slices in the reference, :77).
 * A second line with GPU-side metrics goes to stderr so stdout stays byte-compatible.
 * BSPGEMM_GPU_COO2CSC=1: the reader's COO -> CSC step runs on the GPU too (readCOO_convert + bspgemm_coo2csc). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/bspgemm.h"
#include "../../include/bspgemm_host.h"

int main(int argc, char const *argv[])
{
    if (argc != 5) {
        printf("usage: [BSPGEMM_GPUS=numgpus]  SpGEMM_gpu  path-to-matrix  threadslice_size  number_of_threads  times_to_run\n");
        exit(1);
    }
    const int tBlock = atoi(argv[2]);
    const int threads = atoi(argv[3]);
    const int times = atoi(argv[4]);
    const char *eg = getenv("BSPGEMM_GPUS");
    int st = bspgemm_init(eg ? atoi(eg) : 1);
    if (st != BSPGEMM_OK) { fprintf(stderr, "bspgemm_init: %s: %s\n", bspgemm_strerror(st), bspgemm_last_error()); exit(1); }
    const int numtasks = bspgemm_num_gpus();

    uint32_t *Arow, *Acol, M, N, Annz;
    if (getenv("BSPGEMM_GPU_COO2CSC")) {          /* tokenizer on the host, COO -> CSC on the GPU (bspgemm_coo2csc) */
        int rc = readCOO_convert(argv[1], &Arow, &Acol, &M, &N, &Annz, bspgemm_coo2csc);
        if (rc) { fprintf(stderr, "SpGEMM_gpu: cannot read %s (status %d): %s\n", argv[1], rc, bspgemm_last_error()); exit(1); }
    } else {
        readCOO(argv[1], &Arow, &Acol, &M, &N, &Annz);
    }
    const int An = (int)N;                       /* pointers are indexed by the file's column (transpose-on-read) */
    if (M != N) { fprintf(stderr, "SpGEMM_gpu: C = A*A needs a square matrix (%u x %u)\n", M, N); exit(1); }

    int *nCrow = (int *)calloc((size_t)An + 1, sizeof(int));
    int *nCcol = NULL;
    double *alltimes = (double *)malloc((size_t)(times > 0 ? times : 1) * sizeof(double));
    int64_t ip = 0;
    bspgemm_intermediate_products((int *)Acol, (int *)Arow, An, (int *)Arow, An, &ip);

    for (int i = 0; i < times; i++) {
        tic;
        st = bspgemm_csr((int *)Acol, (int *)Arow, An, (int *)Acol, (int *)Arow, An, (int)M, &nCcol, nCrow);
        alltimes[i] = toc;
        if (st != BSPGEMM_OK) { fprintf(stderr, "SpGEMM_gpu: %s: %s\n", bspgemm_strerror(st), bspgemm_last_error()); exit(1); }
        free(nCcol);
    }
    double mean, median, fastest;
    bs_time_stats(alltimes, times, &mean, &median, &fastest);
    printf("%d,%d,%d,%d,%s,%d,%d,%d,%lf,%lf,%lf\n", numtasks, threads, numtasks * threads, tBlock, argv[1],
           An, (int)Annz, nCrow[An], mean, median, fastest);
    if (fastest > 0)
        fprintf(stderr, "# gpus=%d ip=%lld ip_per_s=%.4e out_nnz_per_s=%.4e (host CSR in -> host CSR out, fastest rep)\n",
                numtasks, (long long)ip, (double)ip / fastest, (double)nCrow[An] / fastest);
    free(alltimes); free(Acol); free(Arow); free(nCrow);
    bspgemm_finalize();
    return 0;
}
