/* host/bspgemm_gen.c — command-line front end of the synthetic generators (writes .mtx files that
 * SpGEMM_gpu / the reference binaries can read).  Stand-in for Matlab/write_spm.m.
 *   bspgemm_gen uniform  <n> <d> <seed> <out.mtx>
 *   bspgemm_gen rmat     <scale> <edge_factor> <a> <b> <c> <seed> <out.mtx>
 *   bspgemm_gen banded   <n> <d> <out.mtx>
 *   bspgemm_gen blockdiag <n> <d> <out.mtx> */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/bspgemm_host.h"

int main(int argc, char **argv)
{
    int32_t *row = NULL, *col = NULL; int64_t nnz = 0; uint32_t n = 0; const char *out = NULL; int rc = 1;
    if (argc == 6 && !strcmp(argv[1], "uniform")) { n = (uint32_t)atol(argv[2]); rc = bs_gen_uniform(n, (uint32_t)atoi(argv[3]), (uint64_t)atoll(argv[4]), &row, &col, &nnz); out = argv[5]; }
    else if (argc == 9 && !strcmp(argv[1], "rmat")) { n = 1u << atoi(argv[2]); rc = bs_gen_rmat((uint32_t)atoi(argv[2]), (uint32_t)atoi(argv[3]), atof(argv[4]), atof(argv[5]), atof(argv[6]), (uint64_t)atoll(argv[7]), &row, &col, &nnz); out = argv[8]; }
    else if (argc == 5 && !strcmp(argv[1], "banded")) { n = (uint32_t)atol(argv[2]); rc = bs_gen_banded(n, (uint32_t)atoi(argv[3]), &row, &col, &nnz); out = argv[4]; }
    else if (argc == 5 && !strcmp(argv[1], "blockdiag")) { n = (uint32_t)atol(argv[2]); rc = bs_gen_blockdiag(n, (uint32_t)atoi(argv[3]), &row, &col, &nnz); out = argv[4]; }
    else { fprintf(stderr, "usage: bspgemm_gen uniform n d seed out | rmat scale ef a b c seed out | banded n d out | blockdiag n d out\n"); return 1; }
    if (rc) { fprintf(stderr, "generation failed\n"); return 1; }
    rc = bs_write_mtx(out, n, row, col);
    fprintf(stderr, "%s: n=%u nnz=%lld\n", out, n, (long long)nnz);
    free(row); free(col);
    return rc;
}
