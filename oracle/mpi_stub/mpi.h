/*
 * oracle/mpi_stub/mpi.h — single-rank MPI shim (TEST INFRASTRUCTURE, not product code).
 *
 * The image has no MPI (no mpicc / mpirun / mpi.h).  The reference's hot path only needs MPI
 * for the outer gather in SpGEMM_mpi (final/SpGEMM_mpi_omp.c:155-225); with one rank every
 * collective degenerates to a local copy.  This header supplies exactly the symbols the two
 * reference drivers use so that the reference sources compile UNMODIFIED, in place, from
 * /root/reference/final (see oracle/Makefile, target `ref`).
 *
 * Symbols used by the reference (file:line in final/SpGEMM_mpi_omp.c):
 *   MPI_Init_thread :352   MPI_Query_thread :353   MPI_Comm_size :161   MPI_Comm_rank :162
 *   MPI_Barrier :319       MPI_Reduce :178         MPI_Gather :186,:204 MPI_Gatherv :203
 *   MPI_Finalize :364      MPI_COMM_WORLD, MPI_INT, MPI_SUM, MPI_THREAD_FUNNELED
 */
#ifndef BSPGEMM_ORACLE_MPI_STUB_H
#define BSPGEMM_ORACLE_MPI_STUB_H

#include <string.h>
#include <stddef.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

#define MPI_COMM_WORLD      0
#define MPI_INT             4   /* value = sizeof(int): the shim uses the datatype as its byte width */
#define MPI_SUM             1
#define MPI_THREAD_FUNNELED 1
#define MPI_SUCCESS         0

static inline int MPI_Init_thread(int *argc, void *argv, int required, int *provided)
{ (void)argc; (void)argv; if (provided) *provided = required; return MPI_SUCCESS; }
static inline int MPI_Query_thread(int *provided)
{ if (provided) *provided = MPI_THREAD_FUNNELED; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }

/* one rank: a reduction / gather to root is a copy of the root's own contribution */
static inline int MPI_Reduce(const void *sb, void *rb, int count, MPI_Datatype dt, MPI_Op op,
                             int root, MPI_Comm c)
{ (void)op; (void)root; (void)c; memmove(rb, sb, (size_t)count * (size_t)dt); return MPI_SUCCESS; }

static inline int MPI_Gather(const void *sb, int scount, MPI_Datatype sdt, void *rb, int rcount,
                             MPI_Datatype rdt, int root, MPI_Comm c)
{ (void)rcount; (void)rdt; (void)root; (void)c;
  memmove(rb, sb, (size_t)scount * (size_t)sdt); return MPI_SUCCESS; }

static inline int MPI_Gatherv(const void *sb, int scount, MPI_Datatype sdt, void *rb,
                              const int *rcounts, const int *displs, MPI_Datatype rdt,
                              int root, MPI_Comm c)
{ (void)rcounts; (void)root; (void)c;
  memmove((char *)rb + (size_t)displs[0] * (size_t)rdt, sb, (size_t)scount * (size_t)sdt);
  return MPI_SUCCESS; }

#endif
