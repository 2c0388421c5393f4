/*
 * oracle/mpi_shm/mpi.h — multi-task MPI shim over fork() + shared memory (TEST INFRASTRUCTURE, not product code).
 *
 * SURVEY.md §8f N2.  The image has no MPI, and oracle/mpi_stub/mpi.h can only run the reference with one task.  The report's
 * fastest CPU mode is many tasks with few threads each (pure MPI beat pure OpenMP by 1.6x on 20 cores), so the CPU baseline
 * timed next to the GPU path should be able to use it.  This header implements the nine MPI calls the reference drivers use
 * (final/SpGEMM_mpi_omp.c:162-204, :297-364) for P tasks ON ONE HOST:
 *
 *   MPI_Init_thread   reads MPI_SHIM_TASKS (default 1), maps a control block and a bounce buffer MAP_SHARED|MAP_ANONYMOUS
 *                     (MPI_SHIM_MB megabytes of address space, default 65536, MAP_NORESERVE: only touched pages cost memory)
 *                     and fork()s P-1 children — the reference calls it first thing in main (:352), before it reads the matrix
 *                     or starts an OpenMP region, so every task then reads the file itself exactly like under mpirun (:309);
 *   MPI_Barrier       sense-reversing barrier on C11 atomics in the control block;
 *   MPI_Reduce        (MPI_INT, MPI_SUM to root), MPI_Gather, MPI_Gatherv: every task copies its contribution into the bounce
 *                     buffer at an offset derived from the per-task byte counts in the control block, barrier, the root copies
 *                     out (to displs[] for Gatherv), barrier.  Non-root receive arguments are ignored, as in MPI;
 *   MPI_Finalize      barrier; the root reaps the children.
 *
 * The reference sources compile UNMODIFIED against it (oracle/Makefile, target `ref`: _ref/SpGEMM_mpi_omp_shm and
 * _ref/SpGEMM_mpi_omp_validity_shm):   MPI_SHIM_TASKS=4 ./SpGEMM_mpi_omp_shm m.mtx <block> <threads> <times>
 */
#ifndef BSPGEMM_ORACLE_MPI_SHM_H
#define BSPGEMM_ORACLE_MPI_SHM_H

#include <stdatomic.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sched.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/types.h>
#include <sys/wait.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

#define MPI_COMM_WORLD      0
#define MPI_INT             4   /* value = sizeof(int): the shim uses the datatype as its byte width */
#define MPI_SUM             1
#define MPI_THREAD_FUNNELED 1
#define MPI_SUCCESS         0
#define MPI_SHIM_MAX_TASKS  256

struct mpi_shim_ctl {
  atomic_int arrived, sense;
  atomic_int failed;
  size_t bytes[MPI_SHIM_MAX_TASKS];      /* contribution of each task to the collective in flight */
};
static struct mpi_shim_ctl *mpi_shim_c;
static char *mpi_shim_buf;
static size_t mpi_shim_cap;
static int mpi_shim_rank, mpi_shim_size = 1, mpi_shim_local_sense;
static pid_t mpi_shim_kids[MPI_SHIM_MAX_TASKS];

static inline void mpi_shim_die(const char *msg)
{
  fprintf(stderr, "mpi_shm shim (task %d): %s\n", mpi_shim_rank, msg);
  if (mpi_shim_c) atomic_store(&mpi_shim_c->failed, 1);
  _exit(3);
}

static inline int MPI_Barrier(MPI_Comm c)
{
  (void)c;
  if (mpi_shim_size == 1) return MPI_SUCCESS;
  mpi_shim_local_sense ^= 1;
  if (atomic_fetch_add(&mpi_shim_c->arrived, 1) == mpi_shim_size - 1) {
    atomic_store(&mpi_shim_c->arrived, 0);
    atomic_store(&mpi_shim_c->sense, mpi_shim_local_sense);
  } else {
    unsigned spins = 0;
    while (atomic_load(&mpi_shim_c->sense) != mpi_shim_local_sense) {
      if (atomic_load(&mpi_shim_c->failed)) _exit(3);
      if (++spins > 2000) { sched_yield(); }
    }
  }
  return MPI_SUCCESS;
}

static inline int MPI_Init_thread(int *argc, void *argv, int required, int *provided)
{
  (void)argc; (void)argv;
  if (provided) *provided = required;
  const char *e = getenv("MPI_SHIM_TASKS");
  int P = e ? atoi(e) : 1;
  if (P < 1) P = 1;
  if (P > MPI_SHIM_MAX_TASKS) P = MPI_SHIM_MAX_TASKS;
  const char *m = getenv("MPI_SHIM_MB");
  mpi_shim_cap = (size_t)(m ? atoll(m) : 65536) << 20;
  mpi_shim_c = (struct mpi_shim_ctl *)mmap(NULL, sizeof *mpi_shim_c, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  for (;;) {                                      /* as much address space as the host hands out */
    mpi_shim_buf = (char *)mmap(NULL, mpi_shim_cap, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (mpi_shim_buf != MAP_FAILED || mpi_shim_cap <= ((size_t)64 << 20)) break;
    mpi_shim_cap >>= 1;
  }
  if (mpi_shim_c == MAP_FAILED || mpi_shim_buf == MAP_FAILED) { mpi_shim_c = NULL; mpi_shim_die("mmap of the shared buffers failed (lower MPI_SHIM_MB)"); }
  memset(mpi_shim_c, 0, sizeof *mpi_shim_c);
  mpi_shim_size = P;
  fflush(stdout); fflush(stderr);
  for (int r = 1; r < P; ++r) {
    pid_t k = fork();
    if (k < 0) mpi_shim_die("fork failed");
    if (k == 0) { mpi_shim_rank = r; break; }
    mpi_shim_kids[r] = k;
  }
  return MPI_SUCCESS;
}
static inline int MPI_Query_thread(int *provided) { if (provided) *provided = MPI_THREAD_FUNNELED; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = mpi_shim_size; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = mpi_shim_rank; return MPI_SUCCESS; }

static inline int MPI_Finalize(void)
{
  MPI_Barrier(MPI_COMM_WORLD);
  if (mpi_shim_rank == 0) {
    fflush(stdout);
    for (int r = 1; r < mpi_shim_size; ++r) { int st; waitpid(mpi_shim_kids[r], &st, 0); }
  }
  return MPI_SUCCESS;
}

/* every task publishes `nbytes` at sb; returns this task's offset in the bounce buffer (tasks laid out in rank order) */
static inline size_t mpi_shim_publish(const void *sb, size_t nbytes)
{
  mpi_shim_c->bytes[mpi_shim_rank] = nbytes;
  MPI_Barrier(MPI_COMM_WORLD);
  size_t off = 0, total = 0;
  for (int r = 0; r < mpi_shim_size; ++r) { if (r < mpi_shim_rank) off += mpi_shim_c->bytes[r]; total += mpi_shim_c->bytes[r]; }
  if (total > mpi_shim_cap) mpi_shim_die("collective larger than the bounce buffer (raise MPI_SHIM_MB)");
  if (nbytes) memcpy(mpi_shim_buf + off, sb, nbytes);
  MPI_Barrier(MPI_COMM_WORLD);
  return off;
}

static inline int MPI_Reduce(const void *sb, void *rb, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm c)
{
  (void)op; (void)c;
  if (mpi_shim_size == 1) { memmove(rb, sb, (size_t)count * (size_t)dt); return MPI_SUCCESS; }
  if (dt != MPI_INT) mpi_shim_die("MPI_Reduce: only MPI_INT / MPI_SUM is implemented");
  mpi_shim_publish(sb, (size_t)count * sizeof(int));
  if (mpi_shim_rank == root) {
    int *out = (int *)rb;
    const int *in = (const int *)mpi_shim_buf;
    for (int i = 0; i < count; ++i) { long long s = 0; for (int r = 0; r < mpi_shim_size; ++r) s += in[(size_t)r * count + i]; out[i] = (int)s; }
  }
  MPI_Barrier(MPI_COMM_WORLD);             /* the bounce buffer is free again */
  return MPI_SUCCESS;
}

static inline int MPI_Gather(const void *sb, int scount, MPI_Datatype sdt, void *rb, int rcount, MPI_Datatype rdt, int root, MPI_Comm c)
{
  (void)c;
  if (mpi_shim_size == 1) { memmove(rb, sb, (size_t)scount * (size_t)sdt); return MPI_SUCCESS; }
  mpi_shim_publish(sb, (size_t)scount * (size_t)sdt);
  if (mpi_shim_rank == root) {
    size_t off = 0;
    for (int r = 0; r < mpi_shim_size; ++r) {
      memcpy((char *)rb + (size_t)r * (size_t)rcount * (size_t)rdt, mpi_shim_buf + off, mpi_shim_c->bytes[r]);
      off += mpi_shim_c->bytes[r];
    }
  }
  MPI_Barrier(MPI_COMM_WORLD);
  return MPI_SUCCESS;
}

static inline int MPI_Gatherv(const void *sb, int scount, MPI_Datatype sdt, void *rb, const int *rcounts, const int *displs,
                              MPI_Datatype rdt, int root, MPI_Comm c)
{
  (void)c; (void)rcounts;
  if (mpi_shim_size == 1) { memmove((char *)rb + (size_t)displs[0] * (size_t)rdt, sb, (size_t)scount * (size_t)sdt); return MPI_SUCCESS; }
  mpi_shim_publish(sb, (size_t)scount * (size_t)sdt);
  if (mpi_shim_rank == root) {
    size_t off = 0;
    for (int r = 0; r < mpi_shim_size; ++r) {
      memcpy((char *)rb + (size_t)displs[r] * (size_t)rdt, mpi_shim_buf + off, mpi_shim_c->bytes[r]);
      off += mpi_shim_c->bytes[r];
    }
  }
  MPI_Barrier(MPI_COMM_WORLD);
  return MPI_SUCCESS;
}

#endif
