// binary-spgemm_b200/csrc/mask.cuh — masked boolean product  C = F .* (A·B)  (SURVEY.md §8f N4).
//
// Replaces SpGEMM_masked (final/SpGEMM_mpi_omp.c:232-288): the reference starts every flag at "seen", clears the flags of the
// mask row F_i (:254-256), runs Gustavson's loop — a product column is appended only where its flag is clear (:264-269) — sorts
// the row and restores the flags.  The result is row i of A·B intersected with the pattern of F_i, ascending and distinct.
// Here the unmasked row comes from the product kernels (sorted, distinct, in the handle's arena) and the mask is applied by
// a merge-free intersection: every column of the row is looked up in the (ascending) mask row by binary search, one warp per
// row, ballot-ranked so that the survivors keep their order; count -> device scan (k_scan) -> fill, like the two-phase
// pipeline.  A mask whose rows are not ascending / distinct is canonicalised first (by the product kernels themselves:
// F' = I·F), so any F the reference accepts is accepted.
#pragma once
#include "kernels.cuh"

namespace bsk {

// err bit 4: a mask row is not strictly ascending; bit 5: a mask column outside [0,Bm)
static __global__ void __launch_bounds__(256) k_mask_check(const int* __restrict__ Frow, const int* __restrict__ Fcol, int An, u32 Bm, DevScalars* sc) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= An) return;
  const u32 lane = lane_id();
  const int f0 = Frow[row], f1 = Frow[row + 1];
  u32 bad = 0;
  for (int p = f0 + (int)lane; p < f1; p += 32) {
    const u32 v = (u32)Fcol[p];
    if (v >= Bm) bad |= 32u;
    if (p > f0 && (u32)Fcol[p - 1] >= v) bad |= 16u;
  }
  bad = __reduce_or_sync(0xffffffffu, bad);
  if (lane == 0 && bad) atomicOr(&sc->err, bad);
}

static __global__ void __launch_bounds__(256) k_iota(int* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}
static __global__ void __launch_bounds__(256) k_narrow_rowptr(const long long* __restrict__ in, int* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)in[i];
}

// is column v in the ascending array f[0..n)?
__device__ __forceinline__ bool mask_has(const int* __restrict__ f, int n, u32 v) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if ((u32)__ldg(&f[mid]) < v) lo = mid + 1; else hi = mid; }
  return lo < n && (u32)__ldg(&f[lo]) == v;
}

// One warp per row of the unmasked product (Cin_row: 64-bit row pointers, Cin_col ascending): MODE_COUNT -> cnt[row];
// MODE_FILL -> the surviving columns at Cout_col[Cout_row[row] ..), in order.
template <int MODE>
static __global__ void __launch_bounds__(256) k_mask_rows(const int* __restrict__ Frow, const int* __restrict__ Fcol,
                                                          const long long* __restrict__ Cin_row, const int* __restrict__ Cin_col, int An,
                                                          u32* __restrict__ cnt, const void* __restrict__ Cout_row, int is64, int* __restrict__ Cout_col) {
  const u32 lane = lane_id();
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < An; row += nw) {
    const int f0 = Frow[row], nf = Frow[row + 1] - f0;
    const long long c0 = Cin_row[row], c1 = Cin_row[row + 1];
    u64 out = (MODE == MODE_FILL) ? ld_rowptr(Cout_row, is64, (size_t)row) : 0;
    u32 total = 0;
    if (nf > 0)
      for (long long p = c0; p < c1; p += 32) {
        const long long q = p + lane;
        u32 v = 0; bool keep = false;
        if (q < c1) { v = (u32)Cin_col[q]; keep = mask_has(Fcol + f0, nf, v); }
        const u32 m = __ballot_sync(0xffffffffu, keep);
        if (MODE == MODE_FILL && keep) Cout_col[out + __popc(m & ((1u << lane) - 1u))] = (int)v;
        out += __popc(m); total += __popc(m);
      }
    if (MODE == MODE_COUNT && lane == 0) cnt[row] = total;
  }
}

}  // namespace bsk
