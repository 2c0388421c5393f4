// binary-spgemm_b200/csrc/rows_l2bm.cuh — big rows (M2 and L bins: more than 2048 intermediate products): one CTA per row, the
// row's columns are marked in a bitmap over ALL of [0,Bm) that lives in global memory but stays in L2 (one private region
// per CTA: 148 x Bm/8 bytes = 76 MB at Bm = 2^22), with a summary bitmap in shared memory (one bit per 32-bit bitmap word)
// that tells the emission which words to read.
//
// Replaces, for these rows, the flag array `xb[Bm]` + quickSort of SpGEMM_bigslice (final/SpGEMM_mpi_omp.c:21,33-47) — it IS
// the flag array, one bit per column, and reading it left to right is the sorted distinct row.  Why not the kernels of round 1:
//   * the CTA-wide bitonic sort (rows_sort.cuh) costs O(IP log^2 IP) with a shared-memory exchange + two barriers per far
//     stage: 174 ms for the 1.04 M rows of 2049..16384 products of R-MAT scale 22 (74 us per row and CTA);
//   * the windowed shared-memory bitmap (rows_window.cuh) walks all products of the row once per 1.5 M-column window (3 times
//     at Bm = 2^22) and pays a shared-memory atomic per product: 152 ms for the 132 K rows above 16384 products.
// Here every product costs one coalesced Bcol load and one L2 atomic (ATOMG.OR, four in flight per thread), the row is
// walked ONCE whatever its column range, and the emission touches only the words that were set (two passes over them: count,
// one block scan, write + clear), so the cost is linear in IP + nnz(row).
#pragma once
#include "kernels.cuh"

namespace bsk {

constexpr u32 L2B_MAX_BM = 1u << 24;      // 2 MB of bitmap per CTA, 64 KB of summary + counts in shared memory
constexpr u32 L2B_LONG = 1024;            // B rows at least this long: one warp per row in a second loop

__host__ __device__ constexpr u32 l2b_summary_words(u32 bm_words) { return (bm_words + 31u) >> 5; }

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_rows_l2bm(Csr m, const u32* __restrict__ list, const u32* __restrict__ nlist, u32* __restrict__ ctr,
                                                       u32* __restrict__ cnt, int G, u32* __restrict__ gbm_all, u32 bm_words,
                                                       const void* __restrict__ Crow, int is64, int* __restrict__ Ccol,
                                                       const u64* __restrict__ tofs, DevScalars* sc) {
  extern __shared__ __align__(16) u32 sm[];
  __shared__ u32 s_red[33];
  __shared__ u32 s_idx, s_bad, s_long;
  const u32 tid = threadIdx.x, nthr = blockDim.x, lane = lane_id(), wid = tid >> 5, nwarps = nthr >> 5;
  const u32 sw = l2b_summary_words(bm_words);
  u32* summary = sm;                       // sw words: bit b of word s = bitmap word 32 s + b is non-zero
  u32* counts = sm + sw;                   // sw words: columns under summary word s, then their exclusive prefix
  u32* gbm = gbm_all + (size_t)blockIdx.x * bm_words;    // all zero between rows (the emission clears what it reads)
  for (u32 s = tid; s < sw; s += nthr) summary[s] = 0;
  const u32 n = *nlist;
  const int ngroups = (int)nthr / G, g = (int)tid / G, l = (int)tid % G;
  while (true) {
    __syncthreads();
    if (tid == 0) { s_idx = atomicAdd(ctr, 1u); s_bad = 0; s_long = 0; }
    __syncthreads();
    const u32 idx = s_idx;
    if (idx >= n) break;
    const int row = (int)list[idx];
    const int a0 = m.Arow[row], a1 = m.Arow[row + 1];
    u32 added = 0, bad = 0, any_long = 0;
    // four products per step: the four ATOMG are in flight together, their results are consumed afterwards
    auto ins4 = [&](const u32 (&v)[4], const int nv) {
      u32 old[4], bit[4], w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        old[k] = 0xffffffffu; bit[k] = 0; w[k] = 0;
        if (k < nv) {
          if (v[k] >= (u32)m.Bm) bad = 1;
          else { w[k] = v[k] >> 5; bit[k] = 1u << (v[k] & 31u); old[k] = atomicOr(&gbm[w[k]], bit[k]); }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (bit[k] && !(old[k] & bit[k])) { ++added; if (old[k] == 0u) atomicOr(&summary[w[k] >> 5], 1u << (w[k] & 31u)); }
    };
    auto walk = [&](int bs, int be, int first, int step) {
      int o = bs + first;
      for (; o + 3 * step < be; o += 4 * step) {
        u32 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (u32)__ldg(&m.Bcol[o + k * step]);
        ins4(v, 4);
      }
      if (o < be) {
        u32 v[4] = {0u, 0u, 0u, 0u};
        int nv = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) if (o + k * step < be) { v[k] = (u32)__ldg(&m.Bcol[o + k * step]); nv = k + 1; }
        ins4(v, nv);
      }
    };
    for (int jj = a0 + g; jj < a1; jj += ngroups) {          // G lanes per (short) B row
      const int j = m.Acol[jj];
      if ((u32)j >= (u32)m.Bn) { bad = 1; continue; }
      const int bs = m.Brow[j], be = m.Brow[j + 1];
      if ((u32)(be - bs) >= L2B_LONG) { any_long = 1; continue; }
      walk(bs, be, l, G);
    }
    if (__any_sync(0xffffffffu, any_long) && lane == 0) s_long = 1;
    __syncthreads();
    if (s_long)
      for (int jj = a0 + (int)wid; jj < a1; jj += (int)nwarps) {   // one warp per long B row (hub rows of R-MAT graphs)
        const int j = m.Acol[jj];
        if ((u32)j >= (u32)m.Bn) continue;
        const int bs = m.Brow[j], be = m.Brow[j + 1];
        if ((u32)(be - bs) < L2B_LONG) continue;
        walk(bs, be, (int)lane, 32);
      }
    if (__any_sync(0xffffffffu, bad) && lane == 0) s_bad = 1;
    __threadfence_block();
    __syncthreads();                                          // every bit of the row is set (the ATOMG have returned)
    // ---- emission.  Pass 1: columns under every summary word (warp w takes summary words w, w + nwarps, ...; lane b reads
    //      bitmap word 32 s + b: one 128-byte line per step)
    for (u32 s = wid; s < sw; s += nwarps) {
      const u32 sumw = summary[s];
      u32 c = 0;
      if (sumw) {
        const u32 wd = ((sumw >> lane) & 1u) ? __ldcg(&gbm[32u * s + lane]) : 0u;
        c = __reduce_add_sync(0xffffffffu, (u32)__popc(wd));
      }
      if (lane == 0) counts[s] = c;
    }
    __syncthreads();
    // exclusive prefix of counts[]: every thread owns a contiguous run of `per` entries
    const u32 per = (sw + nthr - 1u) / nthr, e0 = min(sw, tid * per), e1 = min(sw, e0 + per);
    u32 mine = 0;
    for (u32 s = e0; s < e1; ++s) mine += counts[s];
    u32 total;
    u32 run = block_excl_scan(mine, s_red, &total);
    for (u32 s = e0; s < e1; ++s) { const u32 c = counts[s]; counts[s] = run; run += c; }
    __syncthreads();
    // Pass 2: write the columns in order, clear the bitmap words and the summary
    const u64 base = (MODE == MODE_FILL) ? ld_rowptr(Crow, is64, (size_t)row) : (MODE == MODE_STAGE) ? tofs[row] : 0;   // STAGE: Ccol is the staging arena
    for (u32 s = wid; s < sw; s += nwarps) {
      const u32 sumw = summary[s];
      if (!sumw) continue;
      u32 wd = 0;
      if ((sumw >> lane) & 1u) { wd = __ldcg(&gbm[32u * s + lane]); __stcg(&gbm[32u * s + lane], 0u); }
      if (MODE != MODE_COUNT) {
        const u32 c = (u32)__popc(wd);
        const u32 incl = warp_incl_scan(c);
        int* dst = Ccol + (base + counts[s] + incl - c);
        const u32 col0 = (32u * s + lane) << 5;
        while (wd) { const u32 b = __ffs(wd) - 1; wd &= wd - 1; *dst++ = (int)(col0 + b); }
      }
      __syncwarp();
      if (lane == 0) summary[s] = 0;
    }
    if (MODE != MODE_FILL && tid == 0) cnt[row] = total;
    (void)added;
    __syncthreads();
    if (tid == 0 && s_bad) atomicOr(&sc->err, 4u);
  }
}

}  // namespace bsk
