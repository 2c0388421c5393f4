// binary-spgemm_b200/csrc/rows_window.cuh — rows with many intermediate products (M and L bins): one CTA per row,
// a bitmap over a WINDOW of the column range in shared memory, as many windows as the row's columns need.
//
// Replaces, for these rows, the flag array `xb[Bm]` + per-row quickSort of SpGEMM_bigslice (final/SpGEMM_mpi_omp.c:21,
// 33-47): a set bit is "column seen", reading the bitmap left to right is the sorted distinct row.  The cost of a row
// does not depend on how its columns are distributed (power-law matrices put most candidates of every row on the same
// few hub columns, which defeats order-preserving slot maps), only on IP and on the number of non-empty windows:
//   * window = `wwords` 32-bit words of shared memory (192 KB -> 1.5 M columns);
//   * the first window starts at the smallest first-entry of the selected B rows (exact minimum when the rows of B are
//     ascending; if a smaller column shows up during the walk, the walk restarts from it);
//   * every walk over the row's products sets the bits of the columns inside the window and keeps the smallest column
//     beyond it: that is where the next window starts (empty stretches of the column range cost nothing);
//   * COUNT adds the number of newly set bits; FILL scans the touched words (popc + block scan) and writes the columns
//     at Crow[row]; the touched words are cleared on the way.
// B rows of WIN_LONG entries or more are walked by a whole warp instead of a G-lane group (hub rows of R-MAT graphs).
// Rows are handed out by an atomic counter (costs differ by orders of magnitude).
#pragma once
#include "kernels.cuh"

namespace bsk {

constexpr u32 WIN_LONG = 1024;            // B rows at least this long: one warp per row in a second loop
constexpr u32 WIN_WORDS = 48u * 1024u;    // words per window (multiple of 4096): 192 KB of shared memory
constexpr u32 WIN_MAX_WINDOWS = 8;        // host: matrices with Bm beyond WIN_MAX_WINDOWS windows keep the table / global-bitmap kernels

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_rows_window(Csr m, const u32* __restrict__ list, const u32* __restrict__ nlist,
                                                         u32* __restrict__ ctr, u32* __restrict__ cnt, int G, u32 wwords,
                                                         const void* __restrict__ Crow, int is64, int* __restrict__ Ccol,
                                                         const u64* __restrict__ tofs, DevScalars* sc) {
  extern __shared__ __align__(16) u32 bm[];
  __shared__ u32 s_red[33];
  __shared__ u32 s_idx, s_guess, s_above, s_below, s_wlo, s_whi, s_bad, s_long;
  const u32 tid = threadIdx.x, nthr = blockDim.x;
  const u32 n = *nlist;
  const u32 wbits = wwords << 5;
  uint4* bm4 = reinterpret_cast<uint4*>(bm);
  for (u32 q = tid; q < wwords / 4; q += nthr) bm4[q] = make_uint4(0u, 0u, 0u, 0u);
  const int ngroups = (int)nthr / G, g = (int)tid / G, l = (int)tid % G;
  const int nwarps = (int)nthr >> 5, wid = (int)tid >> 5, lane = (int)lane_id();
  while (true) {
    __syncthreads();
    if (tid == 0) { s_idx = atomicAdd(ctr, 1u); s_guess = EMPTY; s_bad = 0; s_long = 0; }
    __syncthreads();
    const u32 idx = s_idx;
    if (idx >= n) break;
    const int row = (int)list[idx];
    const int a0 = m.Arow[row], a1 = m.Arow[row + 1];
    // where the first window starts: smallest first entry of the selected B rows
    {
      u32 gmin = EMPTY, any_long = 0;
      for (int jj = a0 + (int)tid; jj < a1; jj += (int)nthr) {
        const int j = m.Acol[jj];
        if ((u32)j >= (u32)m.Bn) continue;
        const int bs = m.Brow[j], be = m.Brow[j + 1];
        if (be > bs) gmin = min(gmin, (u32)__ldg(&m.Bcol[bs]));
        if ((u32)(be - bs) >= WIN_LONG) any_long = 1;
      }
      gmin = __reduce_min_sync(0xffffffffu, gmin);
      any_long = __any_sync(0xffffffffu, any_long);
      if (lane == 0) { if (gmin != EMPTY) atomicMin(&s_guess, gmin); if (any_long) s_long = 1; }
    }
    __syncthreads();
    const bool has_long = s_long != 0;
    u32 start = s_guess;
    u32 added = 0;
    u64 done = 0;
    const u64 base = (MODE == MODE_FILL) ? ld_rowptr(Crow, is64, (size_t)row) : (MODE == MODE_STAGE) ? tofs[row] : 0;   // STAGE: Ccol is the staging arena
    bool first = true;
    if (start != EMPTY && start < (u32)m.Bm) {
      start &= ~31u;
      while (true) {
        if (tid == 0) { s_above = EMPTY; s_below = EMPTY; s_wlo = EMPTY; s_whi = 0; }
        __syncthreads();
        const u32 end = start + wbits;                       // Bm <= 2^31 and wbits <= 2^22: no wrap
        u32 add = 0, above = EMPTY, below = EMPTY, wlo = EMPTY, whi = 0, bad = 0;
        auto ins = [&](u32 v) {
          const u32 d = v - start;                           // wraps to a huge value for v < start
          if (v >= (u32)m.Bm) bad = 1;
          else if (d < wbits) {
            const u32 w = d >> 5, bit = 1u << (d & 31);
            const u32 old = atomicOr(&bm[w], bit);
            add += (old & bit) ? 0u : 1u;
            wlo = min(wlo, w); whi = max(whi, w);
          } else if (v >= end) above = min(above, v);
          else below = min(below, v);
        };
        for (int jj = a0 + g; jj < a1; jj += ngroups) {      // G lanes per (short) B row
          const int j = m.Acol[jj];
          if ((u32)j >= (u32)m.Bn) continue;
          const int bs = m.Brow[j], be = m.Brow[j + 1];
          if (has_long && (u32)(be - bs) >= WIN_LONG) continue;
          for (int o = bs + l; o < be; o += G) ins((u32)__ldg(&m.Bcol[o]));
        }
        if (has_long)
          for (int jj = a0 + wid; jj < a1; jj += nwarps) {   // one warp per long B row
            const int j = m.Acol[jj];
            if ((u32)j >= (u32)m.Bn) continue;
            const int bs = m.Brow[j], be = m.Brow[j + 1];
            if ((u32)(be - bs) < WIN_LONG) continue;
            for (int o = bs + lane; o < be; o += 32) ins((u32)__ldg(&m.Bcol[o]));
          }
        above = __reduce_min_sync(0xffffffffu, above);
        below = __reduce_min_sync(0xffffffffu, below);
        wlo = __reduce_min_sync(0xffffffffu, wlo);
        whi = __reduce_max_sync(0xffffffffu, whi);
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) {
          if (above != EMPTY) atomicMin(&s_above, above);
          if (below != EMPTY) atomicMin(&s_below, below);
          if (wlo != EMPTY) { atomicMin(&s_wlo, wlo); atomicMax(&s_whi, whi); }
          if (bad) s_bad = 1;
        }
        __syncthreads();
        const u32 w_lo = s_wlo, w_hi = s_whi, nxt = s_above, blw = s_below;
        const bool restart = first && blw != EMPTY;          // B rows not ascending: a column below the guessed minimum
        first = false;
        if (!restart) added += add;
        if (w_lo != EMPTY) {
          const u32 q0 = w_lo >> 2, q1 = w_hi >> 2;
          if (MODE != MODE_COUNT && !restart) {
            // Emission: every warp owns a contiguous slice of the touched range [q0,q1] (in 16-byte units).  Pass 1 counts the
            // set bits of the slice, ONE block scan orders the 32 slices, pass 2 re-reads the slice in order (warp scan of the
            // lanes' popcounts), writes the columns and clears the words.  (One block scan per 1024 units — up to 12 per window
            // and row — made the kernel barrier-stall bound, profiles/r01_window_m2_*.)
            const u32 span = q1 - q0 + 1u, per = (span + (u32)nwarps - 1u) / (u32)nwarps;
            const u32 s0 = q0 + (u32)wid * per, s1 = min(q1 + 1u, s0 + per);
            u32 c = 0;
            for (u32 q = s0 + (u32)lane; q < s1; q += 32u) { const uint4 v = bm4[q]; c += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }
            c = __reduce_add_sync(0xffffffffu, c);
            u32 tot;
            u32 o = block_excl_scan(lane == 0 ? c : 0u, s_red, &tot);       // exclusive offset of the warp's slice (lane 0 carries the warp's count)
            o = __shfl_sync(0xffffffffu, o, 0);
            for (u32 qb = s0; qb < s1; qb += 32u) {
              const u32 q = qb + (u32)lane;
              uint4 v = make_uint4(0u, 0u, 0u, 0u);
              if (q < s1) { v = bm4[q]; bm4[q] = make_uint4(0u, 0u, 0u, 0u); }
              const u32 cl = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
              const u32 incl = warp_incl_scan(cl);
              int* dst = Ccol + (base + done + o + incl - cl);
              const u32 col0 = start + (q << 7);
              u32 word = v.x; while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; *dst++ = (int)(col0 + b); }
              word = v.y;     while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; *dst++ = (int)(col0 + 32u + b); }
              word = v.z;     while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; *dst++ = (int)(col0 + 64u + b); }
              word = v.w;     while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; *dst++ = (int)(col0 + 96u + b); }
              o += __shfl_sync(0xffffffffu, incl, 31);
            }
            done += tot;
          } else {
            for (u32 q = q0 + tid; q <= q1; q += nthr) bm4[q] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        __syncthreads();
        if (restart) { start = blw & ~31u; continue; }
        if (nxt == EMPTY) break;
        start = nxt & ~31u;
      }
    } else if (start != EMPTY) {
      if (tid == 0) s_bad = 1;                               // first entry of a B row outside [0,Bm)
    }
    if (MODE != MODE_FILL) {
      const u32 c = block_reduce_add(added, s_red);
      if (tid == 0) cnt[row] = c;
    }
    __syncthreads();
    if (tid == 0 && s_bad) atomicOr(&sc->err, 4u);
  }
}

}  // namespace bsk
