// binary-spgemm_b200/csrc/rows_sort.cuh — rows of the M bins (cap_s < IP <= 16384): one CTA per row, the row's candidate
// columns are gathered into shared memory and sorted by a CTA-wide bitonic network held in REGISTERS.
//
// Replaces, for these rows, the flag probes + quickSort of SpGEMM_bigslice (final/SpGEMM_mpi_omp.c:33-47).  The network is
// the one of the warp kernels (bitonic_sort_rows, kernels.cuh) extended over the CTA: element i = t*K + k lives in register
// k of thread t; exchanges at distance < K are register-to-register, < 32*K one SHFL, beyond that the partner's keys come
// through the shared-memory buffer the candidates were gathered into (conflict-free: word k*T + t).  The cost of a row is a
// function of its padded size only — power-law rows (all candidates on a few hub columns) and uniform rows cost the same,
// which the order-preserving slot map of the previous M-bin kernel could not offer (35 s at BASELINE config 4) — and, unlike
// a bitmap over the column range, it does not grow with Bm.  CTAs are small (256 / 512 threads, 8 / 64 KB of shared memory),
// so several rows are in flight on every SM and the dependent loads of one row's gather hide behind the sort of another.
#pragma once
#include "kernels.cuh"

namespace bsk {

constexpr u32 RS_LONG = 512;       // B rows at least this long are gathered by the whole CTA
constexpr int RS_QCAP = 32;        // ... up to this many per output row (the rest by the warp that met them)

// Ascending bitonic sort of K*T keys, K per thread (element index t*K + k).  buf: K*T words of shared memory.
// The levels are template recursion, not loops: with 105 exchange steps (16384 keys) the compiler no longer unrolls a
// loop nest completely, and one rolled loop is enough to push the key array into local memory.
template <int K, int T, int D>
__device__ __forceinline__ void cta_half_cleaners(u32 (&x)[K], const u32 t, u32* buf) {     // i <-> i ^ D, then D/2, ..., 1
  if constexpr (D >= 1) {
    if constexpr (D >= K) {                                        // partner key lives in thread ^ (D/K)
      constexpr u32 td = (u32)(D / K);
      const bool keepmin = (t & td) == 0u;
      if constexpr (D / K < 32) {
#pragma unroll
        for (int k = 0; k < K; ++k) { const u32 y = __shfl_xor_sync(0xffffffffu, x[k], td); x[k] = keepmin ? min(x[k], y) : max(x[k], y); }
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) buf[k * T + t] = x[k];
        __syncthreads();
        const u32 pt = t ^ td;
#pragma unroll
        for (int k = 0; k < K; ++k) { const u32 y = buf[k * T + pt]; x[k] = keepmin ? min(x[k], y) : max(x[k], y); }
        __syncthreads();
      }
    } else {                                                       // both keys in this thread
#pragma unroll
      for (int k = 0; k < K; ++k)
        if ((k & D) == 0) { const u32 lo = min(x[k], x[k | D]), hi = max(x[k], x[k | D]); x[k] = lo; x[k | D] = hi; }
    }
    cta_half_cleaners<K, T, D / 2>(x, t, buf);
  }
}

template <int K, int T, int SIZE>
__device__ __forceinline__ void cta_merge_levels(u32 (&x)[K], const u32 t, u32* buf) {       // merge levels SIZE, 2*SIZE, ..., K*T
  if constexpr (SIZE <= K * T) {
    if constexpr (SIZE <= K) {                                     // mirror inside the thread
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int pk = k ^ (SIZE - 1);
        if (k < pk) { const u32 lo = min(x[k], x[pk]), hi = max(x[k], x[pk]); x[k] = lo; x[pk] = hi; }
      }
    } else {                                                       // mirror across threads: register k <-> K-1-k of thread ^ (SIZE/K-1)
      constexpr u32 tm = (u32)(SIZE / K - 1);
      const bool keepmin = (t & (u32)(SIZE / (2 * K))) == 0u;
      if constexpr (SIZE / K <= 32) {
#pragma unroll
        for (int k = 0; k < K / 2; ++k) {
          const u32 ya = __shfl_xor_sync(0xffffffffu, x[K - 1 - k], tm);
          const u32 yb = __shfl_xor_sync(0xffffffffu, x[k], tm);
          x[k] = keepmin ? min(x[k], ya) : max(x[k], ya);
          x[K - 1 - k] = keepmin ? min(x[K - 1 - k], yb) : max(x[K - 1 - k], yb);
        }
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) buf[k * T + t] = x[k];
        __syncthreads();
        const u32 pt = t ^ tm;
#pragma unroll
        for (int k = 0; k < K; ++k) { const u32 y = buf[(K - 1 - k) * T + pt]; x[k] = keepmin ? min(x[k], y) : max(x[k], y); }
        __syncthreads();
      }
    }
    cta_half_cleaners<K, T, SIZE / 4>(x, t, buf);
    cta_merge_levels<K, T, SIZE * 2>(x, t, buf);
  }
}

template <int K, int T>
__device__ __forceinline__ void cta_bitonic_sort(u32 (&x)[K], const u32 t, u32* buf) { cta_merge_levels<K, T, 2>(x, t, buf); }

// Sort + de-duplicate the ipr candidates in buf[0..ipr) (ipr <= K*T).  Returns the number of distinct columns; MODE_FILL:
// they are written ascending to dst[0..count).  All threads of the CTA call it; buf is clobbered.
template <int K, int T, int MODE>
__device__ __forceinline__ u32 cta_sort_dedup(u32* buf, const u32 ipr, const u32 Bm, int* __restrict__ dst, u32* s_red, u32* s_last, u32* bad_out) {
  const u32 t = threadIdx.x, lane = t & 31u, warp = t >> 5;
  u32 x[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { const u32 i = (u32)k * T + t; x[k] = (i < ipr) ? buf[i] : EMPTY; }   // any order will do
  __syncthreads();
  cta_bitonic_sort<K, T>(x, t, buf);
  if (lane == 31) s_last[warp] = x[K - 1];
  __syncthreads();
  u32 prev = __shfl_up_sync(0xffffffffu, x[K - 1], 1);
  if (lane == 0) prev = warp ? s_last[warp - 1] : EMPTY;
  u32 c = 0, f = 0, bad = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const bool fk = (x[k] != EMPTY) && (x[k] != (k ? x[k - 1] : prev));
    f |= fk ? (1u << k) : 0u; c += fk ? 1u : 0u;
    bad |= (x[k] != EMPTY && x[k] >= Bm) ? 1u : 0u;
  }
  if (bad) *bad_out = 1;
  u32 tot;
  const u32 off = block_excl_scan(c, s_red, &tot);
  if (MODE != MODE_COUNT) {
    u32 o = off;
#pragma unroll
    for (int k = 0; k < K; ++k) if ((f >> k) & 1u) buf[o++] = x[k];
    __syncthreads();
    for (u32 i = t; i < tot; i += T) dst[i] = (int)buf[i];
    __syncthreads();
  }
  return tot;
}

// T threads per CTA, rows of up to KMAX*T candidates; the network is instantiated for KMAX, KMAX/2 and KMAX/4 keys per thread
// and picked per row.  Rows come from a list (k_build_lists) through an atomic counter.
template <int KMAX, int T, int MODE>
__global__ void __maxnreg__(64) k_rows_sort(Csr m, const u32* __restrict__ list, const u32* __restrict__ nlist, u32* __restrict__ ctr,
                                                 const u32* __restrict__ ip, u32* __restrict__ cnt, int G,
                                                 const void* __restrict__ Crow, int is64, int* __restrict__ Ccol,
                                                 const u64* __restrict__ tofs, DevScalars* sc) {
  extern __shared__ __align__(16) u32 buf[];                 // KMAX*T words
  __shared__ u32 s_red[33], s_last[32];
  __shared__ u32 s_idx[2], s_pos, s_nlong, s_bad;
  __shared__ uint4 s_long[RS_QCAP];                          // (start in Bcol, length, position in buf)
  const u32 t = threadIdx.x, lane = t & 31u, warp = t >> 5;
  constexpr u32 NW = T / 32;
  const u32 n = *nlist;
  const u32 SPW = 32u / (u32)G, sub = lane / (u32)G, off0 = lane % (u32)G;
  if (t == 0) { s_bad = 0; s_idx[0] = atomicAdd(ctr, 1u); }
  __syncthreads();
  // The row after this one is fetched while this one is sorted: its list index (an atomic) at the top of the iteration, its list
  // entry after the gather, its product count and row pointers after the sort — three of the six dependent global latencies
  // (counter -> list -> Arow -> Acol -> Brow -> Bcol) that otherwise stand in front of every row's gather.
  u32 idx = s_idx[0], ipr = 0, it = 0;
  int row = 0, a0 = 0, a1 = 0;
  if (idx < n) { row = (int)list[idx]; ipr = ip[row]; a0 = m.Arow[row]; a1 = m.Arow[row + 1]; }
  while (idx < n) {
    u32 nidx_reg = 0;
    if (t == 0) { nidx_reg = atomicAdd(ctr, 1u); s_pos = 0; s_nlong = 0; }
    __syncthreads();
    // ---- gather: every warp takes batches of 32 A nonzeros; G lanes walk one B row; positions from a warp scan + one
    //      shared-memory atomic per batch (the order of the candidates in buf does not matter)
    for (int b0 = a0 + (int)warp * 32; b0 < a1; b0 += (int)NW * 32) {
      const int jj = b0 + (int)lane;
      u32 bs = 0, len = 0;
      if (jj < a1) {
        const int j = m.Acol[jj];
        if ((u32)j < (u32)m.Bn) { bs = (u32)m.Brow[j]; len = (u32)m.Brow[j + 1] - bs; }
      }
      const u32 incl = warp_incl_scan(len);
      const u32 tot = __shfl_sync(0xffffffffu, incl, 31);
      u32 base = 0;
      if (lane == 0 && tot) base = atomicAdd(&s_pos, tot);
      base = __shfl_sync(0xffffffffu, base, 0);
      u32 pos = base + incl - len;
      if (len >= RS_LONG) {                                  // hand a long B row to the whole CTA
        const u32 q = atomicAdd(&s_nlong, 1u);
        if (q < (u32)RS_QCAP) { s_long[q] = make_uint4(bs, len, pos, 0u); len = 0; }
      }
      const int nseg = min(32, a1 - b0);
      for (int s = 0; s < nseg; s += (int)SPW) {
        const int src = (s + (int)sub) & 31;                 // lanes >= nseg carry len 0
        const u32 sbs = __shfl_sync(0xffffffffu, bs, src);
        const u32 slen = __shfl_sync(0xffffffffu, len, src);
        const u32 spos = __shfl_sync(0xffffffffu, pos, src);
#pragma unroll 4
        for (u32 o = off0; o < slen; o += (u32)G) buf[spos + o] = (u32)__ldg(&m.Bcol[sbs + o]);
      }
    }
    if (t == 0) s_idx[(it + 1u) & 1u] = nidx_reg;
    __syncthreads();
    const u32 nidx = s_idx[(it + 1u) & 1u];
    int nrow = 0;
    if (nidx < n) nrow = (int)list[nidx];
    {
      const u32 nl = min(s_nlong, (u32)RS_QCAP);
      for (u32 q = 0; q < nl; ++q) {
        const uint4 d = s_long[q];
#pragma unroll 4
        for (u32 o = t; o < d.y; o += T) buf[d.z + o] = (u32)__ldg(&m.Bcol[d.x + o]);
      }
    }
    __syncthreads();
    // ---- sort, de-duplicate, count / write
    int* dst = (MODE == MODE_FILL) ? Ccol + ld_rowptr(Crow, is64, (size_t)row) : (MODE == MODE_STAGE) ? Ccol + tofs[row] : nullptr;   // STAGE: Ccol is the staging arena
    u32 c;
    if (ipr <= (u32)(KMAX / 4) * T)      c = cta_sort_dedup<KMAX / 4, T, MODE>(buf, ipr, (u32)m.Bm, dst, s_red, s_last, &s_bad);
    else if (ipr <= (u32)(KMAX / 2) * T) c = cta_sort_dedup<KMAX / 2, T, MODE>(buf, ipr, (u32)m.Bm, dst, s_red, s_last, &s_bad);
    else                                 c = cta_sort_dedup<KMAX, T, MODE>(buf, ipr, (u32)m.Bm, dst, s_red, s_last, &s_bad);
    if (MODE != MODE_FILL && t == 0) cnt[row] = c;
    u32 nipr = 0; int na0 = 0, na1 = 0;
    if (nidx < n) { nipr = ip[nrow]; na0 = m.Arow[nrow]; na1 = m.Arow[nrow + 1]; }
    idx = nidx; row = nrow; ipr = nipr; a0 = na0; a1 = na1; ++it;
    __syncthreads();                                   // (s_pos / s_nlong are reset at the top of the next iteration)
  }
  __syncthreads();
  if (t == 0 && s_bad) atomicOr(&sc->err, 4u);
}

}  // namespace bsk
