// binary-spgemm_b200/csrc/band.cuh — one-pass kernel for BANDED / BLOCK-DIAGONAL matrices (BASELINE config 5).
//
// Replaces SpGEMM_bigslice (final/SpGEMM_mpi_omp.c:15-58) and the concatenation / fix-up of SpGEMM_omp (:111-141) for
// matrices whose B rows are runs of consecutive columns (first, first+1, ..., first+len-1) and whose output rows fit a
// window of BAND_BITS columns.  The reference walks every intermediate product of such a row (1024 flag probes for 63
// distinct columns at config 5) and then quick-sorts the row.  Here a B row is its descriptor (first, len): 8 bytes
// instead of 4*len, built once per call by k_build_desc (which also PROVES the run property for every B row), and an
// output row is the OR of its runs in a 128-bit register bitmap:
//   * one warp per output row for the gather: lane l takes A nonzero l (coalesced Acol, coalesced descriptor gather —
//     neighbouring A nonzeros of a banded row select neighbouring B rows), lo/hi by REDUX.MIN/MAX, every lane shifts its
//     run into 4 words, 4 x REDUX.OR give the row: no intermediate product is ever materialised, no atomics, no sort;
//   * lane r of the warp keeps row r: after 32 rows the lanes expand their bitmaps (ffs) into the CTA's staging buffer at
//     the row's offset (block scan of the popcounts), the tile enters the decoupled look-back chain (lookback_exclusive,
//     kernels.cuh) and the staged columns leave with coalesced stores.
// HBM traffic is the compulsory one: A once, 8 bytes per B row, C once (the algorithmic-bytes figure of SURVEY.md §8d
// counts every gathered B entry and is ~10x larger).
// The kernel is OPTIMISTIC: a B row that is not a run (k_build_desc) or an output row wider than BAND_BITS raises
// sc->band_fail and the host redoes the product with the general kernels.
#pragma once
#include "kernels.cuh"

namespace bsk {

constexpr u32 BAND_CONTIG = 0x80000000u;   // descriptor .y: bit 31 = "the row is the run first .. first+len-1", bits 30:0 = len
constexpr u32 BAND_BITS = 128;             // width of the register bitmap of one output row
constexpr int BAND_THREADS = 128;          // = rows per tile (4 warps x 32 rows)
constexpr u32 BAND_STAGE = 8192;           // staging words per CTA (tiles with more distinct columns are written directly)

// desc[j] = (first column of B row j, len | BAND_CONTIG).  One warp per 32 rows, lanes across the row's entries.
__global__ void __launch_bounds__(256) k_build_desc(const int* __restrict__ Brow, const int* __restrict__ Bcol, int Bn, u32 Bm,
                                                    uint2* __restrict__ desc, DevScalars* sc) {
  const u32 lane = lane_id();
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long r0 = w * 32;
  if (r0 >= Bn) return;
  const long long rmine = r0 + lane;
  int bs = 0, be = 0;
  if (rmine < Bn) { bs = Brow[rmine]; be = Brow[rmine + 1]; }
  u32 first = EMPTY;
  if (be > bs) first = (u32)__ldg(&Bcol[bs]);
  u32 ok_mine = 1;
  const int nr = (int)min(32ll, (long long)Bn - r0);
  constexpr int RB = 8;                               // rows whose loads are issued together
#pragma unroll 1
  for (int rb = 0; rb < nr; rb += RB) {
    int s[RB], len[RB];
    u32 v[RB];
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      s[q] = __shfl_sync(0xffffffffu, bs, (rb + q) & 31);
      len[q] = __shfl_sync(0xffffffffu, be, (rb + q) & 31) - s[q];
      if (rb + q >= nr) len[q] = 0;
    }
#pragma unroll
    for (int q = 0; q < RB; ++q) v[q] = ((int)lane < len[q]) ? (u32)__ldg(&Bcol[s[q] + (int)lane]) : 0u;
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const u32 f = __shfl_sync(0xffffffffu, first, (rb + q) & 31);
      u32 good = ((int)lane >= len[q] || v[q] == f + lane) ? 1u : 0u;
      for (int o = 32 + (int)lane; o < len[q]; o += 32) good &= ((u32)__ldg(&Bcol[s[q] + o]) == f + (u32)o) ? 1u : 0u;   // rows longer than a warp
      good = __all_sync(0xffffffffu, good);
      if ((int)lane == rb + q) ok_mine = good;
    }
  }
  const u32 len = (u32)(be - bs);
  u32 bad_col = 0;
  if (len && (first >= Bm || first + len - 1u >= Bm || first + len - 1u < first)) { bad_col = 1; ok_mine = 0; }
  if (rmine < Bn) desc[rmine] = make_uint2(first, len | (ok_mine ? BAND_CONTIG : 0u));
  const u32 fail = __any_sync(0xffffffffu, rmine < Bn && !ok_mine);
  const u32 badc = __any_sync(0xffffffffu, bad_col);
  if (lane == 0) { if (fail) atomicOr(&sc->band_fail, 1u); if (badc) atomicOr(&sc->band_fail, 2u); }   // bit 1: the general kernels will report the column
}

struct BandArgs {
  const int* __restrict__ Arow; const int* __restrict__ Acol; const uint2* __restrict__ desc;
  int An, Bn;
  void* Crow; int is64; int* Ccol;
  u64* status; DevScalars* sc; u32 ntiles;
};

// bits [s,e) of a 128-bit field, word w
__device__ __forceinline__ u32 run_word(u32 s, u32 e, u32 w) {
  const u32 a = max(s, 32u * w), b = min(e, 32u * w + 32u);
  if (a >= b) return 0u;
  const u32 n = b - a;
  return (n >= 32u ? 0xffffffffu : ((1u << n) - 1u)) << (a - 32u * w);
}

__global__ void __launch_bounds__(BAND_THREADS) k_band(const BandArgs p) {
  __shared__ __align__(16) u32 stage[BAND_STAGE];
  __shared__ u32 s_red[33];
  __shared__ u32 s_tile;
  __shared__ u64 s_excl;
  const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  if (tid == 0) s_tile = *(volatile u32*)&p.sc->band_fail;
  __syncthreads();
  if (s_tile) return;                                           // k_build_desc found a B row that is not a run
  u64 ips = 0;
  u32 fail = 0, bad_a = 0;
  while (true) {
    __syncthreads();
    if (tid == 0) s_tile = atomicAdd(&p.sc->tile_counter, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    if (tile >= p.ntiles) break;
    const long long row0 = (long long)tile * BAND_THREADS + warp * 32;
    const int ar = p.Arow[min(row0 + (long long)lane, (long long)p.An)];
    const int arE = p.Arow[min(row0 + 32ll, (long long)p.An)];
    u32 mylo = 0, m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    // blocks of RB rows: all Acol loads of the block are issued together, then all descriptor gathers, then the bitmaps
    // are built (two dependent round trips per RB rows instead of per row)
    constexpr int RB = 8;
#pragma unroll 1
    for (int rb = 0; rb < 32; rb += RB) {
      int a0[RB], la[RB], j[RB];
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        a0[q] = __shfl_sync(0xffffffffu, ar, rb + q);
        const int a1 = (rb + q < 31) ? __shfl_sync(0xffffffffu, ar, (rb + q + 1) & 31) : arE;
        la[q] = a1 - a0[q];
      }
#pragma unroll
      for (int q = 0; q < RB; ++q) j[q] = ((int)lane < la[q] && la[q] <= 32) ? p.Acol[a0[q] + (int)lane] : -1;
      uint2 dsc[RB];
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        dsc[q] = make_uint2(EMPTY, BAND_CONTIG);
        if (j[q] >= 0) { if (j[q] < p.Bn) dsc[q] = p.desc[j[q]]; else bad_a = 1; }
      }
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        const int r = rb + q;
        if (la[q] <= 0) continue;                                 // empty row (or past the end of A): uniform in the warp
        if (la[q] <= 32) {
          u32 f = dsc[q].x;
          const u32 l = dsc[q].y & ~BAND_CONTIG;
          if (!(dsc[q].y & BAND_CONTIG)) fail = 1;
          if (!l) f = EMPTY;
          ips += l;
          const u32 lo = __reduce_min_sync(0xffffffffu, f);
          const u32 hi = __reduce_max_sync(0xffffffffu, l ? f + l : 0u);    // one past the last column
          if (lo == EMPTY) continue;                              // only empty B rows
          if (hi - lo > BAND_BITS) { fail = 1; continue; }
          const u32 s = f - lo, e = s + l;
          const u32 w0 = __reduce_or_sync(0xffffffffu, l ? run_word(s, e, 0) : 0u);
          const u32 w1 = __reduce_or_sync(0xffffffffu, l ? run_word(s, e, 1) : 0u);
          const u32 w2 = __reduce_or_sync(0xffffffffu, l ? run_word(s, e, 2) : 0u);
          const u32 w3 = __reduce_or_sync(0xffffffffu, l ? run_word(s, e, 3) : 0u);
          if ((int)lane == r) { mylo = lo; m0 = w0; m1 = w1; m2 = w2; m3 = w3; }
        } else {                                                  // long A row: first the window, then the runs
          const int b0 = a0[q], b1 = a0[q] + la[q];
          u32 vlo = EMPTY, vhi = 0;
          for (int jj = b0 + (int)lane; jj < b1; jj += 32) {
            const int jx = p.Acol[jj];
            if ((u32)jx >= (u32)p.Bn) { bad_a = 1; continue; }
            const uint2 d = p.desc[jx];
            const u32 ll = d.y & ~BAND_CONTIG;
            if (!(d.y & BAND_CONTIG)) fail = 1;
            ips += ll;
            if (ll) { vlo = min(vlo, d.x); vhi = max(vhi, d.x + ll); }
          }
          const u32 lo = __reduce_min_sync(0xffffffffu, vlo);
          const u32 hi = __reduce_max_sync(0xffffffffu, vhi);
          if (lo == EMPTY) continue;
          if (hi - lo > BAND_BITS) { fail = 1; continue; }
          u32 w0 = 0, w1 = 0, w2 = 0, w3 = 0;
          for (int jj = b0 + (int)lane; jj < b1; jj += 32) {
            const int jx = p.Acol[jj];
            if ((u32)jx >= (u32)p.Bn) continue;
            const uint2 d = p.desc[jx];
            const u32 ll = d.y & ~BAND_CONTIG;
            if (!ll) continue;
            const u32 s = d.x - lo, e = s + ll;
            w0 |= run_word(s, e, 0); w1 |= run_word(s, e, 1); w2 |= run_word(s, e, 2); w3 |= run_word(s, e, 3);
          }
          w0 = __reduce_or_sync(0xffffffffu, w0); w1 = __reduce_or_sync(0xffffffffu, w1);
          w2 = __reduce_or_sync(0xffffffffu, w2); w3 = __reduce_or_sync(0xffffffffu, w3);
          if ((int)lane == r) { mylo = lo; m0 = w0; m1 = w1; m2 = w2; m3 = w3; }
        }
      }
    }
    // thread t holds row tile*128 + t: scan the counts, chain the tile, stage, stream out
    const u32 c = __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
    u32 agg;
    const u32 off = block_excl_scan(c, s_red, &agg);
    if (warp == 0) {
      const u64 ex = lookback_exclusive(p.status, tile, (u64)agg);
      if (lane == 0) s_excl = ex;
    }
    __syncthreads();
    const u64 excl = s_excl;
    const long long row = (long long)tile * BAND_THREADS + tid;
    if (row < p.An) st_rowptr(p.Crow, p.is64, (size_t)row + 1, excl + off + c, &p.sc->err);
    if (tile == 0 && tid == 0) st_rowptr(p.Crow, p.is64, 0, 0, &p.sc->err);
    if (tile == p.ntiles - 1 && tid == 0) p.sc->total_nnz = excl + agg;
    if (agg <= BAND_STAGE) {
      u32 o = off;
      u32 word = m0; while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; stage[o++] = mylo + b; }
      word = m1;     while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; stage[o++] = mylo + 32u + b; }
      word = m2;     while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; stage[o++] = mylo + 64u + b; }
      word = m3;     while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; stage[o++] = mylo + 96u + b; }
      __syncthreads();
      int* dst = p.Ccol + excl;
      for (u32 i = tid; i < agg; i += BAND_THREADS) dst[i] = (int)stage[i];
    } else {
      int* dst = p.Ccol + (excl + off);
      u32 word = m0; while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; *dst++ = (int)(mylo + b); }
      word = m1;     while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; *dst++ = (int)(mylo + 32u + b); }
      word = m2;     while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; *dst++ = (int)(mylo + 64u + b); }
      word = m3;     while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; *dst++ = (int)(mylo + 96u + b); }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ips += __shfl_xor_sync(0xffffffffu, ips, d);
  if (lane == 0 && ips) atomicAdd(&p.sc->total_ip, ips);
  if (__any_sync(0xffffffffu, fail) && lane == 0) atomicOr(&p.sc->band_fail, 1u);
  if (__any_sync(0xffffffffu, bad_a) && lane == 0) atomicOr(&p.sc->err, 1u);
}

}  // namespace bsk
