// binary-spgemm_b200/csrc/fused_ell.cuh — the fast one-pass kernel for matrices whose B rows are short.
//
// Replaces the same reference code as kernels.cuh (SpGEMM_bigslice final/SpGEMM_mpi_omp.c:15-58 + the
// concatenation / row-pointer fix-up of SpGEMM_omp :111-141), for the case max_len(B row) <= 32.
//
// Why a second layout: the v3 kernel (kernels.cuh k_fused) spent 460 L1/shared-memory wavefronts and 2050
// warp instructions per output row at config 3 (profiles/r01_v3_fused_ncu_raw.csv) — descriptor gathers
// (Brow[j], Brow[j+1]: one 128-byte line per lane), unaligned 64-byte B-row reads that straddle 3 sectors,
// bank-conflicted staging stores.  Here B is first re-laid out (k_build_ell, one streaming pass, inside the
// timed region) as ELL: W = 4/8/16/32 columns per row, padded with EMPTY, every row 16·LPR-byte aligned, so
//   * the address of a B row is j*W: no descriptor gather at all;
//   * one LDG.128 per lane fetches 4 columns, a warp instruction fetches 32/LPR whole B rows (LPR = W/4),
//     each row exactly one or two 32-byte sectors... no straddling;
//   * all loads of a tile (R consecutive output rows, <= 64 A nonzeros per chunk) are issued back to back.
// De-duplication + sorting stay the ordered open-addressing table of kernels.cuh (monotone slot map, atomicMin,
// "the larger key moves right"), but: first probes are issued 4 per lane with no dependent branch, losers go to
// a warp queue (ballot-ranked, no atomics) and are drained with all lanes busy; the table is compacted IN PLACE
// by ballot/popc (conflict-free LDS/STS), the tile enters the decoupled look-back chain, and the rows are copied
// to their final position in Ccol with coalesced stores.  B is gathered once, C is written once.
#pragma once
#include "kernels.cuh"

namespace bsk {

constexpr int ELL_MAX_WARPS = 24;        // warps per CTA (one persistent CTA per SM)
constexpr int ELL_QCAP = 192;            // loser-queue entries per warp (a batch adds at most 128)

struct EllArgs {
  const int* __restrict__ Arow;   // An+1 absolute offsets
  const int* __restrict__ Acol;
  const u32* __restrict__ Bell;   // Bn rows of W columns, EMPTY-padded
  int An, Bn;
  u32 unit;                       // floor(2*W*2^32 / Bm): slot scale of a row with one A nonzero
  u32 TW;                         // table words per output row (multiple of 32)
  u32 Bm;
  void* Crow; int is64;
  int* Ccol;
  u64* status;
  DevScalars* sc;
  u32 ntiles;
};

__host__ __device__ constexpr u32 ell_table_limit(u32 lenA, u32 W) { return ((2u * W * lenA + 31u) & ~31u) + 32u; }
__host__ __device__ constexpr u32 ell_warp_words(u32 R, u32 TW) { return R * TW + 2u * ELL_QCAP + 2u * 8u; }

// ---- B (CSR) -> ELL.  LPR = W/4 lanes write one row as uint4 each; also validates B's columns.
template <int W>
__global__ void __launch_bounds__(256) k_build_ell(const int* __restrict__ Brow, const int* __restrict__ Bcol, int Bn, u32 Bm,
                                                   u32* __restrict__ Bell, DevScalars* sc) {
  constexpr int LPR = W / 4;
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = gtid / LPR;
  const int part = (int)(gtid % LPR);
  if (row >= Bn) return;
  const int bs = Brow[row], be = Brow[row + 1];
  uint4 v;
  u32 x[4];
  u32 bad = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int o = bs + part * 4 + k;
    x[k] = (o < be) ? (u32)__ldg(&Bcol[o]) : EMPTY;
    if (o < be && x[k] >= Bm) { bad = 1; x[k] = EMPTY; }
  }
  v.x = x[0]; v.y = x[1]; v.z = x[2]; v.w = x[3];
  reinterpret_cast<uint4*>(Bell)[row * LPR + part] = v;
  if (bad) atomicOr(&sc->err, 4u);
}

// ---- rare path: a row whose optimistic table spilled past its 32 spare slots is rebuilt with TW-cap home
// slots and cap spill slots (cap = lenA*W >= IP): a key is pushed right past at most IP-1 smaller keys, so
// this cannot overflow.  Returns nothing; the table ends ordered like the fast path's.
template <int W>
__device__ __noinline__ void ell_rebuild_row(const int* __restrict__ Acol, const u32* __restrict__ Bell, u32 Bn, u32 Bm, u32 TW,
                                              u32* tabr, int a0, int a1) {
  const u32 lane = lane_id();
  const u32 cap = (u32)(a1 - a0) * W;
  const u32 T = TW - cap;                                         // >= cap + 32 by construction of TW
  const u32 scale = (u32)min((u64)0xffffffffull, (((u64)T) << 32) / Bm);
  for (u32 q = lane * 4; q < TW; q += 128) *reinterpret_cast<uint4*>(tabr + q) = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
  __syncwarp();
  for (u32 idx = lane; idx < cap; idx += 32) {
    const int j = Acol[a0 + (int)(idx / W)];
    if ((u32)j >= Bn) continue;
    u32 x = Bell[(size_t)j * W + (idx % W)];
    if (x == EMPTY) continue;
    u32 s = __umulhi(x, scale);
    while (s < TW) {
      const u32 old = atomicMin(&tabr[s], x);
      if (old == EMPTY || old == x) break;
      x = max(old, x); ++s;
    }
  }
  __syncwarp();
}

// ---- re-insert the queued losers (key, next slot | spill limit << 16; both are word indices into the warp's
// region, < 2^16), all lanes busy.  Returns the rows (bit mask) in which a key ran past the row's spill limit.  Kept out of line: it is called from every unrolled batch position.
__device__ __noinline__ u32 ell_drain(u32* tab, const uint2* queue, u32 qn, u32 TW) {
  u32 ovf = 0;
  __syncwarp();
  for (u32 i = lane_id(); i < qn; i += 32) {
    const uint2 ent = queue[i];
    u32 x = ent.x, s = ent.y & 0xffffu;
    const u32 l = ent.y >> 16;
    while (true) {
      if (s >= l) { ovf |= 1u << ((l - 1u) / TW); break; }      // limit of row r = r*TW + lim_r, lim_r <= TW
      const u32 old = atomicMin(&tab[s], x);
      if (old == EMPTY || old == x) break;
      x = max(old, x); ++s;
    }
  }
  __syncwarp();
  return ovf;
}

// ---- the fused kernel.  Every warp is an independent worker on tiles of R consecutive rows.
template <int W, int R>
__global__ void __launch_bounds__(ELL_MAX_WARPS * 32, 1) k_fused_ell(const EllArgs p) {
  constexpr int LPR = W / 4;               // lanes per B row
  constexpr int NSEG = 32 / LPR;           // B rows per LDG.128 warp instruction ("batch")
  constexpr int NBS = 64 / NSEG;           // batches per chunk of 64 A nonzeros
  constexpr int NBG = NBS < 8 ? NBS : 8;   // batches in flight
  extern __shared__ __align__(16) u32 smem[];
  const u32 warp = threadIdx.x >> 5, lane = lane_id();
  const u32 TW = p.TW;
  u32* tab = smem + (size_t)warp * ell_warp_words(R, TW);
  uint2* queue = reinterpret_cast<uint2*>(tab + R * TW);
  uint2* par = queue + ELL_QCAP;            // per row: (slot scale, spill limit as word index into tab)
  const u32 sub = lane / LPR, part = lane % LPR;
  const u32 ltmask = (1u << lane) - 1u;
  const uint4* __restrict__ Bell4 = reinterpret_cast<const uint4*>(p.Bell);
  u32 ipc = 0;                              // intermediate products seen by this lane

  u32 tile = 0;
  if (lane == 0) tile = atomicAdd(&p.sc->tile_counter, 1u);
  tile = __shfl_sync(0xffffffffu, tile, 0);
  while (tile < p.ntiles) {
    u32 next = 0;
    if (lane == 0) next = atomicAdd(&p.sc->tile_counter, 1u);     // consumed at the end of this tile
    const long long row0 = (long long)tile * R;
    const int nrows = (int)min((long long)R, (long long)p.An - row0);
    const int ar = p.Arow[row0 + min((int)lane, nrows)];
    int a[R + 1];
#pragma unroll
    for (int r = 0; r <= R; ++r) a[r] = __shfl_sync(0xffffffffu, ar, r);
    const int E = a[R] - a[0];
    u32 lim[R];
#pragma unroll
    for (int r = 0; r < R; ++r) lim[r] = ell_table_limit((u32)(a[r + 1] - a[r]), W);
    if (lane < R) {
      u32 len = 0, l = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) if ((int)lane == r) { len = (u32)(a[r + 1] - a[r]); l = lim[r]; }
      par[lane] = make_uint2(len * p.unit, lane * TW + l);
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (a[r + 1] > a[r])
        for (u32 q = lane * 4; q < lim[r]; q += 128) *reinterpret_cast<uint4*>(tab + r * TW + q) = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
    __syncwarp();

    u32 ovf = 0, qn = 0;
    for (int e0 = 0; e0 < E; e0 += 64) {
      int j0 = -1, j1 = -1;
      if (e0 + (int)lane < E) j0 = p.Acol[a[0] + e0 + (int)lane];
      if (e0 + 32 + (int)lane < E) j1 = p.Acol[a[0] + e0 + 32 + (int)lane];
      if ((j0 >= p.Bn) | (j1 >= p.Bn) | (j0 < -1) | (j1 < -1)) {
        atomicOr(&p.sc->err, 1u);
        if ((u32)j0 >= (u32)p.Bn) j0 = -1;
        if ((u32)j1 >= (u32)p.Bn) j1 = -1;
      }
#pragma unroll
      for (int g = 0; g < NBS; g += NBG) {
        if (e0 + g * NSEG >= E) break;
        uint4 v[NBG];
#pragma unroll
        for (int u = 0; u < NBG; ++u) {
          const int seg = (g + u) * NSEG + (int)sub;
          const int j = __shfl_sync(0xffffffffu, ((g + u) * NSEG < 32) ? j0 : j1, seg & 31);
          v[u] = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
          if (j >= 0) v[u] = __ldg(&Bell4[(size_t)j * LPR + part]);
        }
#pragma unroll
        for (int u = 0; u < NBG; ++u) {
          if (e0 + (g + u) * NSEG >= E) break;
          const int e = e0 + (g + u) * NSEG + (int)sub;
          u32 r = 0;
#pragma unroll
          for (int q = 1; q < R; ++q) r += (e >= a[q] - a[0]) ? 1u : 0u;
          const uint2 pr = par[r];
          const u32 tabr = r * TW;
          const u32 x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          u32 s[4], old[4];
#pragma unroll
          // padding (EMPTY) takes a harmless atomicMin(.., EMPTY) on a private bank: no branch around the atomics
          for (int k = 0; k < 4; ++k) s[k] = tabr + ((x[k] != EMPTY) ? __umulhi(x[k], pr.x) : lane);
#pragma unroll
          for (int k = 0; k < 4; ++k) { old[k] = atomicMin(&tab[s[k]], x[k]); ipc += (x[k] != EMPTY) ? 1u : 0u; }
          const u32 hi = (pr.y << 16) + 1u;                  // queue entry: (key, next slot | spill limit << 16)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool lose = (old[k] != EMPTY) && (old[k] != x[k]) && (x[k] != EMPTY);
            const u32 m = __ballot_sync(0xffffffffu, lose);
            if (lose) queue[qn + __popc(m & ltmask)] = make_uint2(max(old[k], x[k]), s[k] + hi);
            qn += __popc(m);
          }
          if (qn > ELL_QCAP - 128) { ovf |= ell_drain(tab, queue, qn, TW); qn = 0; }
        }
      }
    }
    if (qn) ovf |= ell_drain(tab, queue, qn, TW);
    ovf = __reduce_or_sync(0xffffffffu, ovf);
    if (ovf) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if ((ovf >> r) & 1u) { ell_rebuild_row<W>(p.Acol, p.Bell, (u32)p.Bn, p.Bm, TW, tab + r * TW, a[r], a[r + 1]); lim[r] = TW; }
    }

    // ---- compact every table in place (ascending, duplicate-free), count
    u32 c[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      u32 base = 0;
      if (a[r + 1] > a[r]) {
        u32* t = tab + r * TW;
        for (u32 q = 0; q < lim[r]; q += 64) {
          const u32 v0 = t[q + lane];
          const u32 v1 = (q + 32 < lim[r]) ? t[q + 32 + lane] : EMPTY;
          const u32 m0 = __ballot_sync(0xffffffffu, v0 != EMPTY);
          const u32 m1 = __ballot_sync(0xffffffffu, v1 != EMPTY);
          const u32 n0 = __popc(m0);
          if (v0 != EMPTY) t[base + __popc(m0 & ltmask)] = v0;
          if (v1 != EMPTY) t[base + n0 + __popc(m1 & ltmask)] = v1;
          base += n0 + __popc(m1);
        }
      }
      c[r] = base;
    }
    __syncwarp();

    // ---- chain the tile into the scan, write row pointers, stream the rows out
    u32 agg = 0, mine = 0, incl_mine = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) { agg += c[r]; if ((int)lane == r) { mine = c[r]; incl_mine = agg; } }
    (void)mine;
    const u64 excl = lookback_exclusive(p.status, tile, (u64)agg);
    if ((int)lane < nrows) st_rowptr(p.Crow, p.is64, (size_t)(row0 + lane) + 1, excl + incl_mine, &p.sc->err);
    if (tile == 0 && lane == 0) st_rowptr(p.Crow, p.is64, 0, 0, &p.sc->err);
    if (tile == p.ntiles - 1 && lane == 0) p.sc->total_nnz = excl + agg;
    u32 off = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const u32* src = tab + r * TW;
      int* dst = p.Ccol + (excl + off);
      for (u32 q = lane; q < c[r]; q += 32) dst[q] = (int)src[q];
      off += c[r];
    }
    __syncwarp();
    tile = __shfl_sync(0xffffffffu, next, 0);
  }
  u64 ips = ipc;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ips += __shfl_xor_sync(0xffffffffu, ips, d);
  if (lane == 0 && ips) atomicAdd(&p.sc->total_ip, ips);
}

}  // namespace bsk
