// binary-spgemm_b200/csrc/fused_ell.cuh — the fast one-pass kernel for matrices whose B rows are short.
//
// Replaces the same reference code as kernels.cuh (SpGEMM_bigslice final/SpGEMM_mpi_omp.c:15-58 + the
// concatenation / row-pointer fix-up of SpGEMM_omp :111-141), for the case max_len(B row) <= 32.
//
// Why a second layout: the v3 kernel (kernels.cuh k_fused) spent 460 L1/shared-memory wavefronts and 2050
// warp instructions per output row at config 3 (profiles/r01_v3_fused_ncu_raw.csv) — descriptor gathers
// (Brow[j], Brow[j+1]: one 128-byte line per lane), unaligned 64-byte B-row reads that straddle 3 sectors,
// bank-conflicted staging stores.  Here B is first re-laid out (k_build_ell, one streaming pass, inside the
// timed region) as ELL: W = 4/8/16/32 columns per row, padded with EMPTY, every row 4W-byte aligned, so
//   * the address of a B row is j*W: no descriptor gather at all;
//   * one LDG.128 per lane fetches 4 columns, a warp instruction fetches 32/LPR whole B rows (LPR = W/4),
//     every row whole sectors, no straddling;
//   * all loads of a tile (R consecutive output rows, <= 64 A nonzeros per chunk) are issued back to back, one
//     tile AHEAD: tile ids, row pointers, Acol and the B rows of the next tile are in flight while the current
//     tile is inserted / compacted / written (4-deep software pipeline, registers only).
// De-duplication + sorting stay the ordered open-addressing table of kernels.cuh (monotone slot map, atomicMin,
// "the larger key moves right"), but: first probes are issued 4 per lane with no dependent branch, losers go to
// a warp queue (ballot-ranked, no atomics) and are drained with all lanes busy; the table is compacted IN PLACE
// by ballot/popc (conflict-free LDS/STS) and copied to its final position in Ccol with coalesced stores.
// B is gathered once, C is written once.
//
// The scan that gives every tile its output offset is a TWO-LEVEL decoupled look-back (TileChain): with one
// warp per tile ~350 tiles finish per microsecond at the target rate, far more than a flat 32-wide look-back
// window can retire per L2 round trip (profiles/r01_v4_ell_flat_lookback.txt: 76 % of all issued instructions
// were the spin).  Tiles post their aggregate (a) as a flagged word and (b) into a packed per-group counter
// (32 tiles per group, one 64-bit atomicAdd: count<<40 | sum); a tile's offset = exclusive prefix of its group
// (walked 32 groups = 1024 tiles per round trip, published once per group) + the flagged words before it in
// its own group.  The aggregate is known right after the inserts (it is the number of atomicMin that found an
// EMPTY slot) and is posted BEFORE the compaction, so predecessors have normally published by the time a
// warp needs its offset.
#pragma once
#include "kernels.cuh"

namespace bsk {

constexpr int ELL_MAX_WARPS = 24;        // warps per CTA (one persistent CTA per SM)
constexpr int ELL_QCAP = 192;            // loser-queue entries per warp (a batch adds at most 128)

struct TileChain {        // all zero before the launch
  u32* s0;                // [ntiles]  bit31 = posted, low bits = aggregate of the tile
  u64* gsum;              // [ngroups] (tiles posted << 40) | sum of their aggregates
  u64* ginc;              // [ngroups] bit63 = known, low bits = exclusive prefix of the group
};
__host__ __device__ inline size_t tile_chain_words64(size_t ntiles) { const size_t ng = (ntiles + 31) / 32; return 2 * ng + (ntiles + 1) / 2 + 2; }

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
  TileChain chain;
  DevScalars* sc;
  u32 ntiles;
  u32 debug_nochain;              // timing experiments only (BSPGEMM_DEBUG_NOCHAIN): skip the scan, rows land at upper-bound offsets
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

// ------------------------------------------------------------------------------------------------ two-level tile chain
__device__ __forceinline__ u32 ld_relaxed_u32(const u32* p) {
  u32 v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
constexpr u64 CH_KNOWN = 1ull << 63;
constexpr u64 CH_SUM = (1ull << 40) - 1;

// lane 0 publishes the tile's aggregate
__device__ __forceinline__ void chain_post(const TileChain& c, u32 tile, u32 agg) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(c.s0 + tile), "r"(0x80000000u | agg) : "memory");
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(c.gsum + (tile >> 5)), "l"((1ull << 40) + (u64)agg) : "memory");
}

// Exclusive prefix of `tile` (whole warp).  Every tile a running tile waits for belongs to a resident warp (see the
// tile assignment in k_fused_ell), and aggregates are posted before anybody waits, so the spins terminate.
__device__ __noinline__ u64 chain_exclusive(const TileChain c, u32 tile, u32 ntiles) {
  const u32 lane = lane_id();
  const u32 g = tile >> 5, k = tile & 31;
  u32 w = (lane < k) ? ld_relaxed_u32(c.s0 + (g << 5) + lane) : 0x80000000u;
  u64 e = (g == 0) ? CH_KNOWN : ld_status(c.ginc + g);
  while (__any_sync(0xffffffffu, !(w >> 31))) {
    __nanosleep(40);
    if (!(w >> 31)) w = ld_relaxed_u32(c.s0 + (g << 5) + lane);
  }
  u64 intra = __reduce_add_sync(0xffffffffu, w & 0x7fffffffu);
  if (e & CH_KNOWN) return (e & ~CH_KNOWN) + intra;
  u64 acc = 0;
  long long idx = (long long)g - 1;
  while (true) {
    const long long my = idx - lane;
    const u64 need = (my >= 0) ? (u64)min(32ll, (long long)ntiles - 32 * my) : 0ull;
    u64 gi = CH_KNOWN, gs = 0;
    u32 first;
    while (true) {
      if (my >= 0) { gi = ld_status(c.ginc + my); gs = ld_status(c.gsum + my); }
      const u32 known = __ballot_sync(0xffffffffu, (gi & CH_KNOWN) != 0);
      first = known ? (u32)(__ffs(known) - 1) : 32u;
      const bool pending = (lane <= first) && ((gs >> 40) != need);
      if (!__any_sync(0xffffffffu, pending)) break;
      __nanosleep(40);
    }
    u64 v = (lane <= first) ? (gs & CH_SUM) : 0ull;
    if (lane == first) v += gi & ~CH_KNOWN;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    acc += v;
    if (first < 32u) break;
    idx -= 32;
  }
  if (lane == 0) st_status(c.ginc + g, CH_KNOWN | acc);
  return acc + intra;
}

// ------------------------------------------------------------------------------------------------ slow paths (out of line)
// Re-insert the queued losers (key, next slot | spill limit << 16; both are word indices into the warp's region,
// < 2^16), all lanes busy.  Returns (rows whose table spilled past its limit) << 16 | (#keys that found an EMPTY slot).
__device__ __noinline__ u32 ell_drain(u32* tab, const uint2* queue, u32 qn, u32 TW) {
  u32 ovf = 0, added = 0;
  __syncwarp();
  for (u32 i = lane_id(); i < qn; i += 32) {
    const uint2 ent = queue[i];
    u32 x = ent.x, s = ent.y & 0xffffu;
    const u32 l = ent.y >> 16;
    while (true) {
      if (s >= l) { ovf |= 0x10000u << ((l - 1u) / TW); break; }      // limit of row r = r*TW + lim_r, lim_r <= TW
      const u32 old = atomicMin(&tab[s], x);
      if (old == EMPTY) { ++added; break; }
      if (old == x) break;
      x = max(old, x); ++s;
    }
  }
  __syncwarp();
  return ovf | added;
}

// A row whose optimistic table spilled past its 32 spare slots is rebuilt with TW-cap home slots and cap spill
// slots (cap = lenA*W >= IP): a key is pushed right past at most IP-1 smaller keys, so this cannot overflow.
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

__device__ __noinline__ u32 ell_count_table(const u32* tabr, u32 lim) {     // occupied slots (after a rebuild)
  u32 c = 0;
  for (u32 q = lane_id(); q < lim; q += 32) c += (tabr[q] != EMPTY) ? 1u : 0u;
  return __reduce_add_sync(0xffffffffu, c);
}

// ------------------------------------------------------------------------------------------------ the fused kernel
// Every warp is an independent worker on tiles of R consecutive rows.
template <int W, int R>
__global__ void __launch_bounds__(ELL_MAX_WARPS * 32, 1) k_fused_ell(const EllArgs p) {
  constexpr int LPR = W / 4;               // lanes per B row
  constexpr int NSEG = 32 / LPR;           // B rows per LDG.128 warp instruction ("batch")
  constexpr int NBS = 64 / NSEG;           // batches per chunk of 64 A nonzeros
  constexpr int NBG = NBS < 8 ? NBS : 8;   // batches in flight (a "group")
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

  // Tiles are dealt round-robin: in iteration i, warp gw works on tile i*stride + gw.  Consecutive tiles are then
  // processed at the same time by neighbouring warps, so a tile's predecessors publish their aggregates when it
  // does (handing out ids from an atomic counter one tile ahead delayed every look-back by a whole tile time).
  // Every warp of the grid is resident (one CTA per SM), so the chain cannot wait on a tile that never runs.
  const u32 stride = gridDim.x * (blockDim.x >> 5);
  auto load_rowptr = [&](u32 t) -> int {     // lane r (r <= R) gets Arow[t*R + r], clamped to the matrix
    if (t >= p.ntiles) return 0;
    const long long r0 = (long long)t * R;
    const int nr = (int)min((long long)R, (long long)p.An - r0);
    return p.Arow[r0 + min((int)lane, nr)];
  };
  auto load_acol = [&](int abase, int e0, int E, int& j0, int& j1) {
    j0 = -1; j1 = -1;
    if (e0 + (int)lane < E) j0 = p.Acol[abase + e0 + (int)lane];
    if (e0 + 32 + (int)lane < E) j1 = p.Acol[abase + e0 + 32 + (int)lane];
  };
  auto check_acol = [&](int& j0, int& j1) {
    if ((j0 >= p.Bn) | (j1 >= p.Bn) | (j0 < -1) | (j1 < -1)) {
      atomicOr(&p.sc->err, 1u);
      if ((u32)j0 >= (u32)p.Bn) j0 = -1;
      if ((u32)j1 >= (u32)p.Bn) j1 = -1;
    }
  };
  auto load_group = [&](int g, int j0, int j1, uint4 (&v)[NBG]) {      // the B rows of batches g .. g+NBG-1 of a chunk
#pragma unroll
    for (int u = 0; u < NBG; ++u) {
      const int seg = (g + u) * NSEG + (int)sub;
      const int j = __shfl_sync(0xffffffffu, ((g + u) * NSEG < 32) ? j0 : j1, seg & 31);
      v[u] = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
      if (j >= 0) v[u] = __ldg(&Bell4[(size_t)j * LPR + part]);
    }
  };

  // ---- pipeline prologue
  u32 tile = blockIdx.x * (blockDim.x >> 5) + warp;
  int a[R + 1];
  {
    const int ar = load_rowptr(tile);
#pragma unroll
    for (int r = 0; r <= R; ++r) a[r] = __shfl_sync(0xffffffffu, ar, r);
  }
  int j0, j1;
  uint4 v[NBG];
  load_acol(a[0], 0, a[R] - a[0], j0, j1);
  check_acol(j0, j1);
  load_group(0, j0, j1, v);

  while (tile < p.ntiles) {
    const u32 next = (tile + stride < tile) ? 0xffffffffu : tile + stride;
    const int ar_n = load_rowptr(next);                            // in flight during the inserts
    const long long row0 = (long long)tile * R;
    const int nrows = (int)min((long long)R, (long long)p.An - row0);
    const int E = a[R] - a[0];
    int b[R];                                                      // first A nonzero of every row, relative to the tile
#pragma unroll
    for (int r = 0; r < R; ++r) b[r] = a[r] - a[0];
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

    // ---- insert: chunk 0 / group 0 is already in v[] (loaded one tile ago)
    u32 ovf = 0, qn = 0, added = 0;
    for (int e0 = 0; e0 < E; e0 += 64) {
      if (e0 > 0) { load_acol(a[0], e0, E, j0, j1); check_acol(j0, j1); }
#pragma unroll
      for (int g = 0; g < NBS; g += NBG) {
        if (e0 + g * NSEG >= E) break;
        if (g > 0 || e0 > 0) load_group(g, j0, j1, v);
#pragma unroll
        for (int u = 0; u < NBG; ++u) {
          if (e0 + (g + u) * NSEG >= E) break;
          const int e = e0 + (g + u) * NSEG + (int)sub;
          u32 r = 0;
#pragma unroll
          for (int q = 1; q < R; ++q) r += (e >= b[q]) ? 1u : 0u;
          const uint2 pr = par[r];
          const u32 tabr = r * TW;
          const u32 x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          u32 s[4], old[4];
          // padding (EMPTY) takes a harmless atomicMin(.., EMPTY) on a private bank: no branch around the atomics
#pragma unroll
          for (int k = 0; k < 4; ++k) s[k] = tabr + ((x[k] != EMPTY) ? __umulhi(x[k], pr.x) : lane);
#pragma unroll
          for (int k = 0; k < 4; ++k) old[k] = atomicMin(&tab[s[k]], x[k]);
          const u32 hi = (pr.y << 16) + 1u;                  // queue entry: (key, next slot | spill limit << 16)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool valid = x[k] != EMPTY;
            const bool fresh = valid && (old[k] == EMPTY);
            const bool lose = valid && (old[k] != EMPTY) && (old[k] != x[k]);
            ipc += valid ? 1u : 0u;
            added += fresh ? 1u : 0u;
            const u32 m = __ballot_sync(0xffffffffu, lose);
            if (lose) queue[qn + __popc(m & ltmask)] = make_uint2(max(old[k], x[k]), s[k] + hi);
            qn += __popc(m);
          }
          if (qn > ELL_QCAP - 128) { const u32 d = ell_drain(tab, queue, qn, TW); ovf |= d >> 16; added += d & 0xffffu; qn = 0; }
        }
      }
    }
    if (qn) { const u32 d = ell_drain(tab, queue, qn, TW); ovf |= d >> 16; added += d & 0xffffu; }
    ovf = __reduce_or_sync(0xffffffffu, ovf);
    u32 agg = __reduce_add_sync(0xffffffffu, added);
    if (ovf) {                                                     // rare: exact rebuild of the spilled rows, recount
#pragma unroll
      for (int r = 0; r < R; ++r)
        if ((ovf >> r) & 1u) { ell_rebuild_row<W>(p.Acol, p.Bell, (u32)p.Bn, p.Bm, TW, tab + r * TW, a[r], a[r + 1]); lim[r] = TW; }
      agg = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) if (a[r + 1] > a[r]) agg += ell_count_table(tab + r * TW, lim[r]);
    }
    if (lane == 0) chain_post(p.chain, tile, agg);                 // published before the compaction

    // ---- next tile: row pointers have arrived, start its Acol loads
    int an[R + 1];
#pragma unroll
    for (int r = 0; r <= R; ++r) an[r] = __shfl_sync(0xffffffffu, ar_n, r);
    int j0n, j1n;
    load_acol(an[0], 0, an[R] - an[0], j0n, j1n);

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

    // ---- next tile: Acol has arrived, start its B-row loads (v[] is free again)
    check_acol(j0n, j1n);
    load_group(0, j0n, j1n, v);

    // ---- the tile's offset, row pointers, rows to their final position
    u32 incl_mine = 0, run = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) { run += c[r]; if ((int)lane == r) incl_mine = run; }
    const u64 excl = p.debug_nochain ? (u64)tile * (u64)(p.debug_nochain) : chain_exclusive(p.chain, tile, p.ntiles);
    if ((int)lane < nrows) st_rowptr(p.Crow, p.is64, (size_t)(row0 + lane) + 1, excl + incl_mine, &p.sc->err);
    if (tile == 0 && lane == 0) st_rowptr(p.Crow, p.is64, 0, 0, &p.sc->err);
    if (tile == p.ntiles - 1 && lane == 0) p.sc->total_nnz = excl + run;
    u32 off = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const u32* src = tab + r * TW;
      int* dst = p.Ccol + (excl + off);
      for (u32 q = lane; q < c[r]; q += 32) dst[q] = (int)src[q];
      off += c[r];
    }
    __syncwarp();
    tile = next;
#pragma unroll
    for (int r = 0; r <= R; ++r) a[r] = an[r];
    j0 = j0n; j1 = j1n;
  }
  u64 ips = ipc;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ips += __shfl_xor_sync(0xffffffffu, ips, d);
  if (lane == 0 && ips) atomicAdd(&p.sc->total_ip, ips);
}

}  // namespace bsk
