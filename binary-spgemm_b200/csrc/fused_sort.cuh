// binary-spgemm_b200/csrc/fused_sort.cuh — one-pass kernel for REGULAR short-row matrices: register sorting network.
//
// Replaces, like fused_ell.cuh, SpGEMM_bigslice (final/SpGEMM_mpi_omp.c:15-58: Gustavson row product + `xb` flag
// de-duplication + per-row quickSort) and the concatenation / row-pointer fix-up of SpGEMM_omp (:111-141), for
// matrices whose every output row has IP <= LA*W <= 512 (LA = max_len(A) rounded up to a power of two, W = ELL width
// of B).  The hash-table kernel of fused_ell.cuh spends ~560 warp instructions and ~220 L1/shared-memory wavefronts
// per config-3 row on atomics, collision queues, table init and compaction (profiles/r01_v6_*).  Here the candidate
// columns never leave the registers they were loaded into:
//   * a row occupies S lanes with K = LA*W/S <= 16 keys per lane (K/4 LDG.128 of the ELL copy of B, fused_ell.cuh
//     k_build_ell); 32/S rows are handled by one warp pass;
//   * a bitonic network sorts the row in place — exchanges at distance < K are register-to-register min/max, larger
//     distances one SHFL.BFLY + one min/max per key; padding (EMPTY = 0xFFFFFFFF) sorts to the end;
//   * duplicates are adjacent after the sort: neighbour compare, segmented warp scan, and the distinct keys go to a
//     bank-skewed staging buffer in shared memory (the only shared-memory traffic: 1 store + 1 load per output);
//   * no atomics, no collision chains, no overflow path: the cost of a row is a constant, whatever its columns are.
// The scan / deferred commit / tile pipeline are those of fused_ell.cuh (CtaChain): the aggregate of a tile is posted
// when its rows are staged, the tile is committed one tile later from the other half of a ping-pong staging buffer.
#pragma once
#include <stdlib.h>
#include "fused_ell.cuh"

namespace bsk {

constexpr int SORT_MAX_WARPS = 32;        // CtaChain holds 32 warps; the launch uses what the registers of the instantiation allow

struct SortGeom {          // compile-time geometry of k_fused_sort<W, LAL>
  int LPR, LA, S, NQ, K, RP, R, NP;
};
__host__ __device__ constexpr SortGeom sort_geom_rt(int W, int LAL) {
  SortGeom g{};
  g.LPR = W / 4;                               // lanes per B row (one uint4 each)
  g.LA = 1 << LAL;                             // B rows per output row (A row length, rounded up)
  const int cap = g.LA * W;                    // keys per output row (padded)
  // As many keys per lane as possible (exchanges inside a lane cost 1 instruction per key, across lanes 3: SHFL,
  // min, max), subject to: a row spans >= max(LPR, 2) lanes, at most 32 key registers per lane, and the rows of one
  // warp pass own at most 64 A nonzeros (two registers of Acol per lane).
  int k = 32;
  while (k > 4) {
    const int s = cap / k;
    if (s >= (g.LPR > 2 ? g.LPR : 2) && s <= 32 && (32 / s) * g.LA <= 64) break;
    k >>= 1;
  }
  g.K = k;                                     // keys per lane per row
  g.S = cap / k;                               // lanes per row
  g.NQ = g.K / 4;                              // uint4 per lane per row
  g.RP = 32 / g.S;                             // rows per warp pass
  int np = 32 / g.K; if (np < 1) np = 1;       // passes per tile: at most 32 key registers per lane in flight,
  while (np > 1 && (g.RP * np * g.LA > 64 || g.RP * np > 16)) np >>= 1;   // <= 64 A nonzeros and <= 16 rows per tile
  g.NP = np;
  g.R = g.RP * np;                             // rows per tile
  return g;
}
template <int W, int LAL> __host__ __device__ constexpr SortGeom sort_geom() { return sort_geom_rt(W, LAL); }
// The cp.async kernel is the default for big tiles (32 keys per lane, one pass per tile); those also have the
// floating-point network (FLT) when every key is below 2^23.
__host__ __device__ constexpr bool sort_big_tile(int W, int LAL) { const SortGeom g = sort_geom_rt(W, LAL); return g.K == 32 && g.NP == 1; }
constexpr u32 EMPTY_F = 0x3F800000u;        // padding of the ELL copy for the floating-point network: 1.0f
constexpr u32 SORT_FLT_MAX_BM = 1u << 23;   // keys below 2^23: a + b and (a + b) - min(a, b) are exact in binary32
// Host: which sort kernel a (W, LAL) plan runs, and whether with the floating-point network — ONE rule, used where the ELL copy
// is built (its padding value depends on it) and where the kernel is launched.  BSPGEMM_SORT_SYNC / BSPGEMM_SORT_ASYNC force a
// kernel, BSPGEMM_SORT_INT keeps the integer network (A/B runs, tests).
static inline bool sort_plan_async(int W, int LAL) {
  if (getenv("BSPGEMM_SORT_ASYNC")) return true;
  if (getenv("BSPGEMM_SORT_SYNC")) return false;
  return sort_big_tile(W, LAL);
}
static inline bool sort_plan_flt(int W, int LAL, int Bm) {
  return sort_plan_async(W, LAL) && sort_big_tile(W, LAL) && (u32)Bm <= SORT_FLT_MAX_BM && !getenv("BSPGEMM_SORT_INT");
}
// staging words of one tile: R rows of LA*W keys, plus one pad word per 32 (bank skew)
constexpr u32 SORT_HDR = 20;             // per staging buffer: [0] total, [1] tile, [2..2+R) inclusive row counts
__host__ __device__ constexpr u32 sort_stage_words(int R, int LA, int W) { const u32 n = (u32)(R * LA * W); return (SORT_HDR + n + n / 32u + 4u + 3u) & ~3u; }

// Staged keys (key q at word q + q/32 after the SORT_HDR header words: 33 words per 32 keys) -> dst[0..total), whole warp.
// 16-byte stores: `head` keys up to the first 16-byte boundary of dst, then chunks of 4 keys (4 LDS.32 at the skewed staging
// addresses — conflict-free, the skew moves every 8th lane to the next bank — and one STG.128: a warp instruction writes 512
// contiguous bytes), then the `tail` keys: 40 instead of 64 memory instructions per 1024-key tile (config 3, same box: 3.16 -> 3.11 ms).
__device__ __forceinline__ void commit_keys(int* dst, const u32 buf_s, const u32 total) {
  const u32 lane = lane_id();
  const u32 head = min(total, (u32)(((16u - ((u32)(size_t)dst & 15u)) & 15u) >> 2));
  const u32 body = (total - head) >> 2, tail = (total - head) & 3u;
  auto key_s = [&](u32 q) { return buf_s + 4u * (SORT_HDR + q + (q >> 5)); };
  if (lane < head) dst[lane] = (int)lds32(key_s(lane));
  if (lane < tail) dst[head + 4u * body + lane] = (int)lds32(key_s(head + 4u * body + lane));
  int4* dst4 = reinterpret_cast<int4*>(dst + head);
  u32 c = lane;
  for (; c + 32u < body; c += 64u) {
    const u32 q0 = head + 4u * c, q1 = q0 + 128u;
    const u32 a0 = lds32(key_s(q0)), a1 = lds32(key_s(q0 + 1u)), a2 = lds32(key_s(q0 + 2u)), a3 = lds32(key_s(q0 + 3u));
    const u32 b0 = lds32(key_s(q1)), b1 = lds32(key_s(q1 + 1u)), b2 = lds32(key_s(q1 + 2u)), b3 = lds32(key_s(q1 + 3u));
    dst4[c] = make_int4((int)a0, (int)a1, (int)a2, (int)a3);
    dst4[c + 32u] = make_int4((int)b0, (int)b1, (int)b2, (int)b3);
  }
  for (; c < body; c += 32u) {
    const u32 q0 = head + 4u * c;
    const u32 a0 = lds32(key_s(q0)), a1 = lds32(key_s(q0 + 1u)), a2 = lds32(key_s(q0 + 2u)), a3 = lds32(key_s(q0 + 3u));
    dst4[c] = make_int4((int)a0, (int)a1, (int)a2, (int)a3);
  }
}

// (bitonic_sort_rows, the register sorting network, lives in kernels.cuh: the warp-per-row kernels use it too)

// Every warp is an independent worker on tiles of R consecutive rows.  Per iteration (tile t):
//   for each pass: sort the pass's rows in registers, mark first occurrences, scan, write the distinct keys to the
//   current staging buffer; then reload the same registers with the same pass of tile t+1;
//   post the tile's aggregate; commit tile t-1 from the other staging buffer.
// No launch bound: up to 96 registers (no spills — at 80 a reload from thrashed local memory cost 10 % of the warp
// time); the host sizes the CTA from cudaFuncGetAttributes (registers are allocated per SM sub-partition, 16K each).
template <int W, int LAL>
__global__ void __maxnreg__((sort_geom<W, LAL>().K >= 32 ? 96 : 80)) k_fused_sort(const EllArgs p) {
  constexpr SortGeom G = sort_geom<W, LAL>();
  constexpr int LPR = G.LPR, S = G.S, NQ = G.NQ, K = G.K, RP = G.RP, R = G.R, NP = G.NP;
  constexpr u32 SWORDS = sort_stage_words(R, G.LA, W);
  extern __shared__ __align__(16) u32 smem[];
  const u32 warp = threadIdx.x >> 5, lane = lane_id(), nwarps = (blockDim.x >> 5) - 1u;   // compute warps; the last warp is the chain helper
  const u32 nbuf = p.nbuf;                                                                // ring of staging buffers: commit lag = nbuf-1 tiles
  const u32 wwords = nbuf * SWORDS + 64u;                                                 // + the next tile's <= 64 A nonzeros
  const u32 stage_s = (u32)__cvta_generic_to_shared(smem) + warp * (wwords * 4u);
  const u32 acol_s = stage_s + nbuf * SWORDS * 4u;
  CtaChain* cc = reinterpret_cast<CtaChain*>(smem + (size_t)nwarps * wwords);
  for (u32 i = threadIdx.x; i < sizeof(CtaChain) / 4; i += blockDim.x) reinterpret_cast<u32*>(cc)[i] = 0;
  __syncthreads();                          // the only CTA-wide barrier
  if (warp == nwarps) {
    if (!DBG_NOCHAIN(p)) chain_helper_dyn(cc, p.blk_status, &p.sc->tile_counter, p.ntiles, nwarps, p.nbuf - 1u);
    return;
  }
  const u32 ll = lane % S, seg = lane / S;  // lane within its row, row within the pass
  const uint4* __restrict__ Bell4 = reinterpret_cast<const uint4*>(p.Bell);
  u32 ipc = 0;
  const u32 stride = gridDim.x * nwarps;
  const u32 cta_first = blockIdx.x * nwarps;

  auto load_rowptr = [&](u32 t) -> int {     // lane r (r <= R) gets Arow[t*R + r], clamped to the matrix
    if (t >= p.ntiles) return 0;
    const long long r0 = (long long)t * R;
    const int nr = (int)min((long long)R, (long long)p.An - r0);
    return p.Arow[r0 + min((int)lane, nr)];
  };
  auto load_acol = [&](int ar, int& j0, int& j1) {                 // the tile's A nonzeros (<= 64), absent ones select row Bn
    // a launch replayed from a cached plan (bspgemm.cu, mul_launch_fast) got LA from an earlier product: a longer row is flagged,
    // never truncated silently (the host then redoes the product with a fresh plan)
    { const int len = __shfl_down_sync(0xffffffffu, ar, 1) - ar; if ((int)lane < R && len > G.LA) atomicOr(&p.sc->err, 8u); }
    const int a0 = __shfl_sync(0xffffffffu, ar, 0), E = min(__shfl_sync(0xffffffffu, ar, R) - a0, 64);
    j0 = p.Bn; j1 = p.Bn;
    if ((int)lane < E) j0 = acol_checked(p.Acol[a0 + (int)lane], p.Bn);
    if (32 + (int)lane < E) j1 = acol_checked(p.Acol[a0 + 32 + (int)lane], p.Bn);
  };
  auto stash_acol = [&](int j0, int j1) {                          // validate, then park them in shared memory (frees the registers)
    if (((u32)j0 > (u32)p.Bn) | ((u32)j1 > (u32)p.Bn)) {
      atomicOr(&p.sc->err, 1u);
      if ((u32)j0 > (u32)p.Bn) j0 = p.Bn;
      if ((u32)j1 > (u32)p.Bn) j1 = p.Bn;
    }
    sts32(acol_s + 4u * lane, (u32)j0);
    sts32(acol_s + 4u * (32u + lane), (u32)j1);
    __syncwarp();
  };
  // keys of pass q of a tile.  The K keys of lane (seg, ll) are consecutive uint4 of the row's B rows taken in A order:
  // uint4 number t = ll*NQ + u of the row is part t % LPR of B row slot t / LPR, so a lane holds whole (sorted) B rows,
  // or a contiguous piece of one.
  // XPOSE (the config-3 geometry: 4 lanes per B row, 8 lanes and 32 keys per lane per output row): the loads are issued
  // quad-coalesced instead — load u of lane (quad qd, part pp) is part pp of B row slot 2u+qd, a whole 64-byte row per
  // quad and instruction: 8 instead of 32 lines per LDG.128 and no dependence on L1 — and the lane-per-B-row layout is
  // produced when the keys are consumed, by transpose_pass() through the (still unused) staging area of the tile.
  constexpr bool XPOSE = (LPR == 4 && S == 8 && K == 32);
  auto load_pass = [&](int q, int ar, u32 (&x)[K]) {
    const int row = q * RP + (int)seg;
    const int a0 = __shfl_sync(0xffffffffu, ar, 0);
    const int lo = __shfl_sync(0xffffffffu, ar, row), hi = __shfl_sync(0xffffffffu, ar, row + 1);
    if (XPOSE) {
      const int qd = (int)(ll >> 2), pp = (int)(ll & 3u);
#pragma unroll
      for (int u = 0; u < NQ; ++u) {
        const int slot = 2 * u + qd;
        int j = p.Bn;
        if (slot < hi - lo && lo - a0 + slot < 64) j = (int)lds32(acol_s + 4u * (u32)(lo - a0 + slot));   // (< 64: always, unless a stale plan's LA was exceeded — flagged above)
        const uint4 t4 = __ldcg(&Bell4[(size_t)j * LPR + pp]);        // every sector is used exactly once: keep it out of L1
        x[4 * u + 0] = t4.x; x[4 * u + 1] = t4.y; x[4 * u + 2] = t4.z; x[4 * u + 3] = t4.w;
      }
      return;
    }
    constexpr int SLOTS = NQ >= LPR ? NQ / LPR : 1;                // B rows per lane
    constexpr int PARTS = NQ >= LPR ? LPR : NQ;                    // uint4 per B row taken by this lane
#pragma unroll
    for (int g = 0; g < SLOTS; ++g) {
      const int slot = NQ >= LPR ? (int)ll * SLOTS + g : (int)(ll * NQ) / LPR;
      const int part0 = NQ >= LPR ? 0 : (int)(ll * NQ) % LPR;
      int j = p.Bn;
      if (slot < hi - lo && lo - a0 + slot < 64) j = (int)lds32(acol_s + 4u * (u32)(lo - a0 + slot));
#pragma unroll
      for (int c = 0; c < PARTS; ++c) {
        const uint4 t4 = __ldg(&Bell4[(size_t)j * LPR + part0 + c]);
        const int u = g * PARTS + c;
        x[4 * u + 0] = t4.x; x[4 * u + 1] = t4.y; x[4 * u + 2] = t4.z; x[4 * u + 3] = t4.w;
      }
    }
  };
  // XPOSE: quad-coalesced -> lane-per-B-row.  uint4 (row seg, slot, part) lives at seg*64 + (slot>>1)*8 + ((4*(slot&1) +
  // part) ^ (slot>>1)) of the scratch: the XOR makes both the writes (fixed u: the 8 lanes of a row cover one 128-byte
  // line) and the reads (fixed (g,c): the 8 lanes of a row hit 8 different 16-byte bank columns) conflict-free.
  auto transpose_pass = [&](u32 (&x)[K], u32 scratch_s) {
    if (!XPOSE) return;
    const u32 qd = ll >> 2, pp = ll & 3u, rowbase = scratch_s + seg * 1024u;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      sts128(rowbase + 16u * ((u32)u * 8u + ((4u * qd + pp) ^ (u32)u)), x[4 * u], x[4 * u + 1], x[4 * u + 2], x[4 * u + 3]);
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 t4 = lds128(rowbase + 16u * (ll * 8u + ((u32)(4 * g + c) ^ ll)));
        const int u = g * 4 + c;
        x[4 * u + 0] = t4.x; x[4 * u + 1] = t4.y; x[4 * u + 2] = t4.z; x[4 * u + 3] = t4.w;
      }
    __syncwarp();
  };
  // commit of the tile staged in buffer buf_s during iteration `iter` (header: total, tile, inclusive row counts)
  auto commit = [&](u32 iter, u32 buf_s) {
    const u32 total = lds32(buf_s), t = lds32(buf_s + 4u);
    const u32 incl_mine = lds32(buf_s + 8u + 4u * min(lane, (u32)R - 1u));
    const u64 excl = DBG_NOCHAIN(p) ? (u64)t * (u64)DBG_NOCHAIN(p) : chain_resolve(cc, iter, warp);
    const long long row0 = (long long)t * R;
    const int nrows = (int)min((long long)R, (long long)p.An - row0);
    if ((int)lane < nrows) st_rowptr(p.Crow, p.is64, (size_t)(row0 + lane) + 1, excl + incl_mine, &p.sc->err);
    if (t == 0 && lane == 0) st_rowptr(p.Crow, p.is64, 0, 0, &p.sc->err);
    if (t == p.ntiles - 1 && lane == 0) p.sc->total_nnz = excl + total;
    int* dst = p.Ccol + excl;
    u32 src = buf_s + 4u * (SORT_HDR + lane);                      // key q lives at word q + q/32: 33 words per 32 keys
    for (u32 q = lane; q < total; q += 32, src += 132u) dst[q] = (int)lds32(src);   // (commit_keys' 16-byte stores: 0.214 vs 0.210 ms at config 2)
    __syncwarp();
  };

  // ---- pipeline prologue
  // tile of this warp in the CTA's iteration `it`: block ids come from the chain helper (dynamic deal, see chain_helper_dyn)
  const u32 nblocks = (p.ntiles + nwarps - 1u) / nwarps;
  auto tile_of = [&](u32 it) -> u32 {
    if (DBG_NOCHAIN(p)) { const unsigned long long t = (unsigned long long)it * stride + cta_first + warp; return t < p.ntiles ? (u32)t : 0xffffffffu; }
    const u32 blk = chain_block_of(cc, it);
    if (blk >= nblocks) return 0xffffffffu;
    const u32 t = blk * nwarps + warp;
    return t < p.ntiles ? t : 0xffffffffu;
  };
  u32 tile = tile_of(0), iter = 0;
  int arn = load_rowptr(tile_of(1));
  u32 x[NP][K];
  {
    const int ar = load_rowptr(tile);
    int j0, j1;
    load_acol(ar, j0, j1);
    stash_acol(j0, j1);
#pragma unroll
    for (int q = 0; q < NP; ++q) load_pass(q, ar, x[q]);
    __syncwarp();
  }
  u32 cur_buf = 0;                           // iter % nbuf

  while (tile < p.ntiles) {
    const u32 next = tile_of(iter + 1u);
    const int arnn = load_rowptr(tile_of(iter + 2u));
    int j0n, j1n;
    load_acol(arn, j0n, j1n);
    const u32 buf_s = stage_s + cur_buf * (SWORDS * 4u), cur_s = buf_s + 4u * SORT_HDR;
    u32 run = 0, incl_mine = 0;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      u32 (&k)[K] = x[q];
      transpose_pass(k, cur_s);
      bitonic_sort_rows<K, S, (W < K ? W : K)>(k, ll);     // every B row is ascending in the ELL copy (k_build_ell sorts it)
      // the row is ascending along (lane, register); EMPTY (padding) is the largest value
      u32 prev_last = __shfl_up_sync(0xffffffffu, k[K - 1], 1);
      if (ll == 0) prev_last = EMPTY;                              // nothing before the row's first key (EMPTY never counts)
      bool plain = (k[K - 1] != EMPTY) && (k[0] != prev_last);     // no padding in this lane, no duplicate
#pragma unroll
      for (int i = 1; i < K; ++i) plain = plain && (k[i] != k[i - 1]);
      if (__all_sync(0xffffffffu, plain)) {
        // the usual case: no duplicate, no padding anywhere in the pass — every key's place is known in advance
        ipc += (u32)K;
        const u32 pos = lane * (u32)K;                             // rows of the pass back to back, lane-major
        const u32 a0s = cur_s + 4u * (run + pos + ((run + pos) >> 5));
        if (((run + pos) & 31u) + (u32)K <= 32u || (K % 32 == 0 && ((run + pos) & 31u) == 0u)) {
#pragma unroll
          for (int i = 0; i < K; ++i) sts32(a0s + 4u * (u32)(i + i / 32), k[i]);
        } else {
#pragma unroll
          for (int i = 0; i < K; ++i) { const u32 o = run + pos + (u32)i; sts32(cur_s + 4u * (o + (o >> 5)), k[i]); }
        }
#pragma unroll
        for (int sq = 0; sq < RP; ++sq) { run += (u32)(K * S); if ((int)lane == q * RP + sq) incl_mine = run; }
      } else {
        // first occurrences, their count, inclusive scan of the count inside the row's S lanes
        bool f[K];
        u32 cnt = 0;
#pragma unroll
        for (int i = 0; i < K; ++i) {
          f[i] = (k[i] != EMPTY) && (k[i] != (i ? k[i - 1] : prev_last));
          cnt += f[i] ? 1u : 0u;
          ipc += (k[i] != EMPTY) ? 1u : 0u;
        }
        u32 inc = cnt;
#pragma unroll
        for (int d = 1; d < S; d <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, inc, d); if ((int)ll >= d) inc += t; }
        // rows of the pass are staged back to back, in row order
        u32 rowbase = run;
#pragma unroll
        for (int sq = 0; sq < RP; ++sq) {
          const u32 tot = __shfl_sync(0xffffffffu, inc, sq * S + S - 1);
          if ((int)seg > sq) rowbase += tot;
          run += tot;
          if ((int)lane == q * RP + sq) incl_mine = run;
        }
        u32 o = rowbase + inc - cnt;
#pragma unroll
        for (int i = 0; i < K; ++i) if (f[i]) { sts32(cur_s + 4u * (o + (o >> 5)), k[i]); ++o; }
      }
      // the registers of this pass are free: refill them with the same pass of the next tile
      if (q == 0) stash_acol(j0n, j1n);
      load_pass(q, arn, k);
    }
    if (lane == 0) sts64(buf_s, run, tile);
    if (lane < (u32)R) sts32(buf_s + 8u + 4u * lane, incl_mine);
    __syncwarp();
    if (!DBG_NOCHAIN(p)) chain_post(cc, iter, warp, run);
    // commit the tile staged nbuf-1 iterations ago: its buffer is the next one of the ring
    cur_buf = (cur_buf + 1u == nbuf) ? 0u : cur_buf + 1u;
    if (iter + 1u >= nbuf) commit(iter + 1u - nbuf, stage_s + cur_buf * (SWORDS * 4u));
    tile = next; ++iter;
    arn = arnn;
  }
  // drain: the last nbuf-1 tiles of this warp
  for (u32 k = (iter + 1u >= nbuf) ? iter + 1u - nbuf : 0u; k < iter; ++k) commit(k, stage_s + (k % nbuf) * (SWORDS * 4u));
  u64 ips = ipc;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ips += __shfl_xor_sync(0xffffffffu, ips, d);
  if (lane == 0 && ips) atomicAdd(&p.sc->total_ip, ips);
}

// ---------------------------------------------------------------------------------------------- asynchronous-input variant
// k_fused_sort keeps the next tile's B rows in flight in the key registers, so the gather can only be issued after the
// current tile has been sorted and staged: its DRAM latency is covered by the commit alone and the warp then waits
// (long-scoreboard stalls: 22 % of the warp time at config 3, profiles/r01_sort_cfg3_ncu_summary.txt).  Here the B rows of
// tile t+1 are copied global -> shared memory with cp.async (LDGSTS.128, L1 bypassed, no destination registers) right
// after tile t has been read into registers, i.e. the gather is in flight during the whole sort of tile t:
//   iteration t:  wait for the copies of tile t; keys <- input buffer (LDS.128); issue the copies of tile t+1 into the (now
//                 free) input buffer, the Acol loads of tile t+2 (its B-row table) and the row pointers of tile t+3; sort;
//                 commit tile t-LAG from the staging buffer tile t is about to use; stage; post.
// (Measured alternatives, profiles/r01_sort_async_sweeps.txt: committing at the top of the iteration, where no key register
// is live, costs 9 % — the copies' DEPBAR then also waits for the commit's stores; forcing the LDGSTS into a block of their
// own instead of letting ptxas spread them over the first half of the network changes nothing.)
// The copies land directly in the lane-per-B-row layout the network wants (16-byte piece pr of the pass goes to
// L*NQ + (c ^ swz(L)), L = pr / NQ the consuming lane, c = pr % NQ): the copy instructions are row-coalesced (LPR lanes
// per B row) and both the LDGSTS writes and the LDS.128 reads are bank-conflict-free, so the register -> shared -> register
// transpose of k_fused_sort is gone too.
// Shared memory per warp: the input buffer + LAG staging buffers (the tile's B-row table stays in registers).  A tile can
// only be committed once the offsets of ALL earlier tiles of the grid are known (the chain), about one iteration after it
// was posted, and here the commit of tile t-LAG comes before tile t is staged: with LAG = 1 the warps wait for the chain in
// every iteration (config 3: 3.64 ms, 2.61 ms with the scan switched off), so the host launches LAG = 2.
//
// Round 2 (profiles/r02_*): the loop body of the round-1 kernel was 43 KB of code (the instruction cache holds 32 KB: 6 % of
// the warp time was spent waiting for instructions), 1111 of its 2187 warp instructions per 1024-key tile went to the ALU pipe
// (65 % busy, the limiter of the sort phase) and only 619 of those were the comparators' VIMNMX proper.  Changes:
//   * staging in 16-byte groups, XOR-swizzled instead of padded: the sorted keys leave the registers as STS.128 (8 instead of
//     32 stores per lane and tile) and the commit reads aligned LDS.128 pairs, funnel-shifted by the destination's phase
//     (16 LDS.128 instead of 32 LDS.32 with an address computation each).  (Tried: a 36-word stride — the padding cost the 18th
//     warp, +5 %; the rare passes with a duplicate or padding fixed up out of line through shared memory — the call made ptxas
//     spill loop invariants to local memory, one exposed LDL latency per tile, 14 % of the warp time.)
//   * FLT (Bm <= 2^23: every key is a subnormal float with the same bit pattern): the network runs in the floating-point
//     domain.  In-lane comparators take two neighbouring keys at a time, lo = FMNMX (ALU pipe), hi = (a + b) - lo as two
//     FADD2 on packed register pairs (FMA pipe): 1 + 1 instructions per comparator instead of 1 + 2 (IMAD form) or 2 ALU.
//     Cross-lane exchanges: lanes that keep the maximum hold NEGATED keys for the duration of the cross-lane stages, so that
//     both partners execute the same FMNMX(own, -received) (the negation is a free source modifier): one ALU instruction
//     per key instead of VIMNMX + predicated VIMNMX; switching a lane between the two representations is one FMUL2 per
//     two keys.  Padding is 1.0f (EMPTY_F): larger than every key, and x + 1 - x == 1 exactly in this range.
__device__ __forceinline__ void cp_async16(u32 dst_s, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst_s), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__host__ __device__ constexpr u32 sort_input_words(int R, int LA, int W) { return (u32)(R * LA * W); }

// staging buffer of the async kernel: header, then the keys in 16-byte groups of 4, group g = 8*l + j (l = the 32-key block, the
// keys of one lane in the plain case) stored at group slot 8*l + (j ^ (l & 7)): the lanes' STS.128 (fixed j, 8 consecutive l) and
// the commit's LDS.128 (fixed l, 8 consecutive j) are both bank-conflict-free WITHOUT padding words (a 36-word stride cost the
// 18th warp of config 3); 8 spare words for the funnel's look-ahead
__host__ __device__ constexpr u32 sort_stage_words_a(int R, int LA, int W) { const u32 n = (u32)(R * LA * W); return (SORT_HDR + n + 8u + 3u) & ~3u; }
__device__ __forceinline__ u32 sgrp_s(u32 buf_s, u32 g) { return buf_s + 4u * SORT_HDR + 16u * ((g & ~7u) | ((g ^ (g >> 3)) & 7u)); }   // group g (keys 4g..4g+3)
__device__ __forceinline__ u32 skey_s(u32 buf_s, u32 q) { return sgrp_s(buf_s, q >> 2) + 4u * (q & 3u); }

// Staged keys -> dst[0..total), whole warp, 16-byte stores: `head` keys up to the first 16-byte boundary of dst, then chunks
// of 4 keys — chunk c is keys head+4c .. head+4c+3, i.e. the aligned groups c and c+1 of the staging buffer shifted by head —
// then the tail.
template <int H>
__device__ __forceinline__ void commit_body(int4* dst4, const u32 buf_s, const u32 body) {
  const u32 lane = lane_id();
  auto grp = [&](u32 g) { return lds128(sgrp_s(buf_s, g)); };
  auto pick = [&](const uint4 a, const uint4 b) {
    return H == 0 ? make_int4((int)a.x, (int)a.y, (int)a.z, (int)a.w) : H == 1 ? make_int4((int)a.y, (int)a.z, (int)a.w, (int)b.x)
         : H == 2 ? make_int4((int)a.z, (int)a.w, (int)b.x, (int)b.y) : make_int4((int)a.w, (int)b.x, (int)b.y, (int)b.z);
  };
  u32 c = lane;
  for (; c + 32u < body; c += 64u) {
    const uint4 a0 = grp(c), a1 = grp(c + 32u);
    uint4 b0 = a0, b1 = a1;
    if (H) { b0 = grp(c + 1u); b1 = grp(c + 33u); }
    dst4[c] = pick(a0, b0);
    dst4[c + 32u] = pick(a1, b1);
  }
  if (c < body) {
    const uint4 a0 = grp(c);
    uint4 b0 = a0;
    if (H) b0 = grp(c + 1u);
    dst4[c] = pick(a0, b0);
  }
}
__device__ __forceinline__ void commit_keys128(int* dst, const u32 buf_s, const u32 total) {
  const u32 lane = lane_id();
  const u32 head = min(total, (u32)(((16u - ((u32)(size_t)dst & 15u)) & 15u) >> 2));
  const u32 body = (total - head) >> 2, tail = (total - head) & 3u;
  if (lane < head) dst[lane] = (int)lds32(skey_s(buf_s, lane));
  if (lane < tail) dst[head + 4u * body + lane] = (int)lds32(skey_s(buf_s, head + 4u * body + lane));
  int4* dst4 = reinterpret_cast<int4*>(dst + head);
  if (head == 0u) commit_body<0>(dst4, buf_s, body);
  else if (head == 1u) commit_body<1>(dst4, buf_s, body);
  else if (head == 2u) commit_body<2>(dst4, buf_s, body);
  else commit_body<3>(dst4, buf_s, body);
}

// ---- the sorting network in the floating-point domain (see the header comment; same element order as bitonic_sort_rows)
__device__ __forceinline__ float as_f(u32 v) { return __uint_as_float(v); }
__device__ __forceinline__ u32 as_u(float v) { return __float_as_uint(v); }
// two comparators on neighbouring registers: (a0,b0) and (a1,b1); minima to a, maxima to b
__device__ __forceinline__ void cmp2_f(u32& a0, u32& a1, u32& b0, u32& b1) {
  const float2 A = make_float2(as_f(a0), as_f(a1)), B = make_float2(as_f(b0), as_f(b1));
  const float2 S2 = __fadd2_rn(A, B);
  const float l0 = fminf(A.x, B.x), l1 = fminf(A.y, B.y);
  const float2 H = __fadd2_rn(S2, make_float2(-l0, -l1));
  a0 = as_u(l0); a1 = as_u(l1); b0 = as_u(H.x); b1 = as_u(H.y);
}
__device__ __forceinline__ void cmp1_f(u32& a, u32& b) {
  const float x = as_f(a), y = as_f(b);
  a = as_u(fminf(x, y)); b = as_u(fmaxf(x, y));
}
template <int K>
__device__ __forceinline__ void flip_sign(u32 (&x)[K], const float sg) {     // x *= sg (sg = +-1), two keys per FMUL2
  const float2 s2 = make_float2(sg, sg);
#pragma unroll
  for (int k = 0; k < K; k += 2) {
    const float2 v = __fmul2_rn(make_float2(as_f(x[k]), as_f(x[k + 1])), s2);
    x[k] = as_u(v.x); x[k + 1] = as_u(v.y);
  }
}
template <int K, int S, int RUN>
__device__ __forceinline__ void bitonic_sort_rows_f(u32 (&x)[K], const u32 ll) {
  constexpr int N = K * S;
  static_assert(K % 2 == 0, "pairs");
#pragma unroll
  for (int size = 2 * RUN; size <= N; size <<= 1) {
    bool neg = false;                                                // this lane holds negated keys (cross-lane stages only)
    if (size <= K) {                                                 // mirror inside the lane
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int pk = k ^ (size - 1);
        if (k < pk) cmp1_f(x[k], x[pk]);
      }
    } else {                                                         // mirror across lanes: register k <-> K-1-k of lane ^ (size/K-1)
      const u32 lm = (u32)(size / K - 1);
      neg = (ll & (u32)(size / (2 * K))) != 0u;                      // the lane keeps the maxima: as minima of negated keys
      flip_sign<K>(x, neg ? -1.0f : 1.0f);
#pragma unroll
      for (int k = 0; k < K / 2; ++k) {
        const float ya = as_f(__shfl_xor_sync(0xffffffffu, x[K - 1 - k], lm));
        const float yb = as_f(__shfl_xor_sync(0xffffffffu, x[k], lm));
        x[k] = as_u(fminf(as_f(x[k]), -ya));
        x[K - 1 - k] = as_u(fminf(as_f(x[K - 1 - k]), -yb));
      }
    }
#pragma unroll
    for (int d = size >> 2; d >= 1; d >>= 1) {
      if (d >= K) {                                                  // partner key lives in lane ^ (d/K)
        const u32 ld = (u32)(d / K);
        const bool want = (ll & ld) != 0u;
        flip_sign<K>(x, want != neg ? -1.0f : 1.0f);
        neg = want;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float y = as_f(__shfl_xor_sync(0xffffffffu, x[k], ld));
          x[k] = as_u(fminf(as_f(x[k]), -y));
        }
      } else {
        if (d == K / 2 || (size <= K && d == size >> 2)) {           // first in-lane stage after the cross-lane ones: back to plain keys
          if (size > K) { flip_sign<K>(x, neg ? -1.0f : 1.0f); neg = false; }
        }
        if (d >= 2) {                                                // both keys in this lane, two comparators per instruction pair
#pragma unroll
          for (int k = 0; k < K; k += 2)
            if ((k & d) == 0) cmp2_f(x[k], x[k + 1], x[k | d], x[(k | d) + 1]);
        } else {
#pragma unroll
          for (int k = 0; k < K; k += 2) cmp1_f(x[k], x[k + 1]);
        }
      }
    }
  }
}

template <int W, int LAL, bool FLT>
__global__ void __maxnreg__((sort_geom<W, LAL>().K >= 32 ? 96 : 80)) k_fused_sort_async(const EllArgs p) {
  constexpr SortGeom G = sort_geom<W, LAL>();
  constexpr int LPR = G.LPR, LA = G.LA, S = G.S, NQ = G.NQ, K = G.K, RP = G.RP, R = G.R, NP = G.NP;
  constexpr u32 SWORDS = sort_stage_words_a(R, LA, W), IWORDS = sort_input_words(R, LA, W);
  constexpr u32 EMP = FLT ? EMPTY_F : EMPTY;                          // padding value of the ELL copy (k_build_ell's `pad`)
  constexpr int SH = NQ >= 8 ? 0 : NQ == 4 ? 1 : NQ == 2 ? 2 : 3;     // swz(L) = (L >> SH) & (NQ-1): 8 consecutive lanes hit 8 different 16-byte bank columns
  static_assert(NQ <= 8 && R * LA <= 64, "geometry");
  extern __shared__ __align__(16) u32 smem[];
  const u32 warp = threadIdx.x >> 5, lane = lane_id(), nwarps = (blockDim.x >> 5) - 1u;   // compute warps; the last warp is the chain helper
  const u32 lag = p.nbuf - 1u;                                                            // staging buffers = commit lag in tiles
  const u32 wwords = IWORDS + lag * SWORDS;                                               // input buffer, staging ring
  const u32 in_s = (u32)__cvta_generic_to_shared(smem) + warp * (wwords * 4u);
  const u32 stage_s = in_s + IWORDS * 4u;
  CtaChain* cc = reinterpret_cast<CtaChain*>(smem + (size_t)nwarps * wwords);
  for (u32 i = threadIdx.x; i < sizeof(CtaChain) / 4; i += blockDim.x) reinterpret_cast<u32*>(cc)[i] = 0;
  __syncthreads();                          // the only CTA-wide barrier
  if (warp == nwarps) {
    if (!DBG_NOCHAIN(p)) chain_helper_dyn(cc, p.blk_status, &p.sc->tile_counter, p.ntiles, nwarps, lag);
    return;
  }
  const u32 ll = lane % S;                  // lane within its row
  const uint4* __restrict__ Bell4 = reinterpret_cast<const uint4*>(p.Bell);
  u32 ipc = 0;
  const u32 stride = gridDim.x * nwarps;
  const u32 cta_first = blockIdx.x * nwarps;

  auto load_rowptr = [&](u32 t) -> int {     // lane r (r <= R) gets Arow[t*R + r], clamped to the matrix
    if (t >= p.ntiles) return 0;
    const long long r0 = (long long)t * R;
    const int nr = (int)min((long long)R, (long long)p.An - r0);
    return p.Arow[r0 + min((int)lane, nr)];
  };
  // B-row table of a tile: entry e = row*LA + slot is the B row the slot gathers, Bn (the all-EMPTY row) when the A row
  // is shorter.  Lane l loads entries l and l+32 straight from Acol (whole rows: the same coalesced accesses).
  auto load_jtab = [&](int ar, int& j0, int& j1) {
    j0 = p.Bn; j1 = p.Bn;
    // a launch replayed from a cached plan (bspgemm.cu, mul_launch_fast) got LA from an earlier product: a longer row is flagged,
    // never truncated silently (the host then redoes the product with a fresh plan)
    { const int len = __shfl_down_sync(0xffffffffu, ar, 1) - ar; if ((int)lane < R && len > LA) atomicOr(&p.sc->err, 8u); }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int e = h * 32 + (int)lane, row = e / LA, slot = e % LA;
      const int lo = __shfl_sync(0xffffffffu, ar, row & 31), hi = __shfl_sync(0xffffffffu, ar, (row + 1) & 31);
      if (row < R && slot < hi - lo) { const int j = acol_checked(p.Acol[lo + slot], p.Bn); if (h) j1 = j; else j0 = j; }
    }
  };
  auto check_jtab = [&](int& j0, int& j1) {                        // an A column outside [0,Bn): flag it, gather nothing
    if (((u32)j0 > (u32)p.Bn) | ((u32)j1 > (u32)p.Bn)) {
      atomicOr(&p.sc->err, 1u);
      if ((u32)j0 > (u32)p.Bn) j0 = p.Bn;
      if ((u32)j1 > (u32)p.Bn) j1 = p.Bn;
    }
  };
  // copies of a whole tile: NP*NQ LDGSTS.128 per lane.  Copy u of lane l is piece gp = (u % NQ)*32 + l of pass u / NQ:
  // part l % LPR of table entry e = pass*RP*LA + gp / LPR, which lane e % 32 holds in j0 (e < 32) or j1: the table never
  // leaves the registers (one SHFL per copy; in shared memory it cost 256 bytes per warp — the 18th warp of config 3).
  auto issue_tile = [&](int j0, int j1) {
#pragma unroll
    for (int u = 0; u < NP * NQ; ++u) {
      const int q = u / NQ, v = u % NQ;
      const int c0 = q * RP * LA + v * (32 / LPR);                 // multiple of 32/LPR: e >= 32 iff c0 >= 32
      const u32 j = (u32)__shfl_sync(0xffffffffu, c0 >= 32 ? j1 : j0, (c0 & 31) + (int)(lane / (u32)LPR));
      const u32 L = (u32)(v * (32 / NQ)) + lane / (u32)NQ, c = lane % (u32)NQ;
      const u32 pos = (u32)(q * 32 * NQ) + L * (u32)NQ + (c ^ ((L >> SH) & (u32)(NQ - 1)));
      cp_async16(in_s + 16u * pos, &Bell4[(size_t)j * LPR + (lane % (u32)LPR)]);
    }
    cp_async_commit();
  };
  auto read_pass = [&](int q, u32 (&x)[K]) {
#pragma unroll
    for (int c = 0; c < NQ; ++c) {
      const uint4 t4 = lds128(in_s + 16u * ((u32)(q * 32 * NQ) + lane * (u32)NQ + ((u32)c ^ ((lane >> SH) & (u32)(NQ - 1)))));
      x[4 * c + 0] = t4.x; x[4 * c + 1] = t4.y; x[4 * c + 2] = t4.z; x[4 * c + 3] = t4.w;
    }
  };
  // commit of the tile staged in buffer buf_s during iteration `iter` (header: total, tile, inclusive row counts)
  auto commit = [&](u32 iter, u32 buf_s) {
    const u32 total = lds32(buf_s), t = lds32(buf_s + 4u);
    const u32 incl_mine = lds32(buf_s + 8u + 4u * min(lane, (u32)R - 1u));
    const u64 excl = DBG_NOCHAIN(p) ? (u64)t * (u64)DBG_NOCHAIN(p) : chain_resolve(cc, iter, warp);
    const long long row0 = (long long)t * R;
    const int nrows = (int)min((long long)R, (long long)p.An - row0);
    if ((int)lane < nrows) st_rowptr(p.Crow, p.is64, (size_t)(row0 + lane) + 1, excl + incl_mine, &p.sc->err);
    if (t == 0 && lane == 0) st_rowptr(p.Crow, p.is64, 0, 0, &p.sc->err);
    if (t == p.ntiles - 1 && lane == 0) p.sc->total_nnz = excl + total;
    commit_keys128(p.Ccol + excl, buf_s, total);
    __syncwarp();
  };

  // tile of this warp in the CTA's iteration `it`: block ids come from the chain helper (dealt 5 iterations ahead)
  const u32 nblocks = (p.ntiles + nwarps - 1u) / nwarps;
  auto tile_of = [&](u32 it) -> u32 {
    if (DBG_NOCHAIN(p)) { const unsigned long long t = (unsigned long long)it * stride + cta_first + warp; return t < p.ntiles ? (u32)t : 0xffffffffu; }
    const u32 blk = chain_block_of(cc, it);
    if (blk >= nblocks) return 0xffffffffu;
    const u32 t = blk * nwarps + warp;
    return t < p.ntiles ? t : 0xffffffffu;
  };
  // ---- pipeline prologue: copies of tile 0 in flight, table of tile 1 in registers, row pointers of tile 2
  u32 tile = tile_of(0), iter = 0;
  int j0n, j1n, ar2;
  {
    int j0, j1;
    load_jtab(load_rowptr(tile), j0, j1);
    check_jtab(j0, j1);
    issue_tile(j0, j1);
    load_jtab(load_rowptr(tile_of(1)), j0n, j1n);
    ar2 = load_rowptr(tile_of(2));
  }
  u32 cur_buf = 0;                           // iter % lag

  while (tile < p.ntiles) {
    const u32 next = tile_of(iter + 1u);
    const int ar3 = load_rowptr(tile_of(iter + 3u));
    const u32 buf_s = stage_s + cur_buf * (SWORDS * 4u);
    u32 x[NP][K];
    cp_async_wait_all();
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NP; ++q) read_pass(q, x[q]);
    __syncwarp();                            // every lane has read the input buffer
    check_jtab(j0n, j1n);                    // table of tile t+1
    issue_tile(j0n, j1n);                    // ... and its copies, in flight during the sort (a tile that does not exist copies the EMPTY row)
    load_jtab(ar2, j0n, j1n);                // table of tile t+2
    u32 run = 0, incl_mine = 0;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      u32 (&k)[K] = x[q];
      // every B row is ascending in the ELL copy (k_build_ell sorts it)
      if (FLT) bitonic_sort_rows_f<K, S, (W < K ? W : K)>(k, ll);
      else     bitonic_sort_rows<K, S, (W < K ? W : K), (K >= 32 ? 1 : 0)>(k, ll, p.one, p.mone);   // big tiles: two of three in-lane maxima on the FMA pipe (cmpx, kernels.cuh)
      if (q == 0 && iter >= lag) commit(iter - lag, buf_s);        // frees the staging buffer this tile is about to use
      // the row is ascending along (lane, register); padding is the largest value
      u32 prev_last = __shfl_up_sync(0xffffffffu, k[K - 1], 1);
      if (ll == 0) prev_last = EMP;                                // nothing before the row's first key (padding never counts)
      bool plain = (k[K - 1] != EMP) && (k[0] != prev_last);       // no padding in this lane, no duplicate
#pragma unroll
      for (int i = 1; i < K; ++i) plain = plain && (k[i] != k[i - 1]);
      if (__all_sync(0xffffffffu, plain) && (NP == 1 || (run & (u32)(4 * K - 1)) == 0u)) {
        // the usual case: no duplicate, no padding anywhere in the pass — the lane's K keys are K/4 whole groups (STS.128)
        ipc += (u32)K;
        const u32 g0 = (run >> 2) + lane * (u32)(K / 4);
#pragma unroll
        for (int i = 0; i < K; i += 4) sts128(sgrp_s(buf_s, g0 + (u32)(i / 4)), k[i], k[i + 1], k[i + 2], k[i + 3]);
#pragma unroll
        for (int sq = 0; sq < RP; ++sq) { run += (u32)(K * S); if ((int)lane == q * RP + sq) incl_mine = run; }
      } else {
        // first occurrences, their count, inclusive scan of the count inside the row's S lanes
        const u32 seg = lane / S;
        bool f[K];
        u32 cnt = 0;
#pragma unroll
        for (int i = 0; i < K; ++i) {
          f[i] = (k[i] != EMP) && (k[i] != (i ? k[i - 1] : prev_last));
          cnt += f[i] ? 1u : 0u;
          ipc += (k[i] != EMP) ? 1u : 0u;
        }
        u32 inc = cnt;
#pragma unroll
        for (int d = 1; d < S; d <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, inc, d); if ((int)ll >= d) inc += t; }
        // rows of the pass are staged back to back, in row order
        u32 rowbase = run;
#pragma unroll
        for (int sq = 0; sq < RP; ++sq) {
          const u32 tot = __shfl_sync(0xffffffffu, inc, sq * S + S - 1);
          if ((int)seg > sq) rowbase += tot;
          run += tot;
          if ((int)lane == q * RP + sq) incl_mine = run;
        }
        u32 o = rowbase + inc - cnt;
#pragma unroll
        for (int i = 0; i < K; ++i) if (f[i]) { sts32(skey_s(buf_s, o), k[i]); ++o; }
      }
    }
    if (lane == 0) sts64(buf_s, run, tile);
    if (lane < (u32)R) sts32(buf_s + 8u + 4u * lane, incl_mine);
    __syncwarp();
    if (!DBG_NOCHAIN(p)) chain_post(cc, iter, warp, run);
    cur_buf = (cur_buf + 1u == lag) ? 0u : cur_buf + 1u;
    tile = next; ++iter;
    ar2 = ar3;
  }
  cp_async_wait_all();
  // drain: the last `lag` tiles of this warp
  for (u32 k = (iter >= lag) ? iter - lag : 0u; k < iter; ++k) commit(k, stage_s + (k % lag) * (SWORDS * 4u));
  u64 ips = ipc;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ips += __shfl_xor_sync(0xffffffffu, ips, d);
  if (lane == 0 && ips) atomicAdd(&p.sc->total_ip, ips);
}

}  // namespace bsk
