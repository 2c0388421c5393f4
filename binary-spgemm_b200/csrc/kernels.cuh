// binary-spgemm_b200/csrc/kernels.cuh — sm_100a kernels for the boolean CSR product C = A·B.
//
// Replaces the reference's CPU hot loop SpGEMM_bigslice (final/SpGEMM_mpi_omp.c:15-58: Gustavson row
// product + `xb` flag de-duplication + per-row quickSort) and the concatenation / row-pointer fix-up of
// SpGEMM_omp (:111-141).  Result contract: Crow[0]=0, Crow[i+1]-Crow[i] = number of
// distinct k with A(i,j) and B(j,k), Ccol ascending inside each row.
//
// Design: rows are binned by their intermediate-product count IP_i (k_estimate):
//   S  (IP <= cap_s)  one warp per row, table in shared memory              k_rows_warp<G,MODE>
//   M  (IP <= 16384)  one CTA per row, CTA-wide register sort               k_rows_sort<K,T,MODE>   (rows_sort.cuh)
//   L  (larger)       one CTA per row, windowed shared-memory bitmap        k_rows_window<MODE>     (rows_window.cuh)
//                     (matrices with very many columns: bitmap over [0,Bm) in global memory, k_rows_gbitmap<MODE>)
// De-duplication AND sorting are one step: an *ordered* open-addressing table.  Keys are placed by a
// monotone map slot = floor((k-lo)*T/(hi-lo+1)) and collisions are resolved with atomicMin + "the larger
// key moves one slot right" (no wrap-around).  The final table, read left to right, is the sorted set of
// distinct keys — no comparison sort anywhere (the reference spends 2/3 of its time in quickSort).
// Rows whose column span is narrow use a bitmap over [lo,hi] in the same memory instead.
// MODE_FUSED additionally carries a decoupled-look-back scan over row tiles so that B is gathered once
// and C is written once, directly at its final position (one pass; north-star steps 2+3+4 in one launch).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bsk {

typedef unsigned long long u64;
typedef uint32_t u32;

constexpr u32 EMPTY = 0xFFFFFFFFu;
constexpr int MODE_COUNT = 0;   // symbolic: cnt[row] = nnz(C_row)
constexpr int MODE_FILL  = 1;   // numeric : Ccol[Crow[row]..] = sorted distinct columns
constexpr int MODE_FUSED = 2;   // symbolic + scan + numeric in one pass (tiles of consecutive rows)
constexpr int MODE_STAGE = 3;   // big rows, one pass: cnt[row] AND the sorted row at temp[tofs[row]..] (k_copy_rows moves it to Ccol)

constexpr int WARPS_S = 8;      // warps per CTA in the warp-per-row kernel (= rows per tile in MODE_FUSED)

// status word of the decoupled look-back chain: [63:62] flag, [61:0] value
constexpr u64 ST_AGG = 1ull << 62;
constexpr u64 ST_INC = 2ull << 62;
constexpr u64 ST_VAL = (1ull << 62) - 1;

struct Csr {            // A and B as the reference passes them (final/SpGEMM_mpi_omp.c:155-157)
  const int* __restrict__ Arow;   // An+1 ABSOLUTE offsets into Acol (Arow[0] need not be 0, cf. :171)
  const int* __restrict__ Acol;
  const int* __restrict__ Brow;   // Bn+1
  const int* __restrict__ Bcol;
  int An, Bn, Bm;
};

struct DevScalars {     // one per context, in device memory; copied to the host between phases
  u64 total_ip;         // Σ IP_i
  u64 total_nnz;        // nnz(C) (written by the scan / fused kernel)
  u32 hist[34];         // hist[0] = #rows with IP==0, hist[1]: IP==1, hist[b]: IP in (2^(b-2), 2^(b-1)]
  u32 max_ip;
  u32 err;              // bit0: column index of A out of [0,Bn); bit1: 32-bit row pointer overflow; bit2: column of B out of [0,Bm)
  u32 n_m1, n_m2, n_l;  // list lengths (k_build_lists)
  u32 tile_counter;     // dynamic tile ids (fused kernel / scan kernel)
  u32 max_len_a, max_len_b;   // longest row of A / of B (k_maxlen)
  u32 span_rows, span_narrow; // k_probe_span: sampled non-empty rows / those whose candidate columns span < 2^15
  u32 win_ctr[8];             // k_rows_window: next list entry, one counter per launch (3 lists x COUNT/FILL)
  u32 span_runs;              // k_probe_span: sampled rows whose B rows are all runs of consecutive columns
  u32 band_fail;              // band.cuh: bit0 = a B row is not a run / an output row is wider than the register bitmap
  u64 temp_used;              // k_build_lists: words of the staging arena handed to big rows (Σ of their IP)
};

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ u32 warp_incl_scan(u32 v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, v, d); if ((int)lane_id() >= d) v += t; }
  return v;
}

__device__ __forceinline__ u64 ld_status(const u64* p) {
  u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_status(u64* p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// Store / load a row pointer in the caller's width (32-bit ABI of the reference, or the _i64 variant).
__device__ __forceinline__ void st_rowptr(void* Crow, int is64, size_t i, u64 v, u32* err) {
  if (is64) ((long long*)Crow)[i] = (long long)v;
  else { if (v > 0x7fffffffull) atomicOr(err, 2u); ((int*)Crow)[i] = (int)v; }
}
__device__ __forceinline__ u64 ld_rowptr(const void* Crow, int is64, size_t i) {
  return is64 ? (u64)((const long long*)Crow)[i] : (u64)(u32)((const int*)Crow)[i];
}

// Exclusive prefix of this tile in the chain (decoupled look-back, executed by one full warp).
// status[] must be zero before the launch; tiles are handed out by an atomic counter, so every
// predecessor of a running tile has already started: the spin cannot deadlock.
__device__ __forceinline__ u64 lookback_exclusive(u64* status, u32 tile, u64 aggregate) {
  const u32 lane = lane_id();
  if (tile == 0) { if (lane == 0) st_status(&status[0], ST_INC | aggregate); return 0; }
  if (lane == 0) st_status(&status[tile], ST_AGG | aggregate);
  u64 excl = 0;
  long long idx = (long long)tile - 1;
  while (true) {
    long long my = idx - lane;
    u64 s;
    while (true) {
      s = (my >= 0) ? ld_status(&status[my]) : ST_INC;          // before tile 0: inclusive prefix 0
      if (!__any_sync(0xffffffffu, (s >> 62) == 0)) break;
      __nanosleep(64);
    }
    u32 inc_mask = __ballot_sync(0xffffffffu, (s >> 62) == 2);
    u32 first = inc_mask ? (u32)(__ffs(inc_mask) - 1) : 32u;     // nearest predecessor with a full prefix
    u64 v = (lane <= first) ? (s & ST_VAL) : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    excl += v;
    if (inc_mask) break;
    idx -= 32;
  }
  if (lane == 0) st_status(&status[tile], ST_INC | (excl + aggregate));
  return excl;
}

// In-place ascending bitonic sort of 32/S independent rows, K keys per lane, element index i = lane_in_row*K + k.
// "Flip" formulation: every merge level starts with the mirror exchange i <-> i ^ (size-1), then half-cleaners
// i <-> i ^ d; every comparator puts the minimum at the lower index, so exchanges inside a lane need no run-time
// direction (min + max), exchanges between lanes cost SHFL + min + predicated max.
// RUN: the keys arrive as ascending runs of RUN consecutive elements (1 = unsorted): the merge levels up to RUN are skipped.
// MIX = 1: two of three in-lane comparators form the maximum on the FMA pipe (hi = a + b - lo as two IMADs whose multipliers
// `one` = 1 and `mone` = -1 are run-time values, so that ptxas cannot fold them back into an ALU-pipe IADD3): 1 ALU + 2 FMA-pipe
// instructions instead of 2 ALU — for kernels whose limiter is the ALU pipe.
template <int MIX>
__device__ __forceinline__ void cmpx(u32& a, u32& b, const int c, const u32 one, const u32 mone) {
  const u32 lo = min(a, b);
  u32 hi;
  if (MIX == 1 && c % 3 != 0) {
    u32 s;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s) : "r"(a), "r"(one), "r"(b));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(lo), "r"(mone), "r"(s));
  } else {
    hi = max(a, b);
  }
  a = lo; b = hi;
}
template <int K, int S, int RUN = 1, int MIX = 0>
__device__ __forceinline__ void bitonic_sort_rows(u32 (&x)[K], const u32 ll, const u32 one = 1u, const u32 mone = 0xffffffffu) {
  constexpr int N = K * S;
#pragma unroll
  for (int size = 2 * RUN; size <= N; size <<= 1) {
    if (size <= K) {                                               // mirror inside the lane
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int pk = k ^ (size - 1);
        if (k < pk) cmpx<MIX>(x[k], x[pk], k, one, mone);
      }
    } else {                                                       // mirror across lanes: register k <-> K-1-k of lane ^ (size/K-1)
      const u32 lm = (u32)(size / K - 1);
      const bool keepmin = (ll & (u32)(size / (2 * K))) == 0u;
#pragma unroll
      for (int k = 0; k < K / 2; ++k) {
        const u32 ya = __shfl_xor_sync(0xffffffffu, x[K - 1 - k], lm);
        const u32 yb = __shfl_xor_sync(0xffffffffu, x[k], lm);
        x[k] = keepmin ? min(x[k], ya) : max(x[k], ya);
        x[K - 1 - k] = keepmin ? min(x[K - 1 - k], yb) : max(x[K - 1 - k], yb);
      }
    }
#pragma unroll
    for (int d = size >> 2; d >= 1; d >>= 1) {
      if (d >= K) {                                                // partner key lives in lane ^ (d/K)
        const u32 ld = (u32)(d / K);
        const bool keepmin = (ll & ld) == 0u;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const u32 y = __shfl_xor_sync(0xffffffffu, x[k], ld);
          x[k] = keepmin ? min(x[k], y) : max(x[k], y);
        }
      } else {                                                     // both keys in this lane
#pragma unroll
        for (int k = 0, c = 0; k < K; ++k)
          if ((k & d) == 0) { cmpx<MIX>(x[k], x[k | d], c, one, mone); ++c; }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ (1) work estimation
// ip[i] = Σ_{jj in A.row(i)} len(B.row(Acol[jj]))  — the trip count of final/SpGEMM_mpi_omp.c:33-37.
// G lanes per row (G chosen from the mean A row length); Acol is read coalesced, Brow is a gather
// (Brow is small and L2-resident).  Also builds the log2 histogram used to pick the bin thresholds.
template <int G>
__global__ void __launch_bounds__(256) k_estimate(Csr m, u32* __restrict__ ip, DevScalars* sc) {
  __shared__ u32 s_hist[34];
  __shared__ u64 s_sum;
  __shared__ u32 s_max;
  if (threadIdx.x < 34) s_hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_sum = 0; s_max = 0; }
  __syncthreads();
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = gtid / G;
  const int l = (int)(gtid % G);
  u64 sum = 0;
  u32 bad = 0;
  if (row < m.An) {
    const int a0 = m.Arow[row], a1 = m.Arow[row + 1];
    for (int jj = a0 + l; jj < a1; jj += G) {
      const int j = m.Acol[jj];
      if ((u32)j < (u32)m.Bn) sum += (u32)(m.Brow[j + 1] - m.Brow[j]); else bad = 1;
    }
  }
#pragma unroll
  for (int d = G / 2; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if (bad) atomicOr(&sc->err, 1u);
  if (row < m.An && l == 0) {
    const u32 v = sum > 0xfffffffeull ? 0xfffffffeu : (u32)sum;
    ip[row] = v;
    atomicAdd(&s_hist[v == 0 ? 0 : (v == 1 ? 1 : 33 - __clz(v - 1))], 1u);
    atomicAdd(&s_sum, sum);
    atomicMax(&s_max, v);
  }
  __syncthreads();
  if (threadIdx.x < 34 && s_hist[threadIdx.x]) atomicAdd(&sc->hist[threadIdx.x], s_hist[threadIdx.x]);
  if (threadIdx.x == 0) { if (s_sum) atomicAdd(&sc->total_ip, s_sum); atomicMax(&sc->max_ip, s_max); }
}

// Row lists for the CTA-per-row bins (order inside a list is irrelevant).  tofs (optional): every listed row also gets
// IP words (rounded up to 4) of the staging arena (MODE_STAGE), handed out with one 64-bit atomic per warp.
static __global__ void __launch_bounds__(256) k_build_lists(const u32* __restrict__ ip, int An, u32 cap_s, u32 cap_m1, u32 cap_m2,
                                                     u32* __restrict__ list_m1, u32* __restrict__ list_m2,
                                                     u32* __restrict__ list_l, u64* __restrict__ tofs, DevScalars* sc,
                                                     const int* __restrict__ Arow, u32 max_na_m2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const u32 v = (i < An) ? ip[i] : 0u;
  const bool big = v > cap_s;
  if (tofs) {
    const u32 mine = big ? ((v + 3u) & ~3u) : 0u;     // 16-byte aligned rows: k_copy_rows reads them with 16-byte loads
    u64 inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane_id() >= d) inc += t; }
    const u64 tot = __shfl_sync(0xffffffffu, inc, 31);
    u64 base = 0;
    if (lane_id() == 0 && tot) base = atomicAdd(&sc->temp_used, tot);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (big) tofs[i] = base + inc - mine;
  }
  if (!big) return;
  if (v <= cap_m1)      list_m1[atomicAdd(&sc->n_m1, 1u)] = (u32)i;
  else if (v <= cap_m2 && (u32)(Arow[i + 1] - Arow[i]) <= max_na_m2) list_m2[atomicAdd(&sc->n_m2, 1u)] = (u32)i;   // (rows_bm.cuh's register path: <= 1024 A entries)
  else                  list_l[atomicAdd(&sc->n_l, 1u)] = (u32)i;
}

// MODE_STAGE epilogue: the staged rows go to their final place, Ccol[Crow[row] ..) (Crow is complete once the scan / the fused
// kernel has run).  WPR: one warp per row (the M1 list, rows of <= 2048 columns), otherwise one 256-thread CTA per row, grid-stride.
// A staged row starts on a 16-byte boundary (k_build_lists), its destination anywhere: a warp reads 128 columns with one
// 16-byte load per lane (four blocks = 64 bytes per lane in flight), transposes them through shuffles (lane l of store k takes
// component l&3 of lane 8k + l/4) and writes four fully coalesced 128-byte stores.  With 4-byte loads the kernel sat at 58 % of
// the DRAM bandwidth with the load/store queue full (lg_throttle, profiles/r02_rmat20_copy_ncu_summary.txt).  The list entry two
// rows ahead and the (count, source, destination) of the next row are loaded before the current row is copied.
template <bool WPR>
static __global__ void __launch_bounds__(256) k_copy_rows(const u32* __restrict__ list, const u32* __restrict__ nlist, const u32* __restrict__ cnt,
                                                   const u64* __restrict__ tofs, const int* __restrict__ temp,
                                                   const void* __restrict__ Crow, int is64, int* __restrict__ Ccol) {
  constexpr u32 U = 4;                                  // 128-column blocks in flight per warp
  const u32 n = *nlist;
  const u32 lane = lane_id(), wid = threadIdx.x >> 5;
  const u32 first = WPR ? blockIdx.x * 8u + wid : blockIdx.x, step = WPR ? gridDim.x * 8u : gridDim.x;
  const u32 w0 = WPR ? 0u : wid, wstep = WPR ? 1u : 8u;
  if (first >= n) return;
  u32 r1 = (first + step < n) ? list[first + step] : 0u;
  u32 c; u64 so, dofs;
  { const u32 r0 = list[first]; c = cnt[r0]; so = tofs[r0]; dofs = ld_rowptr(Crow, is64, (size_t)r0); }
  const u32 srcl0 = lane >> 2, comp = lane & 3u;
  for (u32 idx = first; idx < n; idx += step) {
    const bool more = idx + step < n;
    const u32 r2 = (idx + 2u * step < n && more) ? list[idx + 2u * step] : 0u;
    u32 nc = 0; u64 nso = 0, ndo = 0;
    if (more) { nc = cnt[r1]; nso = tofs[r1]; ndo = ld_rowptr(Crow, is64, (size_t)r1); }
    const int4* __restrict__ src4 = reinterpret_cast<const int4*>(temp + so);
    int* __restrict__ dst = Ccol + dofs;
    const u32 nblk = (c + 127u) >> 7;
    for (u32 b = w0; b < nblk; b += wstep * U) {
      int4 v[U];
#pragma unroll
      for (u32 u = 0; u < U; ++u) {
        const u32 q = (b + u * wstep) * 32u + lane;     // 16-byte piece of the row; pieces that start beyond the row are not read
        v[u] = make_int4(0, 0, 0, 0);
        if (q * 4u < c) v[u] = __ldcs(src4 + q);
      }
#pragma unroll
      for (u32 u = 0; u < U; ++u) {
        const u32 bb = b + u * wstep;
        if (bb < nblk) {                                // warp-uniform
#pragma unroll
          for (u32 k = 0; k < 4; ++k) {
            const int sl = (int)(8u * k + srcl0);
            const int x0 = __shfl_sync(0xffffffffu, v[u].x, sl), x1 = __shfl_sync(0xffffffffu, v[u].y, sl);
            const int x2 = __shfl_sync(0xffffffffu, v[u].z, sl), x3 = __shfl_sync(0xffffffffu, v[u].w, sl);
            const u32 e = bb * 128u + 32u * k + lane;
            if (e < c) dst[e] = comp == 0u ? x0 : comp == 1u ? x1 : comp == 2u ? x2 : x3;
          }
        }
      }
    }
    c = nc; so = nso; dofs = ndo; r1 = r2;
  }
}

__device__ __forceinline__ u32 slot_scale(u32 T, u32 range) { return (u32)((((u64)T) << 32) / range); }

// Insert key x at/after slot s.  atomicMin keeps the smaller key in the slot; the larger one (the
// newcomer or the evicted resident) moves right.  Returns 1 if a new distinct key was added.
// Invariant on exit of all inserts: non-empty slots read left-to-right are strictly ascending.
__device__ __forceinline__ u32 ordered_insert(u32* tab, u32 s, u32 x, u32& max_slot) {
  while (true) {
    const u32 old = atomicMin(&tab[s], x);
    if (old == EMPTY) { max_slot = max(max_slot, s); return 1u; }
    if (old == x) return 0u;
    x = max(old, x);
    ++s;
  }
}

// ------------------------------------------------------------------------------------------------ (2a) S bin: one warp per row
// Table words per warp: the first attempt of the ordered table needs 2*cap+32 slots rounded up to the
// compaction geometry (128 * odd number of uint4 per lane).  cap=256 -> 640 words.
__host__ __device__ constexpr u32 tab_words(u32 cap) { return cap < 128u ? 384u : 128u * ((((2u * cap + 32u) + 127u) >> 7) | 1u); }

// What a built table looks like to the emitter.
struct RowTable {
  u32 lo;        // smallest candidate column (bitmap origin / slot map origin)
  u32 span;      // bitmap: number of words; ordered table: largest occupied slot
  u32 count;     // distinct columns
  u32 bitmap;    // 1 = bitmap over [lo,hi], 0 = ordered table
  u32 sorted;    // 1 = the distinct columns already stand ascending in the row's staging area (register sort)
  u32 ok;        // 0 = the optimistic build failed (bad lo/hi hint or spill overflow): rebuild exactly
};

__device__ __forceinline__ void table_init_empty(u32* tab, u32 words) {      // words: multiple of 128
  uint4* t4 = reinterpret_cast<uint4*>(tab);
  const uint4 e = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
  for (u32 q = lane_id(); q < words / 4; q += 32) t4[q] = e;
}

// Sort + de-duplicate the staged candidates of one row in registers (one warp, K keys per lane, 32*K >= ipr): the row's
// distinct columns end up ascending in stage[0..count).  The cost is a constant of the (padded) row size — it does not
// depend on how the columns are distributed, unlike an order-preserving slot map, which degenerates into long
// collision chains on power-law rows (R-MAT: most candidates of every row sit on the same few hub columns).
template <int K>
__device__ __noinline__ u32 sort_dedup_staged(u32* stage, const u32 ipr, const u32 Bm, u32* err) {
  const u32 lane = lane_id();
  u32 x[K];
  u32 vmax = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) {                       // any order will do: lane-interleaved = conflict-free
    const u32 i = (u32)k * 32u + lane;
    x[k] = EMPTY;
    if (i < ipr) { x[k] = stage[i]; vmax = max(vmax, x[k]); }
  }
  if (__reduce_max_sync(0xffffffffu, vmax) >= Bm) { if (lane == 0) atomicOr(err, 4u); return 0u; }   // B column outside [0,Bm): refuse
  __syncwarp();
  bitonic_sort_rows<K, 32, 1>(x, lane);
  u32 prev = __shfl_up_sync(0xffffffffu, x[K - 1], 1);
  if (lane == 0) prev = EMPTY;
  u32 c = 0, f = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) { const bool fk = (x[k] != EMPTY) && (x[k] != (k ? x[k - 1] : prev)); f |= fk ? (1u << k) : 0u; c += fk ? 1u : 0u; }
  const u32 inc = warp_incl_scan(c);
  u32 o = inc - c;
#pragma unroll
  for (int k = 0; k < K; ++k) if ((f >> k) & 1u) stage[o++] = x[k];
  __syncwarp();
  return __shfl_sync(0xffffffffu, inc, 31);
}

// Exact build from staged candidates stage[0..ipr): lo/hi by reduction, then
//   narrow span (hi-lo < 32*tabw): bitmap over [lo,hi] in tab[] (emit_sorted reads it back);
//   wide span: register sort (sort_dedup_staged): the sorted distinct row is left in stage[] (t.sorted = 1).
static __device__ __noinline__ RowTable build_staged(u32* stage, const u32 ipr, u32* tab, const u32 tabw, const u32 Bm, u32* err) {
  const u32 lane = lane_id();
  RowTable t; t.ok = 1; t.count = 0; t.span = 0; t.bitmap = 0; t.sorted = 0;
  u32 vmin = EMPTY, vmax = 0;
#pragma unroll 4
  for (u32 p = lane; p < ipr; p += 32) { const u32 v = stage[p]; vmin = min(vmin, v); vmax = max(vmax, v); }
  const u32 lo = __reduce_min_sync(0xffffffffu, vmin);
  const u32 hi = __reduce_max_sync(0xffffffffu, vmax);
  t.lo = lo;
  if (hi >= Bm) { if (lane == 0) atomicOr(err, 4u); return t; }      // B column outside [0,Bm): refuse (count 0)
  const u32 range = hi - lo + 1;
  if (range <= 32u * tabw) {
    const u32 nW = (range + 31) >> 5;
    for (u32 w = lane; w < nW; w += 32) tab[w] = 0;
    __syncwarp();
    u32 added = 0;
#pragma unroll 4
    for (u32 p = lane; p < ipr; p += 32) {
      const u32 v = stage[p] - lo, bit = 1u << (v & 31);
      const u32 old = atomicOr(&tab[v >> 5], bit);
      added += (old & bit) ? 0u : 1u;
    }
    __syncwarp();
    t.bitmap = 1; t.span = nW; t.count = __reduce_add_sync(0xffffffffu, added);
    return t;
  }
  t.sorted = 1;
  if (ipr <= 64u)       t.count = sort_dedup_staged<2>(stage, ipr, Bm, err);
  else if (ipr <= 128u) t.count = sort_dedup_staged<4>(stage, ipr, Bm, err);
  else if (ipr <= 256u) t.count = sort_dedup_staged<8>(stage, ipr, Bm, err);
  else if (ipr <= 512u) t.count = sort_dedup_staged<16>(stage, ipr, Bm, err);
  else                  t.count = sort_dedup_staged<32>(stage, ipr, Bm, err);     // cap_s <= 1024
  return t;
}

// Sorted distinct columns of a built table -> out[0..count).  Bitmap: popc/ffs per word.  Ordered table:
// lane l owns 4*S consecutive slots (S odd -> conflict-free LDS.128), one warp scan gives its offset.
static __device__ __noinline__ void emit_sorted(const u32* tab, const RowTable t, u32* out) {
  const u32 lane = lane_id();
  if (t.bitmap) {
    u32 cnt = 0;
    for (u32 w0 = 0; w0 < t.span; w0 += 32) {
      const u32 w = w0 + lane;
      u32 word = (w < t.span) ? tab[w] : 0u;
      const u32 c = __popc(word);
      const u32 inc = warp_incl_scan(c);
      u32 o = cnt + inc - c;
      const u32 base = t.lo + (w << 5);
      while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; out[o++] = base + b; }
      cnt += __shfl_sync(0xffffffffu, inc, 31);
    }
  } else {
    const u32 S = ((t.span + 128) >> 7) | 1u;       // covers slots [0, span], inside the initialised region
    const uint4* mine = reinterpret_cast<const uint4*>(tab) + lane * S;
    u32 c = 0;
    for (u32 q = 0; q < S; ++q) { const uint4 v = mine[q]; c += (v.x != EMPTY) + (v.y != EMPTY) + (v.z != EMPTY) + (v.w != EMPTY); }
    const u32 inc = warp_incl_scan(c);
    u32 o = inc - c;
    for (u32 q = 0; q < S; ++q) {
      const uint4 v = mine[q];
      if (v.x != EMPTY) out[o++] = v.x;
      if (v.y != EMPTY) out[o++] = v.y;
      if (v.z != EMPTY) out[o++] = v.z;
      if (v.w != EMPTY) out[o++] = v.w;
    }
  }
  __syncwarp();
}

// Gather the candidate columns of one row into stage[] with the whole warp (any row length, any B row
// length): batches of 32 A nonzeros, G lanes walk one B row, positions from a warp scan.  Returns the
// number gathered (= IP of the row); gathers nothing and returns the IP if it exceeds `cap`.
template <int G>
__device__ __noinline__ u32 gather_row(const Csr& m, int row, u32 cap, u32* stage) {
  const u32 lane = lane_id();
  constexpr int SPW = 32 / G;
  const u32 sub = lane / G, off0 = lane % G;
  const int a0 = m.Arow[row], a1 = m.Arow[row + 1];
  u64 total = 0;
  for (int b0 = a0; b0 < a1; b0 += 32) {            // pass 1: IP of the row
    const int jj = b0 + (int)lane;
    u32 len = 0;
    if (jj < a1) { const int j = m.Acol[jj]; if ((u32)j < (u32)m.Bn) len = (u32)(m.Brow[j + 1] - m.Brow[j]); }
    total += __reduce_add_sync(0xffffffffu, len);
  }
  if (total > cap) return total > 0xfffffffeull ? 0xfffffffeu : (u32)total;
  u32 pos_base = 0;
  for (int b0 = a0; b0 < a1; b0 += 32) {
    const int jj = b0 + (int)lane;
    u32 bs = 0, len = 0;
    if (jj < a1) {
      const int j = m.Acol[jj];
      if ((u32)j < (u32)m.Bn) { bs = (u32)m.Brow[j]; len = (u32)m.Brow[j + 1] - bs; }
    }
    const u32 incl = warp_incl_scan(len);
    const u32 excl = incl - len;
    const u32 tot = __shfl_sync(0xffffffffu, incl, 31);
    const int nseg = min(32, a1 - b0);
    for (int s = 0; s < nseg; s += SPW) {
      const int src = s + (int)sub;                     // lanes >= nseg carry len 0
      const u32 sbs = __shfl_sync(0xffffffffu, bs, src & 31);
      const u32 slen = __shfl_sync(0xffffffffu, len, src & 31);
      const u32 spos = __shfl_sync(0xffffffffu, excl, src & 31);
      for (u32 o = off0; o < slen; o += G) stage[pos_base + spos + o] = (u32)__ldg(&m.Bcol[sbs + o]);
    }
    pos_base += tot;
  }
  __syncwarp();
  return pos_base;
}

// Two-phase kernels of the S bin (grid-stride, one warp per row; rows with IP > cap belong to other bins).
//   MODE_COUNT: cnt[row] = nnz(C_row).   MODE_FILL: Ccol[Crow[row] ...] = sorted distinct columns.
template <int G, int MODE>
__global__ void __launch_bounds__(WARPS_S * 32) k_rows_warp(Csr m, const u32* __restrict__ ip, u32* __restrict__ cnt, u32 cap,
                                                           const void* __restrict__ Crow, int is64, int* __restrict__ Ccol,
                                                           DevScalars* sc, u32 tiny_max) {
  extern __shared__ __align__(16) u32 smem[];
  const u32 warp = threadIdx.x >> 5, lane = lane_id();
  u32* tab = smem + (size_t)warp * (tab_words(cap) + cap);
  u32* stage = tab + tab_words(cap);
  const long long nw = (long long)gridDim.x * WARPS_S;
  for (long long row = (long long)blockIdx.x * WARPS_S + warp; row < m.An; row += nw) {
    const u32 ipr = ip[row];
    if (ipr > cap || ipr <= tiny_max) continue;        // (rows of up to tiny_max products: k_rows_tiny; tiny_max = 0: none, not even empty rows)
    u32 c = 0;
    if (ipr) {
      gather_row<G>(m, (int)row, cap, stage);
      const RowTable t = build_staged(stage, ipr, tab, tab_words(cap), (u32)m.Bm, &sc->err);
      c = t.count;
      if (MODE == MODE_FILL && c) {
        if (!t.sorted) emit_sorted(tab, t, stage);
        const u64 base = ld_rowptr(Crow, is64, (size_t)row);
        for (u32 p = lane; p < c; p += 32) Ccol[base + p] = (int)stage[p];
      }
    }
    if (MODE == MODE_COUNT && lane == 0) cnt[row] = c;
    __syncwarp();
  }
}

// Rows of at most TINY_MAX (64) intermediate products — all of a sprand(d=5) matrix (Poisson rows, 25 products on average), the
// small half of every power-law matrix: one warp per row, the row's candidates in ONE or TWO registers per lane, no table.
// k_rows_warp spends 536 warp instructions on such a row (table set-up, a 128-key network) at 80 registers = 24 warps per SM
// (profiles/r02_pub_rows_warp_ncu_summary.txt); here: lane per A entry -> warp scan of the B-row lengths -> lane per product
// (binary search over the scan through shuffles) -> bitonic network over the lanes (15 shuffle stages for <= 32 keys, 21 for
// <= 64) -> neighbour compare + ballot -> count / ranked stores.  Replaces the flag array + quickSort of SpGEMM_bigslice
// (final/SpGEMM_mpi_omp.c:21, 33-47) for these rows like the other bins do.
constexpr u32 TINY_MAX = 64;
template <int NR>
__device__ __forceinline__ void tiny_sort(u32 (&k)[NR], const u32 lane) {       // element i = r * 32 + lane, ascending
#pragma unroll
  for (u32 kk = 2; kk <= 32u * NR; kk <<= 1) {
#pragma unroll
    for (u32 j = kk >> 1; j > 0; j >>= 1) {
      if (j >= 32u) {                                  // partner in the other register of the same lane (NR == 2, kk == 64: ascending)
        const u32 lo = min(k[0], k[NR - 1]), hi = max(k[0], k[NR - 1]);
        k[0] = lo; k[NR - 1] = hi;
      } else {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const u32 v = __shfl_xor_sync(0xffffffffu, k[r], (int)j);
          const u32 i = (u32)r * 32u + lane;
          const bool up = (i & kk) == 0u, lower = (lane & j) == 0u;
          k[r] = (lower == up) ? min(k[r], v) : max(k[r], v);
        }
      }
    }
  }
}
template <int NR, int MODE>
__device__ __forceinline__ u32 tiny_finish(u32 (&k)[NR], const u32 lane, const u32 Bm, int* __restrict__ dst, u32* bad) {
  tiny_sort<NR>(k, lane);
  u32 c = 0, before = 0;
  const u32 last0 = __shfl_sync(0xffffffffu, k[0], 31);          // the element in front of register 1's lane 0
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    u32 prev = __shfl_up_sync(0xffffffffu, k[r], 1);
    if (lane == 0) prev = (r == 0) ? EMPTY : last0;
    if (k[r] != EMPTY && k[r] >= Bm) { *bad = 1; }
    const bool keep = k[r] != EMPTY && k[r] < Bm && k[r] != prev;
    const u32 mask = __ballot_sync(0xffffffffu, keep);
    if (MODE == MODE_FILL && keep) dst[before + __popc(mask & ((1u << lane) - 1u))] = (int)k[r];
    before += __popc(mask);
    c = before;
  }
  return c;
}
template <int MODE>
__global__ void __launch_bounds__(256) k_rows_tiny(Csr m, const u32* __restrict__ ip, u32* __restrict__ cnt,
                                                   const void* __restrict__ Crow, int is64, int* __restrict__ Ccol, DevScalars* sc,
                                                   u32 tiny_max) {                     // <= TINY_MAX; rows above it belong to the other bins
  __shared__ u32 st[8][TINY_MAX];
  const u32 warp = threadIdx.x >> 5, lane = lane_id();
  const long long nw = (long long)gridDim.x * 8;
  u32 bad = 0;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < m.An; row += nw) {
    const u32 ipr = ip[row];
    if (ipr > tiny_max) continue;
    u32 c = 0;
    if (ipr) {
      const int a0 = m.Arow[row], a1 = m.Arow[row + 1];
      u32 filled = 0;
      for (int b0 = a0; b0 < a1; b0 += 32) {           // 32 A entries at a time (one pass unless the row selects empty B rows galore)
        const int jj = b0 + (int)lane;
        u32 bs = 0, len = 0;
        if (jj < a1) { const int j = m.Acol[jj]; if ((u32)j < (u32)m.Bn) { bs = (u32)m.Brow[j]; len = (u32)m.Brow[j + 1] - bs; } }
        const u32 incl = warp_incl_scan(len);
        const u32 tot = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
        for (u32 h = 0; h < 2; ++h) {
          if (h * 32u >= tot) break;                   // warp-uniform
          const u32 q = h * 32u + lane;                // product q of this chunk: entry e = number of lanes whose scan is <= q
          u32 e = 0;
#pragma unroll
          for (u32 step = 16; step; step >>= 1) { const u32 v = __shfl_sync(0xffffffffu, incl, (int)(e + step - 1u)); if (v <= q) e += step; }
          const u32 e_bs = __shfl_sync(0xffffffffu, bs, (int)(e & 31u)), e_excl = __shfl_sync(0xffffffffu, incl - len, (int)(e & 31u));
          if (q < tot && filled + q < TINY_MAX) st[warp][filled + q] = (u32)__ldg(&m.Bcol[e_bs + (q - e_excl)]);
        }
        filled += tot;
      }
      __syncwarp();
      int* dst = (MODE == MODE_FILL) ? Ccol + ld_rowptr(Crow, is64, (size_t)row) : nullptr;
      if (ipr <= 32u) {
        u32 k[1] = { lane < ipr ? st[warp][lane] : EMPTY };
        c = tiny_finish<1, MODE>(k, lane, (u32)m.Bm, dst, &bad);
      } else {
        u32 k[2] = { st[warp][lane], lane + 32u < ipr ? st[warp][lane + 32u] : EMPTY };
        c = tiny_finish<2, MODE>(k, lane, (u32)m.Bm, dst, &bad);
      }
      __syncwarp();
    }
    if (MODE == MODE_COUNT && lane == 0) cnt[row] = c;
  }
  if (bad) atomicOr(&sc->err, 4u);
}

// ------------------------------------------------------------------------------------------------ fused one-pass kernel
// Every WARP is an independent worker: it takes tiles of R=4 consecutive rows from an atomic counter (the
// next tile id is fetched one tile ahead), and there is no __syncthreads anywhere.  Per tile:
//   1. one coalesced load of the R+1 row pointers, one of the tile's contiguous Acol range (<= EMAX entries),
//      one gather of the Brow pairs -> (start,len) descriptors in shared memory, plus the first and last
//      column of every selected B row: 3 dependent round trips per R rows;
//   2. a warp scan of the lengths gives every row's IP; rows with IP > cap are "big": their count comes from
//      cnt_big[] (written beforehand by the M/L symbolic kernels) and they are filled afterwards;
//   3. per row, OPTIMISTIC build straight from the gather registers: [lo,hi] is taken from the B rows' first
//      and last entries (exact when B rows are sorted, which nothing guarantees), every gathered column is
//      validated against it, first probes are issued 4 at a time (independent ATOMS.MIN), colliding keys go
//      to a shared-memory queue that is drained with all lanes busy.  A failed validation or a spill past
//      the 32 spare slots falls back to the exact staged build (gather_row + build_staged);
//   4. emit_sorted into the row's staging area; 5. the tile's 4 counts enter the look-back chain (one status
//      word per tile, the warp's 32 lanes inspect 32 predecessors at once), Crow is written;
//   6. the sorted rows are streamed to their final position in Ccol.
// B is gathered once, C is written once, IP is never materialised.
constexpr int FUSED_R = 4, FUSED_EPL = 2, FUSED_EMAX = 32 * FUSED_EPL;
__host__ __device__ constexpr u32 fused_warp_words(u32 cap) { return tab_words(cap) + FUSED_R * cap + 2 * FUSED_EMAX + 4; }

template <int G, bool NOTAIL, bool BITMAP>
__device__ __forceinline__ void direct_insert_loop(const Csr& m, const uint2* desc, int e0, int e1, u32* tab, u32* queue,
                                                   u32 lo, u32 range, u32 scale, u32& added, u32& max_slot, u32& bad, u32* qn) {
  constexpr int SPW = 32 / G;
  const u32 lane = lane_id(), sub = lane / G, off0 = lane % G;
  for (int q = e0; q < e1; q += 4 * SPW) {
    u32 bs[4], len[4], v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int seg = q + u * SPW + (int)sub;
      const uint2 d = (seg < e1) ? desc[seg] : make_uint2(0u, 0u);
      bs[u] = d.x; len[u] = d.y;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (off0 < len[u]) ? (u32)__ldg(&m.Bcol[bs[u] + off0]) : EMPTY;
    if (BITMAP) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const u32 d = v[u] - lo;
        if (v[u] != EMPTY) { if (d < range) { const u32 bit = 1u << (d & 31); const u32 old = atomicOr(&tab[d >> 5], bit); added += (old & bit) ? 0u : 1u; } else bad = 1; }
      }
    } else {
      u32 s[4], old[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const u32 d = v[u] - lo; if (v[u] != EMPTY && d >= range) { bad = 1; v[u] = EMPTY; } s[u] = __umulhi(d, scale); }
#pragma unroll
      for (int u = 0; u < 4; ++u) old[u] = (v[u] != EMPTY) ? atomicMin(&tab[s[u]], v[u]) : v[u];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (v[u] == EMPTY) continue;
        if (old[u] == EMPTY) { ++added; max_slot = max(max_slot, s[u]); }
        else if (old[u] != v[u]) queue[atomicAdd(qn, 1u)] = max(old[u], v[u]);   // the larger key re-enters from its home slot
      }
    }
    if (!NOTAIL) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        for (u32 o = off0 + G; o < len[u]; o += G) {      // B rows longer than G
          const u32 x = (u32)__ldg(&m.Bcol[bs[u] + o]);
          const u32 d = x - lo;
          if (d >= range) { bad = 1; continue; }
          if (BITMAP) { const u32 bit = 1u << (d & 31); const u32 old = atomicOr(&tab[d >> 5], bit); added += (old & bit) ? 0u : 1u; }
          else queue[atomicAdd(qn, 1u)] = x;
        }
    }
  }
}

template <int G, bool NOTAIL>
__global__ void __launch_bounds__(1024, 1) k_fused(Csr m, const u32* __restrict__ cnt_big, u32 cap,
                                                  void* __restrict__ Crow, int is64, int* __restrict__ Ccol,
                                                  u64* __restrict__ status, DevScalars* sc, u32 ntiles, int acc_ip) {
  constexpr int R = FUSED_R, EPL = FUSED_EPL, EMAX = FUSED_EMAX;
  extern __shared__ __align__(16) u32 smem[];
  const u32 warp = threadIdx.x >> 5, lane = lane_id();
  const u32 tabw = tab_words(cap);
  u32* tab = smem + (size_t)warp * fused_warp_words(cap);
  u32* stage = tab + tabw;
  uint2* desc = reinterpret_cast<uint2*>(stage + R * cap);   // (start, length) of the B row each A nonzero selects
  u32* pfx = tab;                                             // EMAX+1 prefix of the lengths; dead before the table is used
  u32* qn = reinterpret_cast<u32*>(desc + EMAX);              // collision-queue length
  u64 ip_sum = 0;
  u32 ip_max = 0;

  u32 tile = 0;
  if (lane == 0) tile = atomicAdd(&sc->tile_counter, 1u);
  tile = __shfl_sync(0xffffffffu, tile, 0);
  while (tile < ntiles) {
    u32 next = 0;
    if (lane == 0) next = atomicAdd(&sc->tile_counter, 1u);   // consumed at the end of this tile
    const long long row0 = (long long)tile * R;
    const int nrows = (int)min((long long)R, (long long)m.An - row0);
    const int ar = m.Arow[row0 + min((int)lane, nrows)];
    int a[R + 1];
#pragma unroll
    for (int r = 0; r <= R; ++r) a[r] = __shfl_sync(0xffffffffu, ar, r);
    const int E = a[R] - a[0];
    u32 S[R + 1], lo_r[R], hi_r[R];
    const bool fast = E <= EMAX;
    if (fast) {
      u32 run = 0;
      u32 first[EPL], last[EPL], len[EPL];
#pragma unroll
      for (int k = 0; k < EPL; ++k) {
        const int e = k * 32 + (int)lane;
        u32 bs = 0; len[k] = 0;
        if (e < E) {
          const int j = m.Acol[a[0] + e];
          if ((u32)j < (u32)m.Bn) { bs = (u32)m.Brow[j]; len[k] = (u32)m.Brow[j + 1] - bs; } else atomicOr(&sc->err, 1u);
        }
        desc[e] = make_uint2(bs, len[k]);
        first[k] = len[k] ? (u32)__ldg(&m.Bcol[bs]) : EMPTY;
        last[k]  = len[k] ? (u32)__ldg(&m.Bcol[bs + len[k] - 1]) : 0u;
      }
#pragma unroll
      for (int k = 0; k < EPL; ++k) {
        const u32 inc = warp_incl_scan(len[k]);
        pfx[k * 32 + lane] = run + inc - len[k];
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r <= R; ++r) { const int e = a[r] - a[0]; S[r] = (e >= E) ? run : pfx[e]; }
#pragma unroll
      for (int r = 0; r < R; ++r) {          // [lo,hi] hint of every row from its B rows' end points
        u32 mn = EMPTY, mx = 0;
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
          const int e = k * 32 + (int)lane;
          if (len[k] && e >= a[r] - a[0] && e < a[r + 1] - a[0]) { mn = min(mn, min(first[k], last[k])); mx = max(mx, max(first[k], last[k])); }
        }
        lo_r[r] = __reduce_min_sync(0xffffffffu, mn);
        hi_r[r] = __reduce_max_sync(0xffffffffu, mx);
      }
      __syncwarp();
    }
    u32 c_mine = 0;                               // lane r keeps the count of row r
    u32 mine_mask = 0;
#pragma unroll 1
    for (int r = 0; r < nrows; ++r) {
      u32 c = 0;
      u32* out = stage + r * cap;
      bool need_exact = !fast;
      u32 ipr = 0;
      RowTable t; t.ok = 0; t.count = 0; t.bitmap = 0; t.sorted = 0; t.lo = 0; t.span = 0;
      if (fast) {
        u32 Sa = S[0], Sb = S[1], lo = lo_r[0], hi = hi_r[0]; int e0 = a[0], e1 = a[1];
#pragma unroll
        for (int q = 1; q < R; ++q) if (r == q) { Sa = S[q]; Sb = S[q + 1]; lo = lo_r[q]; hi = hi_r[q]; e0 = a[q]; e1 = a[q + 1]; }
        e0 -= a[0]; e1 -= a[0];
        ipr = Sb - Sa;
        if (ipr > cap || ipr == 0) need_exact = false;
        else if (hi >= (u32)m.Bm || hi < lo) need_exact = true;
        else {
          const u32 range = hi - lo + 1;
          u32 added = 0, max_slot = 0, bad = 0;
          t.lo = lo;
          if (range <= 32u * tabw) {
            const u32 nW = (range + 31) >> 5;
            for (u32 w = lane; w < nW; w += 32) tab[w] = 0;
            __syncwarp();
            direct_insert_loop<G, NOTAIL, true>(m, desc, e0, e1, tab, out, lo, range, 0u, added, max_slot, bad, qn);
            __syncwarp();
            t.bitmap = 1; t.span = nW;
          } else bad = 1;                                   // wide span: staged gather + register sort (build_staged)
          t.count = __reduce_add_sync(0xffffffffu, added);
          t.ok = __any_sync(0xffffffffu, bad) ? 0u : 1u;
          need_exact = !t.ok;
        }
      }
      if (need_exact) {
        ipr = gather_row<G>(m, (int)(row0 + r), cap, out);
        t.ok = 0;
        if (ipr <= cap && ipr > 0) { t = build_staged(out, ipr, tab, tabw, (u32)m.Bm, &sc->err); }
      }
      if (!fast) { ip_sum += ipr; ip_max = max(ip_max, ipr); }
      if (ipr > cap) c = cnt_big[row0 + r];
      else if (ipr > 0 && t.ok && t.count) { if (!t.sorted) emit_sorted(tab, t, out); c = t.count; mine_mask |= 1u << r; }
      if ((int)lane == r) c_mine = c;
    }
    if (fast) { ip_sum += S[R] - S[0]; 
#pragma unroll
      for (int r = 0; r < R; ++r) ip_max = max(ip_max, S[r + 1] - S[r]); }

    // ---- chain the tile into the scan, write row pointers, stream the rows out
    const u32 v = ((int)lane < nrows) ? c_mine : 0u;
    const u32 inc = warp_incl_scan(v);
    const u64 agg = __shfl_sync(0xffffffffu, inc, 31);
    const u64 excl = lookback_exclusive(status, tile, agg);
    if ((int)lane < nrows) st_rowptr(Crow, is64, (size_t)(row0 + lane) + 1, excl + inc, &sc->err);
    if (tile == 0 && lane == 0) st_rowptr(Crow, is64, 0, 0, &sc->err);
    if (tile == ntiles - 1 && lane == 0) sc->total_nnz = excl + agg;
#pragma unroll 1
    for (int r = 0; r < nrows; ++r) {
      const u32 cr = __shfl_sync(0xffffffffu, v, r);
      const u32 off = __shfl_sync(0xffffffffu, inc - v, r);
      if (!((mine_mask >> r) & 1u)) continue;
      const u32* src = stage + r * cap;
      int* dst = Ccol + (excl + off);
      for (u32 p = lane; p < cr; p += 32) dst[p] = (int)src[p];
    }
    __syncwarp();
    tile = __shfl_sync(0xffffffffu, next, 0);
  }
  if (acc_ip && lane == 0) {   // per-lane copies are identical: lane 0 publishes
    if (ip_sum) atomicAdd(&sc->total_ip, ip_sum);
    atomicMax(&sc->max_ip, ip_max);
  }
}

// Longest row of A and of B (two streaming passes over the row pointers).  If maxA*maxB <= the S-bin capacity
// no row can leave the S bin and Σip <= nnzA*maxB, so the work-estimation pass can be skipped entirely.
static __global__ void __launch_bounds__(256) k_maxlen(const int* __restrict__ Arow, int An, const int* __restrict__ Brow, int Bn, DevScalars* sc) {
  __shared__ u32 s_a, s_b;
  if (threadIdx.x == 0) { s_a = 0; s_b = 0; }
  __syncthreads();
  u32 la = 0, lb = 0;
  const long long n = max(An, Bn), step = (long long)gridDim.x * blockDim.x;
  const bool same = Arow == Brow && An == Bn;                     // C = A*A (the reference's drivers): one pass
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {   // grid-stride: one global atomic pair per CTA
    if (i < An) la = max(la, (u32)(Arow[i + 1] - Arow[i]));
    if (!same && i < Bn) lb = max(lb, (u32)(Brow[i + 1] - Brow[i]));
  }
  if (same) lb = la;
  la = __reduce_max_sync(0xffffffffu, la);
  lb = __reduce_max_sync(0xffffffffu, lb);
  if (lane_id() == 0) { if (la) atomicMax(&s_a, la); if (lb) atomicMax(&s_b, lb); }
  __syncthreads();
  if (threadIdx.x == 0) { if (s_a) atomicMax(&sc->max_len_a, s_a); if (s_b) atomicMax(&sc->max_len_b, s_b); }
}

// Column-span probe: one warp per sampled row of A gathers the row's candidate columns and records whether they fit a
// window of 2^15 columns (banded / block-diagonal rows: the bitmap path of k_rows_warp / k_fused is the right tool, the
// global slot map of k_fused_ell would send every key of such a row to the same few slots).
static __global__ void __launch_bounds__(256) k_probe_span(Csr m, int nsamples, DevScalars* sc) {
  const int s = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (s >= nsamples) return;
  const long long row = (long long)m.An * s / nsamples;
  const u32 lane = lane_id();
  u32 lo = EMPTY, hi = 0, runs = 1;
  const int a0 = m.Arow[row], a1 = m.Arow[row + 1];
  // lane l takes A nonzeros l, l+32 (the first 64 are enough for a verdict) and walks the first 256 entries of their B rows:
  // three dependent loads per lane instead of three per A nonzero (one lane per entry cost 28 us per call at config 3)
  for (int jj = a0 + (int)lane; jj < a1 && jj < a0 + 64; jj += 32) {
    const int j = m.Acol[jj];
    if ((u32)j >= (u32)m.Bn) continue;
    const int b0 = m.Brow[j], b1 = m.Brow[j + 1];
    const u32 first = b1 > b0 ? (u32)m.Bcol[b0] : 0u;
    for (int o = b0; o < b1 && o < b0 + 256; ++o) {
      const u32 v = (u32)m.Bcol[o]; lo = min(lo, v); hi = max(hi, v);
      if (v != first + (u32)(o - b0)) runs = 0;
    }
  }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  runs = __all_sync(0xffffffffu, runs);
  if (lane == 0 && lo != EMPTY) {
    atomicAdd(&sc->span_rows, 1u);
    if (hi - lo < (1u << 15)) atomicAdd(&sc->span_narrow, 1u);
    if (runs && hi - lo < 128u) atomicAdd(&sc->span_runs, 1u);       // candidate for the run/bitmap kernel (band.cuh)
  }
}

// ------------------------------------------------------------------------------------------------ (2b) M bin: one CTA per row
// Walk all candidate columns of a row with the whole CTA: groups of G lanes take one A nonzero each.
template <class F>
__device__ __forceinline__ void cta_for_each_product(const Csr& m, int a0, int a1, int G, F f) {
  const int ngroups = (int)blockDim.x / G, g = (int)threadIdx.x / G, l = (int)threadIdx.x % G;
  for (int jj = a0 + g; jj < a1; jj += ngroups) {
    const int j = m.Acol[jj];
    if ((u32)j >= (u32)m.Bn) continue;
    const int bs = m.Brow[j], be = m.Brow[j + 1];
    for (int o = bs + l; o < be; o += G) f((u32)__ldg(&m.Bcol[o]));
  }
}

__device__ __forceinline__ u32 block_reduce_add(u32 v, u32* s_red) {   // s_red: 33 words
  v = __reduce_add_sync(0xffffffffu, v);
  if (lane_id() == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    u32 w = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : 0u;
    w = __reduce_add_sync(0xffffffffu, w);
    if (threadIdx.x == 0) s_red[32] = w;
  }
  __syncthreads();
  const u32 r = s_red[32];
  __syncthreads();
  return r;
}

// Block-wide exclusive scan of one value per thread; returns the exclusive prefix, total in *total.
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32* s_red, u32* total) {
  const u32 inc = warp_incl_scan(v);
  const u32 w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane_id() == 31) s_red[w] = inc;
  __syncthreads();
  if (threadIdx.x < 32) {
    const u32 x = (threadIdx.x < nw) ? s_red[threadIdx.x] : 0u;
    const u32 xi = warp_incl_scan(x);
    s_red[threadIdx.x] = xi - x;
    if (threadIdx.x == 31) s_red[32] = xi;
  }
  __syncthreads();
  const u32 r = s_red[w] + inc - v;
  *total = s_red[32];
  __syncthreads();
  return r;
}

// ------------------------------------------------------------------------------------------------ (2c) L bin: global bitmap
// One CTA per row; bitmap over [0,Bm) in global memory (one per resident CTA, zero between rows).
template <int MODE>
__global__ void __launch_bounds__(1024) k_rows_gbitmap(Csr m, const u32* __restrict__ list, const u32* __restrict__ nlist,
                                                       u32* __restrict__ cnt, int G, u32* __restrict__ bitmaps, u32 words_per_cta,
                                                       const void* __restrict__ Crow, int is64, int* __restrict__ Ccol,
                                                       DevScalars* sc) {
  __shared__ u32 s_red[33];
  __shared__ u32 s_wlo, s_whi, s_bad;
  u32* bm = bitmaps + (size_t)blockIdx.x * words_per_cta;
  const u32 n = *nlist;
  for (u32 idx = blockIdx.x; idx < n; idx += gridDim.x) {
    const int row = (int)list[idx];
    const int a0 = m.Arow[row], a1 = m.Arow[row + 1];
    if (threadIdx.x == 0) { s_wlo = EMPTY; s_whi = 0; s_bad = 0; }
    __syncthreads();
    u32 added = 0, wlo = EMPTY, whi = 0, bad = 0;
    cta_for_each_product(m, a0, a1, G, [&](u32 v) {
      if (v >= (u32)m.Bm) { bad = 1; return; }
      const u32 w = v >> 5, bit = 1u << (v & 31);
      const u32 old = atomicOr(&bm[w], bit);
      added += (old & bit) ? 0u : 1u;
      wlo = min(wlo, w); whi = max(whi, w);
    });
    wlo = __reduce_min_sync(0xffffffffu, wlo);
    whi = __reduce_max_sync(0xffffffffu, whi);
    if (lane_id() == 0) { atomicMin(&s_wlo, wlo); atomicMax(&s_whi, whi); if (bad) s_bad = 1; }
    if (__any_sync(0xffffffffu, bad) && lane_id() == 0) s_bad = 1;
    __syncthreads();
    if (s_bad && threadIdx.x == 0) atomicOr(&sc->err, 4u);
    const u32 w_lo = s_wlo, w_hi = s_whi;
    if (MODE == MODE_COUNT) {
      const u32 c = block_reduce_add(added, s_red);
      if (threadIdx.x == 0) cnt[row] = c;
    } else if (w_lo != EMPTY) {
      const u64 base = ld_rowptr(Crow, is64, (size_t)row);
      u64 done = 0;
      for (u32 w0 = w_lo; w0 <= w_hi; w0 += blockDim.x) {
        const u32 w = w0 + threadIdx.x;
        u32 word = (w <= w_hi) ? __ldcg(&bm[w]) : 0u, tot;
        u64 o = done + block_excl_scan(__popc(word), s_red, &tot);
        const u32 b0 = w << 5;
        while (word) { const u32 b = __ffs(word) - 1; word &= word - 1; Ccol[base + o++] = (int)(b0 + b); }
        done += tot;
        if (w0 + blockDim.x < w0) break;   // overflow guard
      }
    }
    __syncthreads();
    if (w_lo != EMPTY) for (u32 w = w_lo + threadIdx.x; w <= w_hi; w += blockDim.x) bm[w] = 0;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ (3) device prefix scan
// Crow[i+1] = Σ_{r<=i} cnt[r], Crow[0] = 0 — single pass, decoupled look-back over tiles of
// SCAN_THREADS*SCAN_ITEMS counts.  Replaces the serial fix-up loops (final/SpGEMM_mpi_omp.c:135-141, :215-221).
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8;
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan(const u32* __restrict__ cnt, int An, void* __restrict__ Crow, int is64,
                                                       u64* __restrict__ status, DevScalars* sc, u32 ntiles) {
  __shared__ u32 s_tile;
  __shared__ u64 s_wsum[SCAN_THREADS / 32];
  __shared__ u64 s_excl;
  if (threadIdx.x == 0) s_tile = atomicAdd(&sc->tile_counter, 1u);
  __syncthreads();
  const u32 tile = s_tile;
  if (tile >= ntiles) return;
  const long long base = (long long)tile * (SCAN_THREADS * SCAN_ITEMS) + (long long)threadIdx.x * SCAN_ITEMS;
  u32 v[SCAN_ITEMS];
  u64 tsum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < An) ? cnt[base + k] : 0u; tsum += v[k]; }
  // warp scan of thread sums (64-bit)
  u64 inc = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { u64 t = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane_id() >= d) inc += t; }
  const u32 w = threadIdx.x >> 5;
  if (lane_id() == 31) s_wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    u64 x = (lane_id() < SCAN_THREADS / 32) ? s_wsum[lane_id()] : 0;
    u64 xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u64 t = __shfl_up_sync(0xffffffffu, xi, d); if ((int)lane_id() >= d) xi += t; }
    const u64 agg = __shfl_sync(0xffffffffu, xi, 31);
    if (lane_id() < SCAN_THREADS / 32) s_wsum[lane_id()] = xi - x;
    const u64 excl = lookback_exclusive(status, tile, agg);
    if (lane_id() == 0) { s_excl = excl; if (tile == ntiles - 1) sc->total_nnz = excl + agg; if (tile == 0) st_rowptr(Crow, is64, 0, 0, &sc->err); }
  }
  __syncthreads();
  u64 run = s_excl + s_wsum[w] + inc - tsum;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    run += v[k];
    if (base + k < An) st_rowptr(Crow, is64, (size_t)(base + k) + 1, run, &sc->err);
  }
}

// ------------------------------------------------------------------------------------------------ multi-GPU helper
// Adds a shard's displacement to its slice-relative row pointers while copying them into the gathered
// array (replaces the root's fix-up loop final/SpGEMM_mpi_omp.c:211-223).
static __global__ void k_offset_rowptr(const void* __restrict__ src, void* __restrict__ dst, int is64, long long n, long long disp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (is64) ((long long*)dst)[i] = ((const long long*)src)[i] + disp;
  else ((int*)dst)[i] = (int)(((const int*)src)[i] + disp);
}

}  // namespace bsk
