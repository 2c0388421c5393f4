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
// The scan that gives every tile its output offset must not stall the row work: with the rows written in order,
// any warp-granular look-back makes every warp wait for the slowest earlier tile (profiles/r01_v4_ell_*: 25 ms with
// the chain, 6.4 ms without).  So (a) the aggregate of a tile is posted right after its inserts (it is the number
// of atomicMin that found an EMPTY slot), (b) the tile is compacted into a staging buffer and COMMITTED ONE TILE
// LATER, after the next tile's inserts — by then its offset is known without waiting — and (c) the chain is
// hierarchical (CtaChain below): shared memory inside a CTA, a flat decoupled look-back over CTA blocks in global
// memory.
#pragma once
#include "kernels.cuh"

namespace bsk {

constexpr int ELL_MAX_WARPS = 24;        // warps per CTA (one persistent CTA per SM)
constexpr int ELL_QCAP = 192;            // loser-queue entries per warp (a batch adds at most 128)


struct EllArgs {
  const int* __restrict__ Arow;   // An+1 absolute offsets
  const int* __restrict__ Acol;
  const u32* __restrict__ Bell;   // Bn+1 rows of W columns, EMPTY-padded; row Bn is all EMPTY (what an absent A nonzero selects)
  int An, Bn;
  float inv_bm;                   // slightly less than 2^32 / Bm: slot = umulhi(key, floor(T_home * inv_bm)) < T_home
  u32 TW;                         // table words per output row (multiple of 128)
  u32 SW;                         // staging words per tile = R * max_len(A) * W
  u32 lf16;                       // table sizing rule, see ell_table_limit
  u32 Bm;
  void* Crow; int is64;
  int* Ccol;
  u64* blk_status;                // [iterations * gridDim.x] block-level look-back words, zero before the launch
  DevScalars* sc;
  u32 ntiles;
  u32 nbuf;                       // sort kernel: staging buffers per warp (commit lag = nbuf - 1 tiles)
  u32 debug_nochain;              // timing experiments only, builds with -DBSPGEMM_DEBUG_KNOBS (env BSPGEMM_DEBUG_NOCHAIN): skip the scan, rows land at upper-bound offsets
  u32 one, mone;                  // 1 and 0xFFFFFFFF as run-time values (IMAD-form comparators, see cmpx in kernels.cuh)
};

// The "no chain" timing knob produces WRONG RESULTS by design; release builds compile it out (the field is ignored).
#ifdef BSPGEMM_DEBUG_KNOBS
#define DBG_NOCHAIN(p) ((p).debug_nochain)
#else
#define DBG_NOCHAIN(p) 0u
#endif

// Table geometry of a row with lenA nonzeros in A (cap = lenA*W >= its IP): `lim` slots, a multiple of 128 (the
// compaction reads 128 slots per warp instruction); keys are mapped to the first lim-32 "home" slots, the last 32 only
// take keys pushed right by collisions.  Load factor of the home slots <= 4/7.
__host__ __device__ constexpr u32 ell_table_limit(u32 lenA, u32 W, u32 lf16 = 28u) {     // lf16/16 = min home slots per key
  const u32 cap = lenA * W;
  u32 lim = (2u * cap + 127u) & ~127u;
  if (lim < 128u) lim = 128u;
  if ((lim - 32u) * 16u < cap * lf16) lim += 128u;
  return lim;
}
__host__ __device__ constexpr u32 ell_warp_words(u32 R, u32 TW, u32 SW) { return R * TW + SW + 2u * ELL_QCAP + 4u * 8u; }
constexpr u32 ELL_CTA_WORDS = 320;       // CtaChain, after the warp regions

// An entry loaded from Acol must lie in [0,Bn): Bn itself is the internal "no row" sentinel of absent slots, so a stored Bn (or
// anything beyond) becomes 0xFFFFFFFF here, which the `> Bn` test of the kernels flags as BSPGEMM_ERR_BADARG like k_estimate does.
__device__ __forceinline__ int acol_checked(int j, int Bn) { return (u32)j < (u32)Bn ? j : -1; }

// ---- B (CSR) -> ELL.  LPR = W/4 lanes write one row as uint4 each; also validates B's columns.  SORTED: every ELL row
// is sorted ascending (EMPTY padding last) by a small register network — the sorting-network kernel (fused_sort.cuh)
// starts its merges from these runs; the reference accepts unsorted rows, so nothing may be assumed.
// Every thread converts its 16-byte part of ELL_RPT rows (row, row + H, ...: H rows apart, so that a warp still reads and writes
// contiguous spans): the Brow loads of all of them, then the Bcol loads of all of them are in flight together — with one row per
// thread the two dependent load phases left half of the memory latency uncovered (4.1 TB/s; config 3: 0.134 ms per call).
constexpr int ELL_RPT = 2;
template <int W, bool SORTED>
__global__ void __launch_bounds__(256) k_build_ell(const int* __restrict__ Brow, const int* __restrict__ Bcol, int Bn, u32 Bm,
                                                   u32* __restrict__ Bell, DevScalars* sc, const u32 pad) {   // pad: EMPTY, or EMPTY_F for the floating-point network (fused_sort.cuh)
  constexpr int LPR = W / 4;
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long H = ((long long)Bn + ELL_RPT) / ELL_RPT;             // rows per slab; row Bn (the all-EMPTY "no row" row) included
  const int part = (int)(gtid % LPR);
  long long row[ELL_RPT];
  bool live[ELL_RPT];
  int bs[ELL_RPT], be[ELL_RPT];
#pragma unroll
  for (int r = 0; r < ELL_RPT; ++r) {
    row[r] = gtid / LPR + r * H;
    live[r] = gtid / LPR < H && row[r] <= Bn;    // lanes beyond the matrix only take part in the shuffles
    if (!live[r]) row[r] = Bn;
    bs[r] = row[r] < Bn ? Brow[row[r]] : 0;
    be[r] = row[r] < Bn ? Brow[row[r] + 1] : 0;
  }
  u32 x[ELL_RPT][4];
  u32 bad = 0;
#pragma unroll
  for (int r = 0; r < ELL_RPT; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int o = bs[r] + part * 4 + k;
      x[r][k] = (o < be[r]) ? (u32)__ldg(&Bcol[o]) : pad;
    }
#pragma unroll
  for (int r = 0; r < ELL_RPT; ++r) {
#pragma unroll
    for (int k = 0; k < 4; ++k) if (bs[r] + part * 4 + k < be[r] && x[r][k] >= Bm) { bad = 1; x[r][k] = pad; }   // a column outside [0,Bm)
    if (SORTED) bitonic_sort_rows<4, LPR, 1>(x[r], (u32)part);          // 256 threads = whole rows: LPR divides 32
    if (live[r]) reinterpret_cast<uint4*>(Bell)[row[r] * LPR + part] = make_uint4(x[r][0], x[r][1], x[r][2], x[r][3]);
  }
  if (bad) atomicOr(&sc->err, 4u);
}

// ------------------------------------------------------------------------------------------------ hierarchical tile chain
// Level 1 (shared memory): the compute warps of a CTA work on NW consecutive tiles per iteration ("block" b =
// iteration * gridDim.x + blockIdx.x) and post their aggregates into a 4-deep ring of slots.  Level 2 (global): a flat
// decoupled look-back over blocks — ~16 blocks per microsecond at the target rate instead of ~350 tiles, and one
// poller per CTA instead of one per warp (3256 warps polling the same few status lines starved the L2 slices that
// also had to take the posts).  The look-back is run by a HELPER WARP (the last warp of the CTA, no row work): as
// soon as the block's last aggregate is posted it publishes the block total, walks back, and leaves the block's
// offset in shared memory.  The compute warps need that offset only one whole tile later (deferred commit), so they
// normally never wait: with the walk done lazily by the first committing warp the chain cost 12 % of config 3 and
// 40 % of config 2 (profiles/r01_sort_nochain_sweep.txt).
constexpr u32 CH_RING = 8;   // slots; supports a commit lag of up to 3 tiles (ring >= 2*lag + 2, see chain_helper)
struct CtaChain {
  u32 cnt[CH_RING];           // compute warps that have posted in the slot (recycled by the helper)
  u64 base[CH_RING];          // (tag << 48) | exclusive prefix of the block, tag = (iteration & 0xfff) + 1
  u32 agg[CH_RING][32];       // (tag << 20) | aggregate of warp w's tile
  u32 blkid[CH_RING];         // dynamic block assignment (chain_helper_dyn): block id of the CTA's iteration ...
  u32 blkit[CH_RING];         // ... valid when this holds iteration + 1
};
constexpr u32 CH_AGG_MASK = (1u << 20) - 1;
constexpr u64 CH_BASE_MASK = (1ull << 48) - 1;

// Publish this warp's tile aggregate for `iter` (lane 0 does the work).
__device__ __forceinline__ void chain_post(CtaChain* cc, u32 iter, u32 warp, u32 agg) {
  if (lane_id() == 0) {
    const u32 s = iter & (CH_RING - 1u), tag = (iter & 0xfffu) + 1u;
    *reinterpret_cast<volatile u32*>(&cc->agg[s][warp]) = (tag << 20) | agg;
    __threadfence_block();
    atomicAdd(&cc->cnt[s], 1u);
  }
}

// Flat decoupled look-back over blocks: exclusive prefix of block `blk` (whole warp, blocking).  K windows of 32
// blocks are loaded together (one L2 round trip).
static __device__ __noinline__ u64 chain_walk(const u64* blk_status, u32 blk) {
  constexpr int K = 5;
  const u32 lane = lane_id();
  u64 excl = 0;
  long long idx = (long long)blk - 1;
  bool done = (blk == 0);
  while (!done) {
    u64 st[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { const long long my = idx - 32 * k - lane; st[k] = (my >= 0) ? ld_status(&blk_status[my]) : ST_INC; }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (done) break;
      const long long my = idx - 32 * k - lane;
      const u32 inc0 = __ballot_sync(0xffffffffu, (st[k] >> 62) == 2);
      const u32 first0 = inc0 ? (u32)(__ffs(inc0) - 1) : 32u;
      while (__any_sync(0xffffffffu, lane <= first0 && (st[k] >> 62) == 0)) {   // only blocks nearer than the first INC matter
        __nanosleep(100);
        if ((st[k] >> 62) == 0) st[k] = ld_status(&blk_status[my]);
      }
      const u32 inc_mask = __ballot_sync(0xffffffffu, lane <= first0 && (st[k] >> 62) == 2);
      const u32 first = inc_mask ? (u32)(__ffs(inc_mask) - 1) : 32u;
      u64 v = (lane <= first) ? (st[k] & ST_VAL) : 0;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      excl += v;
      if (inc_mask) done = true;
    }
    idx -= 32 * K;
  }
  return excl;
}

// The helper warp: for every iteration in which this CTA owns tiles, wait for the compute warps' aggregates, publish
// the block total, walk back, publish the block's inclusive prefix and leave its offset in shared memory.
// ncompute = compute warps of the CTA; tiles of iteration i: [i*stride + cta_first, ... + ncompute) clipped to ntiles.
// lag = tiles between a compute warp's post(i) and its commit(i).  Slot recycling: when every warp has posted
// iteration i, every warp has finished iteration i-1, i.e. committed iteration i-1-lag: that slot is reset.  Warps may
// by then have posted up to iteration i+lag (they need base(i-1), already published), so the ring must hold
// iterations i-lag .. i+lag plus the one being reset: CH_RING >= 2*lag + 2.
static __device__ __noinline__ void chain_helper(CtaChain* cc, u64* blk_status, u32 ntiles, u32 stride, u32 cta_first, u32 ncompute, u32 lag) {
  const u32 lane = lane_id();
  for (u32 iter = 0;; ++iter) {
    const unsigned long long first_tile = (unsigned long long)iter * stride + cta_first;
    if (first_tile >= ntiles) break;
    const u32 expected = (u32)min((unsigned long long)ncompute, (unsigned long long)ntiles - first_tile);
    const u32 s = iter & (CH_RING - 1u), tag = (iter & 0xfffu) + 1u, blk = iter * gridDim.x + blockIdx.x;
    while (*reinterpret_cast<volatile u32*>(&cc->cnt[s]) != expected) __nanosleep(100);
    __threadfence_block();
    const u32 w = (lane < expected) ? (*reinterpret_cast<volatile u32*>(&cc->agg[s][lane]) & CH_AGG_MASK) : 0u;
    const u32 total = __reduce_add_sync(0xffffffffu, w);
    if (lane == 0) {
      if (iter >= 1u + lag) *reinterpret_cast<volatile u32*>(&cc->cnt[(iter - 1u - lag) & (CH_RING - 1u)]) = 0;
      __threadfence_block();
      st_status(&blk_status[blk], ST_AGG | (u64)total);
    }
    const u64 excl = chain_walk(blk_status, blk);
    if (lane == 0) {
      st_status(&blk_status[blk], ST_INC | (excl + (u64)total));
      *reinterpret_cast<volatile u64*>(&cc->base[s]) = ((u64)tag << 48) | excl;
      __threadfence_block();
    }
    __syncwarp();
  }
}

// Dynamic variant: the CTA's block of iteration i is not i*gridDim.x + blockIdx.x but the next id of a global counter,
// claimed five iterations ahead (the compute warps look up to three tiles ahead).  An SM that runs slower than the others
// (the static deal made every CTA process the same number of blocks: the kernel ran at the pace of the slowest SM and
// the warps of the others spent 9 % of their time waiting for offsets, profiles/r01_sort_cfg3_static_blocks.txt) then
// simply takes fewer blocks.  Ids are consecutive in time, so neighbours in the chain still run at the same time.
static __device__ __noinline__ void chain_helper_dyn(CtaChain* cc, u64* blk_status, u32* counter, u32 ntiles, u32 ncompute, u32 lag) {
  const u32 lane = lane_id();
  const u32 nblocks = (ntiles + ncompute - 1u) / ncompute;
  // The first five iterations take the static ids it * gridDim.x + blockIdx.x, the global counter hands out the rest.  (All
  // five claimed from the counter at kernel start gave every CTA five CONSECUTIVE ids: block 42, the first of one CTA, then
  // waited for block 41, the fifth of another — every first commit of the kernel waited four iterations, ~30 us at config 3,
  // and a shard of six iterations per CTA ran 2.5 times longer than its work: tools/shard_sweep.py, profiles/r02_sweeps.txt.)
  auto claim = [&](u32 it) {
    if (lane == 0) {
      const u32 id = it < 5u ? it * gridDim.x + blockIdx.x : 5u * gridDim.x + atomicAdd(counter, 1u);
      *reinterpret_cast<volatile u32*>(&cc->blkid[it & (CH_RING - 1u)]) = id;
      __threadfence_block();
      *reinterpret_cast<volatile u32*>(&cc->blkit[it & (CH_RING - 1u)]) = it + 1u;
    }
  };
  for (u32 it = 0; it < 5u; ++it) claim(it);
  __syncwarp();
  for (u32 iter = 0;; ++iter) {
    const u32 s = iter & (CH_RING - 1u), tag = (iter & 0xfffu) + 1u;
    const u32 blk = *reinterpret_cast<volatile u32*>(&cc->blkid[s]);
    if (blk >= nblocks) break;
    const u32 expected = min(ncompute, ntiles - blk * ncompute);
    while (*reinterpret_cast<volatile u32*>(&cc->cnt[s]) != expected) __nanosleep(400);   // (shorter sleeps here and in chain_resolve: 3.14 ms against 3.11 at config 3)
    __threadfence_block();
    const u32 w = (lane < expected) ? (*reinterpret_cast<volatile u32*>(&cc->agg[s][lane]) & CH_AGG_MASK) : 0u;
    const u32 total = __reduce_add_sync(0xffffffffu, w);
    if (lane == 0) {
      if (iter >= 1u + lag) *reinterpret_cast<volatile u32*>(&cc->cnt[(iter - 1u - lag) & (CH_RING - 1u)]) = 0;
      __threadfence_block();
      st_status(&blk_status[blk], ST_AGG | (u64)total);
    }
    claim(iter + 5u);                        // every warp has posted `iter`: nobody reads the ids of iterations <= iter any more (ring of 8)
    const u64 excl = chain_walk(blk_status, blk);
    if (lane == 0) {
      st_status(&blk_status[blk], ST_INC | (excl + (u64)total));
      *reinterpret_cast<volatile u64*>(&cc->base[s]) = ((u64)tag << 48) | excl;
      __threadfence_block();
    }
    __syncwarp();
  }
}
// Block id of the CTA's iteration `it` (compute warps; normally already there).
__device__ __forceinline__ u32 chain_block_of(CtaChain* cc, u32 it) {
  const u32 s = it & (CH_RING - 1u);
  while (*reinterpret_cast<volatile u32*>(&cc->blkit[s]) != it + 1u) __nanosleep(100);
  __threadfence_block();
  return *reinterpret_cast<volatile u32*>(&cc->blkid[s]);
}

// Exclusive prefix of this warp's tile of iteration `iter` (whole warp).  Called `lag` tiles after chain_post(iter).
__device__ __forceinline__ u64 chain_resolve(CtaChain* cc, u32 iter, u32 warp) {
  const u32 lane = lane_id();
  const u32 s = iter & (CH_RING - 1u), tag = (iter & 0xfffu) + 1u;
  volatile u64* basep = &cc->base[s];
  u64 bw;
  while ((u32)((bw = *basep) >> 48) != tag) __nanosleep(200);
  const u32 w = (lane < warp) ? (*reinterpret_cast<volatile u32*>(&cc->agg[s][lane]) & CH_AGG_MASK) : 0u;
  return (bw & CH_BASE_MASK) + __reduce_add_sync(0xffffffffu, w);
}

// ------------------------------------------------------------------------------------------------ slow paths (out of line)
// Shared memory is addressed with 32-bit shared-window addresses in the hot path: through generic pointers the compiler
// re-derives the window base (S2R SR_CgaCtaId, LEA) next to every predicated store.
__device__ __forceinline__ u32 atoms_min(u32 saddr, u32 x) {
  u32 old; asm volatile("atom.shared.min.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(x)); return old;
}
__device__ __forceinline__ uint4 lds128(u32 saddr) {
  uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr)); return v;
}
__device__ __forceinline__ uint2 lds64(u32 saddr) {
  uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr)); return v;
}
__device__ __forceinline__ u32 lds32(u32 saddr) {
  u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr)); return v;
}
__device__ __forceinline__ void sts32(u32 saddr, u32 v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(saddr), "r"(v)); }
__device__ __forceinline__ void sts64(u32 saddr, u32 a, u32 b) { asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(saddr), "r"(a), "r"(b)); }
__device__ __forceinline__ void sts128(u32 saddr, u32 a, u32 b, u32 c, u32 d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d));
}

// Queue entry: (key, y) with y = next slot (shared byte address, 18 bits) | (lim/128) << 18 | row << 24.
// Re-insert the queued losers [lo,hi) in rounds of 32 (one entry per lane, every lane walks its collision chain).
// Returns (rows whose table spilled past its limit) << 16 | (#keys that found an EMPTY slot).
static __device__ __noinline__ u32 ell_drain(u32 tab_s, u32 queue_s, u32 lo, u32 hi, u32 TW) {
  u32 ovf = 0, added = 0;
  __syncwarp();
  for (u32 i = lo + lane_id(); i < hi; i += 32) {
    const uint2 ent = lds64(queue_s + 8u * i);
    u32 x = ent.x, addr = ent.y & 0x3ffffu;
    const u32 r = ent.y >> 24;
    const u32 lim = tab_s + 4u * (r * TW + (((ent.y >> 18) & 63u) << 7));
    while (true) {
      if (addr >= lim) { ovf |= 0x10000u << r; break; }
      const u32 old = atoms_min(addr, x);
      if (old == EMPTY) { ++added; break; }
      if (old == x) break;
      x = max(old, x); addr += 4u;
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
  const u32 T = TW - cap;                                         // >= cap by construction of TW
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

static __device__ __noinline__ u32 ell_count_table(const u32* tabr, u32 lim) {     // occupied slots (after a rebuild)
  u32 c = 0;
  for (u32 q = lane_id(); q < lim; q += 32) c += (tabr[q] != EMPTY) ? 1u : 0u;
  return __reduce_add_sync(0xffffffffu, c);
}

// ------------------------------------------------------------------------------------------------ the fused kernel
// Every warp is an independent worker on tiles of R consecutive rows.  Per iteration (tile t):
//   1. init the R tables; insert the B rows of tile t (loaded one iteration ago); drain the losers;
//   2. post the tile's aggregate (CtaChain);
//   3. COMMIT tile t-1 (its offset is known by now): row pointers + staging buffer -> Ccol, coalesced;
//   4. start the Acol loads of tile t+1; compact the tables of tile t into the staging buffer;
//   5. start the B-row loads of tile t+1.
template <int W, int R>
__global__ void __launch_bounds__(ELL_MAX_WARPS * 32, 1) k_fused_ell(const EllArgs p) {
  constexpr int LPR = W / 4;               // lanes per B row
  constexpr int NSEG = 32 / LPR;           // B rows per LDG.128 warp instruction ("batch")
  constexpr int NBS = 64 / NSEG;           // batches per chunk of 64 A nonzeros
  constexpr int NBG = NBS < 8 ? NBS : 8;   // batches in flight (a "group")
  extern __shared__ __align__(16) u32 smem[];
  const u32 warp = threadIdx.x >> 5, lane = lane_id(), nwarps = (blockDim.x >> 5) - 1u;   // compute warps; the last warp is the chain helper
  const u32 TW = p.TW;
  const u32 wwords = ell_warp_words(R, TW, p.SW);
  u32* tab = smem + (size_t)warp * wwords;
  // warp region: R tables | staging (compacted rows of the tile awaiting its commit, back to back) | loser queue |
  // per-row parameters (slot scale, table byte address, queue tag, -)
  const u32 tab_s = (u32)__cvta_generic_to_shared(tab);
  const u32 stage_s = tab_s + R * TW * 4u, queue_s = stage_s + p.SW * 4u, par_s = queue_s + ELL_QCAP * 8u;
  CtaChain* cc = reinterpret_cast<CtaChain*>(smem + (size_t)nwarps * wwords);
  for (u32 i = threadIdx.x; i < sizeof(CtaChain) / 4; i += blockDim.x) reinterpret_cast<u32*>(cc)[i] = 0;
  __syncthreads();                          // the only CTA-wide barrier
  if (warp == nwarps) {
    if (!DBG_NOCHAIN(p)) chain_helper(cc, p.blk_status, p.ntiles, gridDim.x * nwarps, blockIdx.x * nwarps, nwarps, 1u);
    return;
  }
  const u32 sub = lane / LPR, part = lane % LPR;
  const u32 ltmask = (1u << lane) - 1u;
  const uint4* __restrict__ Bell4 = reinterpret_cast<const uint4*>(p.Bell);
  u32 ipc = 0;                              // intermediate products seen by this lane

  // Tiles are dealt round-robin: in iteration i, warp w of CTA c works on tile i*stride + c*nwarps + w.  Consecutive
  // tiles are processed at the same time by neighbouring warps (ids from an atomic counter, taken a tile ahead,
  // delayed every commit by a tile time).  Every warp of the grid is resident (one CTA per SM), so the chain
  // cannot wait on a tile that never runs.
  const u32 stride = gridDim.x * nwarps;
  const u32 cta_first = blockIdx.x * nwarps;

  auto load_rowptr = [&](u32 t) -> int {     // lane r (r <= R) gets Arow[t*R + r], clamped to the matrix
    if (t >= p.ntiles) return 0;
    const long long r0 = (long long)t * R;
    const int nr = (int)min((long long)R, (long long)p.An - r0);
    return p.Arow[r0 + min((int)lane, nr)];
  };
  auto load_acol = [&](int abase, int e0, int E, int& j0, int& j1) {      // absent entries select the all-EMPTY row Bn
    j0 = p.Bn; j1 = p.Bn;
    if (e0 + (int)lane < E) j0 = acol_checked(p.Acol[abase + e0 + (int)lane], p.Bn);
    if (e0 + 32 + (int)lane < E) j1 = acol_checked(p.Acol[abase + e0 + 32 + (int)lane], p.Bn);
  };
  auto check_acol = [&](int& j0, int& j1) {
    if (((u32)j0 > (u32)p.Bn) | ((u32)j1 > (u32)p.Bn)) {
      atomicOr(&p.sc->err, 1u);
      if ((u32)j0 > (u32)p.Bn) j0 = p.Bn;
      if ((u32)j1 > (u32)p.Bn) j1 = p.Bn;
    }
  };
  auto load_group = [&](int g, int j0, int j1, uint4 (&v)[NBG]) {      // the B rows of batches g .. g+NBG-1 of a chunk
#pragma unroll
    for (int u = 0; u < NBG; ++u) {
      const int seg = (g + u) * NSEG + (int)sub;
      const int j = __shfl_sync(0xffffffffu, ((g + u) * NSEG < 32) ? j0 : j1, seg & 31);
      v[u] = __ldg(&Bell4[(size_t)j * LPR + part]);
    }
  };
  // commit of a finished tile: its rows are in stage[0..total), lane r holds the inclusive count of row r
  auto commit = [&](u32 t, u32 iter, u32 incl_mine, u32 total) {
    const u64 excl = DBG_NOCHAIN(p) ? (u64)t * (u64)DBG_NOCHAIN(p) : chain_resolve(cc, iter, warp);
    const long long row0 = (long long)t * R;
    const int nrows = (int)min((long long)R, (long long)p.An - row0);
    if ((int)lane < nrows) st_rowptr(p.Crow, p.is64, (size_t)(row0 + lane) + 1, excl + incl_mine, &p.sc->err);
    if (t == 0 && lane == 0) st_rowptr(p.Crow, p.is64, 0, 0, &p.sc->err);
    if (t == p.ntiles - 1 && lane == 0) p.sc->total_nnz = excl + total;
    int* dst = p.Ccol + excl;
    for (u32 q = lane; q < total; q += 32) dst[q] = (int)lds32(stage_s + 4u * q);
    __syncwarp();
  };

  // ---- pipeline prologue: row pointers two tiles ahead, Acol one tile ahead, B rows from the end of the previous
  // tile's inserts
  u32 tile = cta_first + warp, iter = 0;
  int a[R + 1], an[R + 1];
  {
    const int ar = load_rowptr(tile);
    const int arn = load_rowptr(tile + stride < tile ? 0xffffffffu : tile + stride);
#pragma unroll
    for (int r = 0; r <= R; ++r) { a[r] = __shfl_sync(0xffffffffu, ar, r); an[r] = __shfl_sync(0xffffffffu, arn, r); }
  }
  int j0, j1;
  uint4 v[NBG];
  load_acol(a[0], 0, a[R] - a[0], j0, j1);
  check_acol(j0, j1);
  load_group(0, j0, j1, v);
  u32 prev_tile = 0xffffffffu, prev_incl = 0, prev_total = 0;      // the tile awaiting its commit

  while (tile < p.ntiles) {
    const u32 next = (tile + stride < tile) ? 0xffffffffu : tile + stride;
    const u32 next2 = (next + stride < next) ? 0xffffffffu : next + stride;
    const int ar_nn = load_rowptr(next2);                          // consumed at the end of the iteration
    int j0n, j1n;
    load_acol(an[0], 0, an[R] - an[0], j0n, j1n);                  // consumed after the inserts
    const int E = a[R] - a[0];
    int b[R];                                                      // first A nonzero of every row, relative to the tile
#pragma unroll
    for (int r = 0; r < R; ++r) b[r] = a[r] - a[0];
    u32 lim[R];
#pragma unroll
    for (int r = 0; r < R; ++r) lim[r] = ell_table_limit((u32)(a[r + 1] - a[r]), W, p.lf16);
    if (lane < R) {
      u32 l = 128;
#pragma unroll
      for (int r = 0; r < R; ++r) if ((int)lane == r) l = lim[r];
      const u32 scale = __float2uint_rd(__fmul_rd((float)(l - 32u), p.inv_bm));
      sts128(par_s + 16u * lane, scale, tab_s + lane * TW * 4u, (lane << 24) | ((l >> 7) << 18) | 4u, 0u);
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (a[r + 1] > a[r])
        for (u32 q = lane * 4; q < lim[r]; q += 128) sts128(tab_s + (r * TW + q) * 4u, EMPTY, EMPTY, EMPTY, EMPTY);
    __syncwarp();

    // ---- 1. insert: chunk 0 / group 0 is already in v[] (loaded one tile ago)
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
          const uint4 pr = lds128(par_s + 16u * r);
          const u32 x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          u32 sa[4], old[4];
          // padding (EMPTY) takes a harmless atomicMin(.., EMPTY) on a private bank: no branch around the atomics
#pragma unroll
          for (int k = 0; k < 4; ++k) sa[k] = pr.y + 4u * ((x[k] != EMPTY) ? __umulhi(x[k], pr.x) : lane);
#pragma unroll
          for (int k = 0; k < 4; ++k) old[k] = atoms_min(sa[k], x[k]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool valid = x[k] != EMPTY;
            const bool fresh = valid && (old[k] == EMPTY);
            const bool lose = valid && (old[k] != EMPTY) && (old[k] != x[k]);
            ipc += valid ? 1u : 0u;
            added += fresh ? 1u : 0u;
            const u32 m = __ballot_sync(0xffffffffu, lose);
            if (lose) sts64(queue_s + 8u * (qn + __popc(m & ltmask)), max(old[k], x[k]), sa[k] + pr.z);
            qn += __popc(m);
          }
          if (qn > ELL_QCAP - 128) {                              // full rounds only; the remainder waits for company
            const u32 d = ell_drain(tab_s, queue_s, qn & 31u, qn, TW); ovf |= d >> 16; added += d & 0xffffu; qn &= 31u;
          }
        }
      }
    }
    // ---- next tile: its Acol has arrived; v[] is dead until the next iteration: start the B-row loads now
    check_acol(j0n, j1n);
    load_group(0, j0n, j1n, v);
    if (qn) { const u32 d = ell_drain(tab_s, queue_s, 0u, qn, TW); ovf |= d >> 16; added += d & 0xffffu; }
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
    // ---- 2. publish the aggregate
    if (!DBG_NOCHAIN(p)) chain_post(cc, iter, warp, agg);
    // ---- 3. commit the previous tile (frees the staging buffer)
    if (prev_tile != 0xffffffffu) commit(prev_tile, iter - 1u, prev_incl, prev_total);

    // ---- 4. compact this tile into the staging buffer
    u32 run = 0, incl_mine = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (a[r + 1] > a[r]) {
        const u32 tb = tab_s + (r * TW + lane * 4u) * 4u;          // lane l owns slots q+4l .. q+4l+3 of every 128-slot chunk
        for (u32 q = 0; q < lim[r]; q += 128) {
          const uint4 t4 = lds128(tb + q * 4u);
          const bool p0 = t4.x != EMPTY, p1 = t4.y != EMPTY, p2 = t4.z != EMPTY, p3 = t4.w != EMPTY;
          const u32 m0 = __ballot_sync(0xffffffffu, p0), m1 = __ballot_sync(0xffffffffu, p1);
          const u32 m2 = __ballot_sync(0xffffffffu, p2), m3 = __ballot_sync(0xffffffffu, p3);
          u32 o = stage_s + 4u * (run + __popc(m0 & ltmask) + __popc(m1 & ltmask) + __popc(m2 & ltmask) + __popc(m3 & ltmask));
          if (p0) { sts32(o, t4.x); o += 4u; }
          if (p1) { sts32(o, t4.y); o += 4u; }
          if (p2) { sts32(o, t4.z); o += 4u; }
          if (p3) sts32(o, t4.w);
          run += __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
        }
      }
      if ((int)lane == r) incl_mine = run;
    }
    __syncwarp();

    prev_tile = tile; prev_incl = incl_mine; prev_total = run;
    tile = next; ++iter;
#pragma unroll
    for (int r = 0; r <= R; ++r) { a[r] = an[r]; an[r] = __shfl_sync(0xffffffffu, ar_nn, r); }
    j0 = j0n; j1 = j1n;
  }
  if (prev_tile != 0xffffffffu) commit(prev_tile, iter - 1u, prev_incl, prev_total);
  u64 ips = ipc;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ips += __shfl_xor_sync(0xffffffffu, ips, d);
  if (lane == 0 && ips) atomicAdd(&p.sc->total_ip, ips);
}

}  // namespace bsk
