// binary-spgemm_b200/csrc/rows_bm.cuh — big rows (more than 2048 intermediate products: the M2 and L lists) of skewed matrices:
// one 1024-thread CTA per row, a bitmap over a 1.4 M-column WINDOW in shared memory with a one-byte-per-128-columns SUMMARY,
// a LOAD-BALANCED walk over the row's products and a LOAD-BALANCED emission.
//
// Replaces, for these rows, the flag array `xb[Bm]` + per-row quickSort of SpGEMM_bigslice (final/SpGEMM_mpi_omp.c:21, 33-47):
// a set bit is "column seen", reading the bitmap left to right is the sorted distinct row.  What the kernels of round 1 paid
// for on R-MAT scale 22 (profiles/r02_rmat20_*_ncu_summary.txt):
//   * rows_sort.cuh (2049..16384 products, CTA-wide bitonic sort): O(IP log^2 IP) compare-exchanges, two barriers per far
//     stage — 173 ms for 1.01 M rows;
//   * rows_window.cuh (above 16384 products): B rows dealt to 16-lane groups, so a group that drew a 1000-entry hub row
//     walked it alone while the CTA waited at the barrier (barrier stall 15 of 26 warp-cycles per issue), one shared-memory
//     atomic per product (ATOMS: 2 cycles per LANE), every window scanned and cleared all 49152 words it might have
//     touched — 154 ms for 132 K rows.
// Here:
//   * WALK.  The B-row lengths of up to 1024 A entries are scanned into shared memory once per row; the row's products
//     0..T-1 are dealt to the warps in equal contiguous shares, lane l of a warp takes products l, l+32, ... of the share
//     and finds "which B row" by a binary search at the start and a step forward afterwards: every lane does the same amount
//     of work whatever the B-row lengths, and consecutive lanes read consecutive Bcol entries.
//   * INSERT WITHOUT ATOMICS.  Eight products per thread and step: pass A sets the bits with a plain load / OR / store (and a
//     plain byte store into the summary), which can lose updates when two threads hit the same word at the same time;
//     barrier; pass B re-reads the word of every product still held in registers and repairs a missing bit with an atomic
//     (only atomics write in pass B, so nothing is lost again); barrier.  Power-law rows are sparse in most of the window:
//     almost no repairs, and a plain LDS + STS pair costs 2 issue slots per 32 products instead of 64 cycles.
//   * EMISSION.  The summary bytes are compacted (one block scan) into the ordered list of non-empty 16-byte pieces of the
//     bitmap; every thread sums the popcounts of an equal share of the list, a second block scan gives the shares' offsets,
//     and then thread t writes output positions [t Q, (t+1) Q) of the window, Q = columns / 1024: a binary search over the
//     shares' offsets, a few pieces forward, skip the bits that belong to the thread before, emit Q columns across piece
//     boundaries.  Every thread writes the same number of columns whatever the distribution of the columns.  (Power-law
//     rows are dense on the hub columns and sparse elsewhere at every scale: with a fixed slice of the window per thread,
//     and still with one piece of the list per thread, the threads that drew the hub columns emitted tens of columns while
//     the others waited — 28 % of the kernel's instructions at 9 active lanes of 32.  A word-by-word variant, the warp
//     emitting one non-empty word per coalesced store, serialises what the lanes otherwise do in parallel: slower still.)
//   * one barrier per block scan (every warp scans the 32 warp totals itself, double-buffered), not three.
//   * the window slides: the next one starts at the smallest column seen beyond the current one, so empty stretches of the
//     column range cost nothing.  B rows need not be sorted.
//   * MEDIUM ROWS (at most 16384 products and 1024 A entries: their own list and kernel instantiation, SMALLK) keep their products
//     in registers (at most 16 per thread) and, where the matrix is wider than one window, take a COMPRESSED SINGLE PASS instead
//     of one pass per window: the 128-column pieces the row touches are marked in a piece bitmap over all of [0,Bm) (4 KB at
//     Bm = 2^22), ranked by one block scan (piece -> slot, slots in column order), the products' bits are set in the slots
//     (16 bytes each; the bitmap area holds 11264 of them; more pieces than that: back to windows) and the slots are emitted
//     like the pieces of a window.  A medium row pays the per-pass fixed costs once instead of three times at Bm = 2^22.
//   * the next row's list entry and row pointers are fetched while the current row is processed.
// MODE_COUNT: cnt[row]; MODE_FILL: columns at Ccol[Crow[row]..); MODE_STAGE: both, columns at Ccol[tofs[row]..) (staging arena).
#pragma once
#include "kernels.cuh"

namespace bsk {

constexpr u32 BM_THREADS = 1024;
constexpr u32 BM_WORDS = 44u * 1024u;                 // bitmap words per window: 176 KB = 1,441,792 columns
constexpr u32 BM_PIECES = BM_WORDS / 4u;              // 16-byte pieces (128 columns) = summary bytes = longest piece list
constexpr u32 BM_FLAGW = BM_PIECES / 4u;              // summary bytes, counted in 32-bit words
constexpr u32 BM_LISTW = BM_PIECES / 2u;              // piece list (16-bit ids), counted in 32-bit words
constexpr u32 BM_CHUNK = BM_THREADS;                  // A entries per chunk: one per thread
constexpr u32 BM_STEP = 8;                            // products per thread between two barriers
constexpr u32 BM_REGS = 16;                           // rows of up to 16 * 1024 products: the products stay in registers over all windows
constexpr u32 BM_MAX_WINDOWS = 8;                     // host: wider matrices keep the sort / global-bitmap kernels
constexpr size_t BM_SMEM = (size_t)(BM_WORDS + BM_FLAGW + BM_LISTW + 3u * (BM_CHUNK + 4u)) * 4u;
static_assert(BM_PIECES <= 65536u, "piece ids are 16 bits");
static_assert(3u * BM_THREADS >= BM_FLAGW, "three summary words per thread cover the summary");
static_assert(2u * BM_THREADS <= BM_FLAGW && 2u * BM_THREADS * 2u <= (BM_THREADS + 4u) * 4u, "compressed pass: 2048 piece-bitmap words in the summary area, 2048 u16 ranks in cb");
constexpr u32 BM_COMP_MAX_BM = 2u * BM_THREADS * 32u * 128u;   // 8,388,608 columns: piece ids are 16 bits

// Block-wide exclusive scan with ONE barrier: lane 31 of every warp posts the warp's total, after the barrier every warp scans
// the 32 totals itself.  red: 2 x 32 words, used alternately (a thread can be at most one scan ahead of the slowest one).
__device__ __forceinline__ u32 bm_scan(u32 v, u32* red, u32& flip, u32* total) {
  const u32 lane = lane_id(), w = threadIdx.x >> 5;
  const u32 inc = warp_incl_scan(v);
  u32* r = red + flip;
  flip ^= 32u;
  if (lane == 31) r[w] = inc;
  __syncthreads();
  const u32 x = r[lane];
  const u32 xi = warp_incl_scan(x);
  *total = __shfl_sync(0xffffffffu, xi, 31);
  return __shfl_sync(0xffffffffu, xi - x, (int)w) + inc - v;
}

__device__ __forceinline__ u32 bm_popc4(const uint4& x) { return __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w); }
__device__ __forceinline__ u32 bm_word(const uint4& x, u32 i) { return i == 0 ? x.x : i == 1 ? x.y : i == 2 ? x.z : x.w; }

// SMALLK: the list holds rows of at most 16384 products and 1024 A entries (k_build_lists): products in registers, compressed single
// pass when cmode; the walking code is compiled out.  !SMALLK: any row, walked per window; the register path is compiled out.
template <int MODE, bool SMALLK>
__global__ void __launch_bounds__(BM_THREADS, 1) k_rows_bm(Csr m, const u32* __restrict__ list_a, const u32* __restrict__ nlist_a,
                                                           u32* __restrict__ ctr, u32* __restrict__ cnt,
                                                           const void* __restrict__ Crow, int is64, int* __restrict__ Ccol,
                                                           const u64* __restrict__ tofs, DevScalars* sc, u32 cmode) {
  extern __shared__ __align__(16) u32 bm[];
  u32* const flw = bm + BM_WORDS;                     // summary bytes, read as words by the emission
  unsigned char* const fl = reinterpret_cast<unsigned char*>(flw);
  unsigned short* const plist = reinterpret_cast<unsigned short*>(flw + BM_FLAGW);
  u32* const off = flw + BM_FLAGW + BM_LISTW;         // BM_CHUNK + 1: exclusive scan of the chunk's B-row lengths
  u32* const bst = off + BM_CHUNK + 4u;               // BM_CHUNK: where every B row starts in Bcol, minus its off[]
  u32* const cb = bst + BM_CHUNK + 4u;                // BM_THREADS + 1: output offset of every thread's share of the piece list
  uint4* const bm4 = reinterpret_cast<uint4*>(bm);
  unsigned short* const pre16 = reinterpret_cast<unsigned short*>(cb);   // compressed pass: slots before every word of the piece bitmap (dead before cb is written)
  __shared__ u32 s_red[64];
  __shared__ u32 s_idx[2], s_above;
  const u32 tid = threadIdx.x, lane = lane_id(), wid = tid >> 5;
  const u32 FULL = 0xffffffffu;
  constexpr u32 wbits = BM_WORDS << 5;
  u32 flip = 0;
  for (u32 q = tid; q < (BM_WORDS + BM_FLAGW) / 4u; q += BM_THREADS) bm4[q] = make_uint4(0u, 0u, 0u, 0u);
  const u32 n = *nlist_a;
  if (tid == 0) { s_idx[0] = atomicAdd(ctr, 1u); s_above = EMPTY; }
  __syncthreads();
  u32 idx = s_idx[0];
  int row = 0, a0 = 0, a1 = 0;
  if (idx < n) { row = (int)list_a[idx]; a0 = m.Arow[row]; a1 = m.Arow[row + 1]; }
  u32 it = 0;
  while (idx < n) {
    u32 nidx_reg = 0;
    if (tid == 0) nidx_reg = atomicAdd(ctr, 1u);      // the next row's list index: consumed after the first walk
    u32 nidx = n; int nrow = 0, na0 = 0, na1 = 0;
    const u32 nA = (u32)(a1 - a0);
    const u64 base = (MODE == MODE_FILL) ? ld_rowptr(Crow, is64, (size_t)row) : (MODE == MODE_STAGE) ? tofs[row] : 0;
    u64 done = 0;
    u32 start = 0;
    bool first = true, bad = false, small = false, comp = false;
    u32 rv[BM_REGS], have = 0, nslots = 0;
    while (true) {                                    // windows
      u32 above = EMPTY;
      const u32 end = start + wbits;                  // Bm <= 2^31 and wbits < 2^21: no wrap
      if (first && nA <= BM_CHUNK) {                  // one chunk: build the product index here, then decide how to walk
        u32 len = 0, bs = 0;
        if (tid < nA) {
          const int j = m.Acol[(size_t)a0 + tid];
          if ((u32)j < (u32)m.Bn) { bs = (u32)m.Brow[j]; len = (u32)m.Brow[j + 1] - bs; }
        }
        u32 tot;
        const u32 o = bm_scan(len, s_red, flip, &tot);
        off[tid] = o; bst[tid] = bs - o;
        if (tid == BM_THREADS - 1) off[BM_CHUNK] = tot;
        __syncthreads();
        small = SMALLK && tot <= BM_REGS * BM_THREADS;
        if (SMALLK && !small) bad = true;             // (cannot happen: the list was built from the same product counts)
        if (SMALLK && small) {                        // every thread loads its (at most 16) products once
          const u32 per = ((tot + BM_THREADS - 1u) / BM_THREADS) * 32u;
          const u32 ws = wid * per, we = min(tot, ws + per);
          u32 p = ws + lane;
          have = 0;
          nslots = per >> 5;                          // the same for every thread: whole register slots are skipped below
          if (p < we) {
            u32 lo = 0, hi = nA;
            while (hi - lo > 1u) { const u32 mid = (lo + hi) >> 1; if (off[mid] <= p) lo = mid; else hi = mid; }
            u32 e = lo, nx = off[e + 1], badd = bst[e];
#pragma unroll
            for (int k = 0; k < (int)BM_REGS; ++k) {
              if ((u32)k >= nslots) break;
              if (p < we) {
                if (p >= nx) { do { ++e; nx = off[e + 1]; } while (p >= nx); badd = bst[e]; }
                rv[k] = (u32)__ldg(&m.Bcol[(u32)(badd + p)]);
                have |= 1u << k;
                p += 32u;
              }
            }
          }
        }
      }
      u32 npieces = 0;
      if (SMALLK && small && cmode && first) {
        // ---- COMPRESSED SINGLE PASS (rows of up to 16384 products, matrices wider than one window): instead of one bitmap pass
        // per 1.44 M-column window, (1) mark the 128-column PIECES the row touches in a piece bitmap over all of [0,Bm) (Bm/128
        // bits: 4 KB at Bm = 2^22), (2) rank them — piece -> slot, slots in column order, (3) set the products' bits in the
        // slots (16 bytes each, the bitmap area holds 11264 of them), (4) emit the slots.  Marks and bits are plain
        // read-modify-writes repaired by a second pass, like the windowed insert.  More than 11264 pieces: back to windows.
        const u32 uBm = (u32)m.Bm;
#pragma unroll
        for (int k = 0; k < (int)BM_REGS; ++k) {
          if ((u32)k >= nslots) break;
          if ((have >> k) & 1u) {
            const u32 v = rv[k];
            if (v >= uBm) bad = true;
            else { const u32 pc = v >> 7, b = 1u << (pc & 31u), o = flw[pc >> 5]; if (!(o & b)) flw[pc >> 5] = o | b; }
          }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < (int)BM_REGS; ++k) {
          if ((u32)k >= nslots) break;
          if ((have >> k) & 1u) {
            const u32 v = rv[k];
            if (v < uBm) { const u32 pc = v >> 7, b = 1u << (pc & 31u); if (!(flw[pc >> 5] & b)) atomicOr(&flw[pc >> 5], b); }
          }
        }
        if (tid == 0) s_idx[(it + 1u) & 1u] = nidx_reg;
        __syncthreads();
        const u32 w0 = flw[2u * tid], w1 = flw[2u * tid + 1u];            // thread t ranks the pieces of bitmap words 2 t, 2 t + 1
        u32 pos = bm_scan(__popc(w0) + __popc(w1), s_red, flip, &npieces);
        comp = npieces <= BM_PIECES;
        if (comp) {
          pre16[2u * tid] = (unsigned short)pos;
          pre16[2u * tid + 1u] = (unsigned short)(pos + __popc(w0));
          for (u32 b = w0; b; b &= b - 1u) plist[pos++] = (unsigned short)(tid * 64u + (u32)__ffs((int)b) - 1u);
          for (u32 b = w1; b; b &= b - 1u) plist[pos++] = (unsigned short)(tid * 64u + 32u + (u32)__ffs((int)b) - 1u);
        } else { flw[2u * tid] = 0u; flw[2u * tid + 1u] = 0u; npieces = 0; }
        __syncthreads();
        if (comp) {
#pragma unroll
          for (int k = 0; k < (int)BM_REGS; ++k) {
            if ((u32)k >= nslots) break;
            if ((have >> k) & 1u) {
              const u32 v = rv[k];
              if (v < uBm) {
                const u32 pc = v >> 7, pw = flw[pc >> 5];
                const u32 a = ((u32)pre16[pc >> 5] + __popc(pw & ((1u << (pc & 31u)) - 1u))) * 4u + ((v >> 5) & 3u);
                const u32 bit = 1u << (v & 31u), o = bm[a];
                if (!(o & bit)) bm[a] = o | bit;
              }
            }
          }
          __syncthreads();
#pragma unroll
          for (int k = 0; k < (int)BM_REGS; ++k) {
            if ((u32)k >= nslots) break;
            if ((have >> k) & 1u) {
              const u32 v = rv[k];
              if (v < uBm) {
                const u32 pc = v >> 7, pw = flw[pc >> 5];
                const u32 a = ((u32)pre16[pc >> 5] + __popc(pw & ((1u << (pc & 31u)) - 1u))) * 4u + ((v >> 5) & 3u);
                const u32 bit = 1u << (v & 31u);
                if (!(bm[a] & bit)) atomicOr(&bm[a], bit);
              }
            }
          }
          __syncthreads();
        }
      }
      if (comp) {
        // (the slots are filled; nothing to walk)
      } else if (SMALLK) {
        // pass A / barrier / pass B straight from the registers: no walk, no loads in the second and later windows
#pragma unroll
        for (int k = 0; k < (int)BM_REGS; ++k) {
          if ((u32)k >= nslots) break;
          if ((have >> k) & 1u) {
            const u32 d = rv[k] - start;
            if (d < wbits) {
              const u32 w = d >> 5, bit = 1u << (d & 31u), o = bm[w];
              if (!(o & bit)) bm[w] = o | bit;
              fl[w >> 2] = 1;
            } else if (rv[k] >= end) above = min(above, rv[k]);
          }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < (int)BM_REGS; ++k) {
          if ((u32)k >= nslots) break;
          if ((have >> k) & 1u) {
            const u32 d = rv[k] - start;
            if (d < wbits) { const u32 w = d >> 5, bit = 1u << (d & 31u); if (!(bm[w] & bit)) atomicOr(&bm[w], bit); }
          }
        }
        __syncthreads();
      } else {
        for (u32 c0 = 0; c0 < nA; c0 += BM_CHUNK) {
          const u32 nc = min(BM_CHUNK, nA - c0);
          if (nA > BM_CHUNK) {                          // several chunks: (re)build the chunk's product index, one A entry per thread
            u32 len = 0, bs = 0;                        // (the scan's barrier comes after everybody's last use of off[] / bst[])
            if (tid < nc) {
              const int j = m.Acol[(size_t)a0 + c0 + tid];
              if ((u32)j < (u32)m.Bn) { bs = (u32)m.Brow[j]; len = (u32)m.Brow[j + 1] - bs; }
            }
            u32 tot;
            const u32 o = bm_scan(len, s_red, flip, &tot);
            off[tid] = o; bst[tid] = bs - o;            // address of product p of this B row: Bcol + bst + p (mod 2^32 arithmetic)
            if (tid == BM_THREADS - 1) off[BM_CHUNK] = tot;
            __syncthreads();
          }
          const u32 T = off[BM_CHUNK];
          const u32 per = ((T + BM_THREADS - 1u) / BM_THREADS) * 32u;       // products per warp (a multiple of 32)
          const u32 steps = (per / 32u + BM_STEP - 1u) / BM_STEP;           // the same for every thread: the loop holds barriers
          const u32 ws = wid * per, we = min(T, ws + per);
          u32 p = ws + lane;
          u32 e = 0, nx = 0, badd = 0;
          if (p < we) {
            u32 lo = 0, hi = nc;                        // off[lo] <= p < off[hi]
            while (hi - lo > 1u) { const u32 mid = (lo + hi) >> 1; if (off[mid] <= p) lo = mid; else hi = mid; }
            e = lo; nx = off[e + 1]; badd = bst[e];
          }
          for (u32 st = 0; st < steps; ++st) {
            u32 v[BM_STEP], have = 0;
  #pragma unroll
            for (int k = 0; k < (int)BM_STEP; ++k) {
              v[k] = 0;
              if (p < we) {
                have |= 1u << k;
                if (p >= nx) {                          // next B row (one step for B rows of 32+ entries, a few for shorter ones)
                  do { ++e; nx = off[e + 1]; } while (p >= nx);
                  badd = bst[e];
                }
                v[k] = (u32)__ldg(&m.Bcol[(u32)(badd + p)]);
                p += 32u;
              }
            }
            u32 w[BM_STEP], bit[BM_STEP];               // bit == 0: nothing to insert
  #pragma unroll
            for (int k = 0; k < (int)BM_STEP; ++k) {
              const u32 d = v[k] - start;               // wraps to a huge value below the window (already emitted)
              w[k] = d >> 5; bit[k] = 0;
              if ((have >> k) & 1u) {
                if (d < wbits) bit[k] = 1u << (d & 31u);
                else if (v[k] >= end) above = min(above, v[k]);
              }
            }
            // pass A: plain read-modify-write (may lose concurrent updates of the same word) + summary byte
  #pragma unroll
            for (int k = 0; k < (int)BM_STEP; ++k)
              if (bit[k]) {
                const u32 o = bm[w[k]];
                if (!(o & bit[k])) bm[w[k]] = o | bit[k];
                fl[w[k] >> 2] = 1;
              }
            __syncthreads();
            // pass B: whoever lost its bit repairs it atomically (no plain store runs concurrently)
  #pragma unroll
            for (int k = 0; k < (int)BM_STEP; ++k)
              if (bit[k] && !(bm[w[k]] & bit[k])) atomicOr(&bm[w[k]], bit[k]);
            __syncthreads();
          }
        }
      }
      u32 nxt = EMPTY;
      if (comp) {
        nidx = s_idx[(it + 1u) & 1u];                   // next row, stage 1: its list entry (published before the ranking scan)
        if (nidx < n) nrow = (int)list_a[nidx];
      } else {
        above = __reduce_min_sync(FULL, above);
        if (lane == 0 && above != EMPTY) atomicMin(&s_above, above);
        if (first && tid == 0) s_idx[(it + 1u) & 1u] = nidx_reg;
        // ---- emission, step 1: the ordered list of non-empty 16-byte pieces (thread t scans summary words 3 t .. 3 t + 2)
        u32 f[3], np = 0;
  #pragma unroll
        for (int q = 0; q < 3; ++q) {
          const u32 iw = tid * 3u + (u32)q;
          f[q] = (iw < BM_FLAGW) ? (flw[iw] & 0x01010101u) : 0u;            // the walk's last barrier ordered the summary stores
          np += __popc(f[q]);
        }
        u32 pos = bm_scan(np, s_red, flip, &npieces);                      // (its barrier also publishes s_above / s_idx)
  #pragma unroll
        for (int q = 0; q < 3; ++q) {
          const u32 iw = tid * 3u + (u32)q;
          if (f[q]) flw[iw] = 0u;
          for (u32 b = f[q]; b; b &= b - 1u) plist[pos++] = (unsigned short)(iw * 4u + (((u32)__ffs((int)b) - 1u) >> 3));
        }
        nxt = s_above;
        if (first) {                                    // next row, stage 1: its list entry
          nidx = s_idx[(it + 1u) & 1u];
          if (nidx < n) nrow = (int)list_a[nidx];
        }
        __syncthreads();
      }
      {
        // steps 2 and 3 read piece i of the list at bm4[comp ? i : plist[i]]; its columns start at 128 * plist[i] (+ window start)
        if (npieces) {
          // ---- step 2: every thread sums an equal share of the list; offsets of the shares
          const u32 K = (npieces + BM_THREADS - 1u) / BM_THREADS;
          const u32 i0 = min(tid * K, npieces), i1 = min(i0 + K, npieces);
          u32 c = 0;
          for (u32 i = i0; i < i1; ++i) c += bm_popc4(bm4[comp ? i : (u32)plist[i]]);
          u32 tot;
          const u32 cbase = bm_scan(c, s_red, flip, &tot);
          if (MODE != MODE_COUNT) {
            cb[tid] = cbase;
            if (tid == BM_THREADS - 1) cb[BM_THREADS] = tot;
            __syncthreads();
            // ---- step 3: thread t writes output positions [t Q, (t+1) Q) of this window
            const u32 Q = (tot + BM_THREADS - 1u) / BM_THREADS;
            const u32 o0 = tid * Q;
            if (o0 < tot) {
              u32 left = min(Q, tot - o0);
              u32 lo = 0, hi = BM_THREADS;              // cb[lo] <= o0 < cb[hi]
              while (hi - lo > 1u) { const u32 mid = (lo + hi) >> 1; if (cb[mid] <= o0) lo = mid; else hi = mid; }
              u32 i = lo * K, run = cb[lo], k;
              uint4 x;
              while (true) { k = plist[i]; x = bm4[comp ? i : k]; const u32 pc = bm_popc4(x); if (run + pc > o0) break; run += pc; ++i; }
              // drop the bits that belong to the threads before: whole words first, then bit by bit
              u32 skip = o0 - run;
  #pragma unroll
              for (int q = 0; q < 3; ++q) {
                const u32 lowest = x.x ? x.x : x.y ? x.y : x.z;             // the lowest non-empty word of (x.x, x.y, x.z)
                const u32 pc = __popc(lowest);
                if ((x.x | x.y | x.z) && skip >= pc) { skip -= pc; if (x.x) x.x = 0u; else if (x.y) x.y = 0u; else x.z = 0u; }
              }
              for (; skip; --skip) { if (x.x) x.x &= x.x - 1u; else if (x.y) x.y &= x.y - 1u; else if (x.z) x.z &= x.z - 1u; else x.w &= x.w - 1u; }
              int* q_out = Ccol + (base + done + o0);
              // one loop of `left` iterations, the same trip count for every thread: lowest set bit of the 128-bit piece, next piece
              // when this one is used up (listed pieces are never empty)
              for (; left; --left) {
                if (!(x.x | x.y | x.z | x.w)) { ++i; k = plist[i]; x = bm4[comp ? i : k]; }
                const bool e0 = x.x != 0u, e1 = !e0 && x.y != 0u, e2 = !e0 && !e1 && x.z != 0u, e3 = !e0 && !e1 && !e2;   // selects, no branches
                const u32 word = e0 ? x.x : e1 ? x.y : e2 ? x.z : x.w;
                const u32 rest = word & (word - 1u);
                const u32 col = start + (k << 7) + (e0 ? 0u : e1 ? 32u : e2 ? 64u : 96u) + ((u32)__ffs((int)word) - 1u);
                x.x = e0 ? rest : x.x; x.y = e1 ? rest : x.y; x.z = e2 ? rest : x.z; x.w = e3 ? rest : x.w;
                *q_out++ = (int)col;
              }
            }
          }
          __syncthreads();                              // everybody has read the pieces: clear them
          for (u32 i = tid; i < npieces; i += BM_THREADS) bm4[comp ? i : (u32)plist[i]] = make_uint4(0u, 0u, 0u, 0u);
          done += tot;
        }
      }
      if (comp) { flw[2u * tid] = 0u; flw[2u * tid + 1u] = 0u; }              // the piece bitmap (its owner clears it)
      if (tid == 0) s_above = EMPTY;
      if (first && nidx < n) { na0 = m.Arow[nrow]; na1 = m.Arow[nrow + 1]; }   // next row, stage 2: its row pointers
      first = false;
      __syncthreads();
      if (nxt == EMPTY) break;
      if (nxt >= (u32)m.Bm) { bad = true; break; }    // a column of B outside [0,Bm)
      start = nxt & ~31u;
    }
    if (MODE != MODE_FILL && tid == 0) cnt[row] = (u32)done;
    if (bad) atomicOr(&sc->err, 4u);                  // (windowed passes: every thread saw it; compressed pass: the thread that held the product)
    idx = nidx; row = nrow; a0 = na0; a1 = na1; ++it;
  }
}

}  // namespace bsk
