// binary-spgemm_b200/csrc/coo2csc.cuh — COO -> CSC/CSR on the device: the stable counting sort of final/coo2csc.c:22-64.
//
// The reference sorts the coordinate entries by their `col_coo` key with a serial histogram (:37-38), a serial scan
// (:41-46) and a serial scatter in input order (:48-56); entries of one column keep their input order, and since the
// result feeds the product as Acol, the order is part of the bit-exact contract.  A scatter through atomic cursors is not
// order-stable, so the device version is a least-significant-digit radix sort of (key = col_coo, value = row_coo), 8 bits
// per pass, every pass stable:
//   k_c2c_count    per-column counts (global atomics) -> k_scan (kernels.cuh) gives the pointer array col[0..n];
//   per pass       k_c2c_hist     digit histogram of every block's chunk of C2C_CHUNK consecutive entries, digit-major
//                  k_scan         exclusive scan over (digit, block)
//                  k_c2c_scatter  rank of an entry = entries with the same digit in earlier blocks (the scan), in earlier
//                                 warps of the block (per-warp counters, scanned over the warps), in earlier rounds of its
//                                 warp (the same counters, read before they are advanced) and in lower lanes of its round
//                                 (MATCH.ANY + popc): all in input order, no atomics in the ranking.
// ceil(log2(n) / 8) passes (3 at n = 2^22); the last pass writes only the values, straight into row[].
#pragma once
#include "kernels.cuh"

namespace bsk {

constexpr int C2C_WARPS = 8, C2C_ITEMS = 16;
constexpr u32 C2C_CHUNK = 32u * C2C_WARPS * C2C_ITEMS;     // entries per block and pass

// cnt[c] += 1 for every entry; err bit 0: a key outside [0,n) (after removing the index base)
__global__ void __launch_bounds__(256) k_c2c_count(const u32* __restrict__ col_coo, u32 nnz, u32 n, u32 base, u32* __restrict__ cnt, u32* err) {
  u32 bad = 0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (size_t)gridDim.x * blockDim.x) {
    const u32 c = col_coo[e] - base;
    if (c < n) atomicAdd(&cnt[c], 1u); else bad = 1;
  }
  if (bad) atomicOr(err, 1u);
}

__global__ void __launch_bounds__(32 * C2C_WARPS) k_c2c_hist(const u32* __restrict__ keys, u32 nnz, u32 base, int shift, u32* __restrict__ hist, u32 nblocks) {
  __shared__ u32 s_cnt[256];
  s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const size_t first = (size_t)blockIdx.x * C2C_CHUNK;
#pragma unroll 4
  for (u32 i = threadIdx.x; i < C2C_CHUNK; i += 32 * C2C_WARPS) {
    const size_t e = first + i;
    if (e < nnz) atomicAdd(&s_cnt[((keys[e] - base) >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// pos[]: exclusive scan of hist (k_scan output: pos[i] = sum of hist[0..i)).  keys_out may be null (last pass).
__global__ void __launch_bounds__(32 * C2C_WARPS) k_c2c_scatter(const u32* __restrict__ keys_in, const u32* __restrict__ vals_in, u32 nnz,
                                                                u32 kbase, u32 vbase, int shift, const int* __restrict__ pos, u32 nblocks,
                                                                u32* __restrict__ keys_out, u32* __restrict__ vals_out) {
  __shared__ u32 s_cnt[C2C_WARPS][256];
  const u32 w = threadIdx.x >> 5, lane = lane_id();
  for (u32 i = threadIdx.x; i < C2C_WARPS * 256; i += 32 * C2C_WARPS) (&s_cnt[0][0])[i] = 0;
  __syncthreads();
  const size_t first = (size_t)blockIdx.x * C2C_CHUNK + (size_t)w * (32u * C2C_ITEMS);
  u32 key[C2C_ITEMS], val[C2C_ITEMS], rank[C2C_ITEMS];
#pragma unroll
  for (int i = 0; i < C2C_ITEMS; ++i) {
    const size_t e = first + (size_t)i * 32u + lane;
    const bool valid = e < nnz;
    key[i] = valid ? keys_in[e] - kbase : 0u;
    val[i] = valid ? vals_in[e] - vbase : 0u;
    const u32 d = (key[i] >> shift) & 255u;
    const u32 m = __match_any_sync(0xffffffffu, valid ? d : 0x100u + lane);      // lanes of this round with my digit
    const u32 leader = (u32)__ffs((int)m) - 1u;
    u32 old = 0;
    if (valid && lane == leader) { old = s_cnt[w][d]; s_cnt[w][d] = old + (u32)__popc(m); }
    __syncwarp();                                                                  // the next round's leader may be another lane
    rank[i] = __shfl_sync(0xffffffffu, old, leader) + (u32)__popc(m & ((1u << lane) - 1u));
  }
  __syncthreads();
  {                                       // thread d: counts of the warps -> where each warp's entries of digit d start
    const u32 d = threadIdx.x;
    u32 run = (u32)pos[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < C2C_WARPS; ++ww) { const u32 t = s_cnt[ww][d]; s_cnt[ww][d] = run; run += t; }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < C2C_ITEMS; ++i) {
    const size_t e = first + (size_t)i * 32u + lane;
    if (e < nnz) {
      const u32 p = s_cnt[w][(key[i] >> shift) & 255u] + rank[i];
      if (keys_out) keys_out[p] = key[i];
      vals_out[p] = val[i];
    }
  }
}

}  // namespace bsk
