// tools/microbench.cu — shared-memory primitive throughput on sm_100a (design input for the ordered-table
// kernels: is ATOMS.MIN cheap enough, or should the insert be LDS/STS based?).  Not part of the product.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu && ./microbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int WARPS = 8, TABW = 1024, ITERS = 4096;

__device__ __forceinline__ uint32_t mixu(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int OP>
__global__ void __launch_bounds__(WARPS * 32) k(uint32_t* out, long long* clk) {
  __shared__ uint32_t tab[WARPS][TABW];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < TABW; i += 32) tab[w][i] = 0xffffffffu;
  __syncwarp();
  long long c0 = clock64();
  uint32_t acc = 0, x = mixu(threadIdx.x + blockIdx.x * 977u);
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
    x = x * 1664525u + 1013904223u;
    const uint32_t s = (x >> 8) & (TABW - 1);
    if (OP == 0) acc += atomicMin(&tab[w][s], x);                       // ATOMS.MIN random slot
    else if (OP == 1) acc += atomicOr(&tab[w][s], 1u << (x & 31));      // ATOMS.OR
    else if (OP == 2) { acc += tab[w][s]; }                             // LDS random
    else if (OP == 3) { tab[w][s] = x; }                                // STS random
    else if (OP == 4) { uint32_t o = tab[w][s]; if (o > x) tab[w][s] = x; acc += o; }   // LDS+STS
    else if (OP == 5) acc += atomicMin(&tab[w][(i * 32 + lane) & (TABW - 1)], x);       // ATOMS.MIN conflict-free
    else if (OP == 6) acc += atomicCAS(&tab[w][s], 0xffffffffu, x);     // ATOMS.CAS
    else if (OP == 7) { atomicMin(&tab[w][s], x); }                     // ATOMS.MIN no return (RED-like)
  }
  long long c1 = clock64();
  if (acc == 0x12345678u) out[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = c1 - c0;
}

template <int OP> void run(const char* name, int sms) {
  uint32_t* out; long long* clk; cudaMalloc(&out, 4); cudaMalloc(&clk, 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = sms * 4;
  k<OP><<<grid, WARPS * 32>>>(out, clk); cudaDeviceSynchronize();
  cudaEventRecord(a); k<OP><<<grid, WARPS * 32>>>(out, clk); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long cyc; cudaMemcpy(&cyc, clk, 8, cudaMemcpyDeviceToHost);
  const double warp_ops_per_sm = 4.0 * WARPS * ITERS;      // 4 CTAs/SM resident (32 warps/SM)
  printf("%-28s %8.3f ms  block0 cycles %10lld  -> %.2f cycles per warp-op per SM (32 warps/SM resident)\n", name, ms, cyc,
         (double)cyc / warp_ops_per_sm);
  cudaFree(out); cudaFree(clk);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs, smem/SM %zu\n", p.name, p.multiProcessorCount, p.sharedMemPerMultiprocessor);
  run<2>("LDS random", p.multiProcessorCount);
  run<3>("STS random", p.multiProcessorCount);
  run<4>("LDS+cond STS random", p.multiProcessorCount);
  run<0>("ATOMS.MIN random", p.multiProcessorCount);
  run<7>("ATOMS.MIN random noret", p.multiProcessorCount);
  run<5>("ATOMS.MIN conflict-free", p.multiProcessorCount);
  run<1>("ATOMS.OR random", p.multiProcessorCount);
  run<6>("ATOMS.CAS random", p.multiProcessorCount);
  return 0;
}
