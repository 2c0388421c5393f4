// tools/microbench_alu.cu — issue rates of the instructions the sorting-network kernel is made of (VIMNMX, SHFL.BFLY,
// IMAD, IADD3, LOP3) on sm_100a: 16 independent chains per thread, 32 resident warps per SM.  Not part of the product.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_alu microbench_alu.cu && ./microbench_alu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int WARPS = 8, ITERS = 2048, CH = 16;

template <int OP>
__global__ void __launch_bounds__(WARPS * 32) k(uint32_t* out, long long* clk, uint32_t seed, uint32_t one, uint32_t mone) {
  uint32_t x[CH], y[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) { x[c] = seed * (threadIdx.x + 1) + c * 977u; y[c] = x[c] ^ (0x9e3779b9u * (c + 1)); }
  const bool p = (threadIdx.x & seed) == 0;
  long long c0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (OP == 0) x[c] = min(x[c], y[c]) + 0;                                   // VIMNMX (+ nothing: keep a dependence chain)
      else if (OP == 1) x[c] = p ? min(x[c], y[c]) : max(x[c], y[c]);            // VIMNMX + @p VIMNMX
      else if (OP == 2) x[c] = __shfl_xor_sync(0xffffffffu, x[c], 1);            // SHFL.BFLY
      else if (OP == 3) x[c] = x[c] * one + y[c];                                // IMAD (3 registers)
      else if (OP == 4) x[c] = x[c] + y[c] + seed;                               // IADD3
      else if (OP == 5) x[c] = (x[c] & y[c]) ^ seed;                             // LOP3
      else if (OP == 6) { const uint32_t t = __shfl_xor_sync(0xffffffffu, x[c], 1); x[c] = p ? min(x[c], t) : max(x[c], t); }  // the cross-lane exchange
      else if (OP == 7) { const uint32_t lo = min(x[c], y[c]), hi = max(x[c], y[c]); x[c] = lo; y[c] = hi; }                  // the in-lane exchange
      else if (OP == 8 || (OP == 9 && c % 3 != 0) || (OP == 10 && (c & 1))) {                                                   // in-lane exchange, max formed on the FMA pipe
        const uint32_t lo = min(x[c], y[c]); uint32_t s, hi;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s) : "r"(x[c]), "r"(one), "r"(y[c]));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(lo), "r"(mone), "r"(s));
        x[c] = lo; y[c] = hi;
      }
      else if (OP == 9 || OP == 10) { const uint32_t lo = min(x[c], y[c]), hi = max(x[c], y[c]); x[c] = lo; y[c] = hi; }
      else if (OP == 11) { x[c] = min(x[c], y[c] + i); }                                                                          // VIMNMX + IADD3-or-IMAD.IADD (ptxas' choice)
    }
    if (OP == 0) {
#pragma unroll
      for (int c = 0; c < CH; ++c) y[c] ^= x[c];                                 // keep min() from being hoisted (counted below)
    }
  }
  long long c1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) acc ^= x[c] ^ y[c];
  if (acc == 0x12345678u) out[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = c1 - c0;
}

template <int OP> void run(const char* name, int sms, double ops_per_iter) {
  uint32_t* out; long long* clk; cudaMalloc(&out, 4); cudaMalloc(&clk, 8);
  const int grid = sms * 4;
  k<OP><<<grid, WARPS * 32>>>(out, clk, 3, 1, 0xffffffffu); cudaDeviceSynchronize();
  k<OP><<<grid, WARPS * 32>>>(out, clk, 3, 1, 0xffffffffu); cudaDeviceSynchronize();
  long long cyc; cudaMemcpy(&cyc, clk, 8, cudaMemcpyDeviceToHost);
  const double warp_ops_per_smsp = 8.0 * ITERS * ops_per_iter;      // 32 warps per SM = 8 per sub-partition
  printf("%-44s block0 cycles %10lld -> %.2f cycles per warp instruction per SM sub-partition\n", name, cyc, (double)cyc / warp_ops_per_smsp);
  cudaFree(out); cudaFree(clk);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<0>("VIMNMX + LOP3 (2 instr)", p.multiProcessorCount, 2 * CH);
  run<1>("VIMNMX + @p VIMNMX (2 instr)", p.multiProcessorCount, 2 * CH);
  run<2>("SHFL.BFLY", p.multiProcessorCount, CH);
  run<3>("IMAD r,r,r", p.multiProcessorCount, CH);
  run<4>("IADD3", p.multiProcessorCount, CH);
  run<5>("LOP3", p.multiProcessorCount, CH);
  run<6>("SHFL + VIMNMX + @p VIMNMX (3 instr)", p.multiProcessorCount, 3 * CH);
  run<7>("VIMNMX min + VIMNMX max (2 instr)", p.multiProcessorCount, 2 * CH);
  run<7>("exchange: min + max          (per exchange)", p.multiProcessorCount, CH);
  run<8>("exchange: min + IMAD + IMAD  (per exchange)", p.multiProcessorCount, CH);
  run<9>("exchange: 2 of 3 IMAD form   (per exchange)", p.multiProcessorCount, CH);
  run<10>("exchange: 1 of 2 IMAD form   (per exchange)", p.multiProcessorCount, CH);
  run<11>("VIMNMX + add (2 instr)", p.multiProcessorCount, 2 * CH);
  return 0;
}
