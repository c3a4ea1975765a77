// Microbenchmark: do the alu pipe (LOP3 / SHF / IADD3) and the FP64 pipe (DFMA) of an sm_100a sub-partition issue
// independently, or do they share an issue port?  Not product code.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/ubench_pipe_share tools/ubench/pipe_share.cu
// Each kernel runs ITER iterations of a block of NA independent alu chains and/or ND independent DFMA chains and/or NF
// independent IMAD chains per thread (8 chains each, so latency is hidden inside one warp); 148 x 8 CTAs of 128 threads.
// If two pipes issue independently, the mixed kernel takes max(t_a, t_b); if they share a port, t_a + t_b.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NA, int ND, int NF>
__global__ void __launch_bounds__(128) k_mix(uint32_t *out, double *outd, int iters, uint32_t seed, double dseed, uint32_t mul) {
  uint32_t a[8];
  double d[8];
  uint32_t f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = seed + threadIdx.x * 8 + j; d[j] = dseed + j; f[j] = seed * 3 + j + threadIdx.x; }
  const double m = 1.0000001, c = 1e-9;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < NA) a[j] = __funnelshift_l(a[j], a[j], 7) ^ (a[j] + 0x9e3779b9u);  // SHF + IADD3 + LOP3 (alu pipe)
        if (j < ND) d[j] = fma(d[j], m, c);                                       // DFMA
        if (j < NF) f[j] = f[j] * mul + 12345u;                                    // IMAD (fma pipe)
      }
    }
  }
  uint32_t s = 0; double sd = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { s ^= a[j] ^ f[j]; sd += d[j]; }
  if (s == 0x12345678u && sd == 1.5) { out[0] = s; outd[0] = sd; }
}

template <int NA, int ND, int NF>
float run(const char *name, uint32_t *out, double *outd) {
  const int iters = 4000, grid = 148 * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_mix<NA, ND, NF><<<grid, 128>>>(out, outd, 100, 1u, 1.0, 3u);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_mix<NA, ND, NF><<<grid, 128>>>(out, outd, iters, 1u, 1.0, 3u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  // warp instructions per SM sub-partition: 8 CTAs x 4 warps / 4 sub-partitions = 8 warps each
  const double alu = 3.0 * NA * 8 * iters * 8, dfma = 1.0 * ND * 8 * iters * 8, imad = 1.0 * NF * 8 * iters * 8;
  printf("%-28s %8.3f ms   alu %5.2f  dfma %5.2f  imad %5.2f  (10^6 warp-instr per sub-partition)\n", name, ms, alu / 1e6,
         dfma / 1e6, imad / 1e6);
  return ms;
}

int main() {
  uint32_t *out; double *outd;
  cudaMalloc(&out, 64); cudaMalloc(&outd, 64);
  const float ta = run<8, 0, 0>("alu only (8 chains)", out, outd);
  const float td = run<0, 8, 0>("dfma only (8 chains)", out, outd);
  const float tf = run<0, 0, 8>("imad only (8 chains)", out, outd);
  const float tad = run<8, 8, 0>("alu + dfma", out, outd);
  const float taf = run<8, 0, 8>("alu + imad", out, outd);
  const float tdf = run<0, 8, 8>("dfma + imad", out, outd);
  const float tall = run<8, 8, 8>("alu + dfma + imad", out, outd);
  printf("alu+dfma : %.3f  vs sum %.3f, max %.3f\n", tad, ta + td, ta > td ? ta : td);
  printf("alu+imad : %.3f  vs sum %.3f, max %.3f\n", taf, ta + tf, ta > tf ? ta : tf);
  printf("dfma+imad: %.3f  vs sum %.3f, max %.3f\n", tdf, td + tf, td > tf ? td : tf);
  printf("all three: %.3f  vs sum %.3f\n", tall, ta + td + tf);
  return 0;
}
