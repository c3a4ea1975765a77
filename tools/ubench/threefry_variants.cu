// Microbenchmark: Threefry2x64-20 formulations on sm_100a (pipe balance alu vs fma).  Not product code.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/ubench_threefry tools/ubench/threefry_variants.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t rotl_plain(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

__device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// V1: funnel shifts
template <int R>
__device__ __forceinline__ uint64_t rotl_fs(uint64_t x) {
  uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  if (R == 32) return pack(hi, lo);
  if (R < 32) return pack(__funnelshift_l(hi, lo, R), __funnelshift_l(lo, hi, R));
  // R > 32: swap then rotate by R-32
  return pack(__funnelshift_l(lo, hi, R - 32), __funnelshift_l(hi, lo, R - 32));
}

// V2: x1 = rotl(x1,R) ^ x0 with two wide multiplies and two 3-input LOP3
template <int R>
__device__ __forceinline__ uint64_t rotxor_mul(uint64_t x1, uint64_t x0) {
  uint32_t lo = (uint32_t)x1, hi = (uint32_t)(x1 >> 32);
  if (R == 32) return pack(hi, lo) ^ x0;
  if (R > 32) { uint32_t t = lo; lo = hi; hi = t; }
  constexpr int S = R & 31;
  uint64_t P, Q;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(P) : "r"(lo), "r"(1u << S));
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(Q) : "r"(hi), "r"(1u << S));
  const uint32_t nlo = ((uint32_t)P | (uint32_t)(Q >> 32)) ^ (uint32_t)x0;
  const uint32_t nhi = ((uint32_t)Q | (uint32_t)(P >> 32)) ^ (uint32_t)(x0 >> 32);
  return pack(nlo, nhi);
}

// 64-bit add on the fma pipe: IMAD.WIDE.U32 (x1.lo * one + x0) then the high word; `one` is a run-time 1 so that
// ptxas cannot fold the multiply back into IADD3 (alu pipe)
__device__ uint32_t g_one;
__device__ __forceinline__ uint64_t add_wide(uint64_t x0, uint64_t x1) {
  const uint32_t one = g_one;
  uint64_t t;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"((uint32_t)x1), "r"(one), "l"(x0));
  uint32_t hi = (uint32_t)(t >> 32), add = (uint32_t)(x1 >> 32);
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(add), "r"(one), "r"(hi));
  return pack((uint32_t)t, hi);
}

// rotl + xor with ONE 32-bit half formed on the fma pipe: (hi << S) | (lo >> (32 - S)) = hi * 2^S + mulhi(lo, 2^S),
// the multiplier 2^S read from constant memory so that ptxas cannot turn the multiplies back into alu-pipe shifts
__constant__ uint32_t c_pow[32];
template <int R, int HALVES>
__device__ __forceinline__ uint64_t rotxor_imad(uint64_t x1, uint64_t x0) {
  uint32_t lo = (uint32_t)x1, hi = (uint32_t)(x1 >> 32);
  if (R == 32) return pack(hi, lo) ^ x0;
  if (R > 32) { uint32_t t = lo; lo = hi; hi = t; }
  constexpr int S = R & 31;
  const uint32_t c = c_pow[S];
  uint32_t nhi, nlo;
  {
    uint32_t t;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(lo), "r"(c));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(nhi) : "r"(hi), "r"(c), "r"(t));
  }
  if (HALVES == 2) {
    uint32_t t;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(hi), "r"(c));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(nlo) : "r"(lo), "r"(c), "r"(t));
  } else {
    nlo = __funnelshift_l(hi, lo, S);
  }
  return pack(nlo ^ (uint32_t)x0, nhi ^ (uint32_t)(x0 >> 32));
}

template <int V, int R>
__device__ __forceinline__ void tf_round(uint64_t &x0, uint64_t &x1) {
  if (V == 8) { x0 += x1; x1 = rotxor_imad<R, 1>(x1, x0); return; }
  if (V == 9) { x0 += x1; x1 = rotxor_imad<R, 2>(x1, x0); return; }
  if (V == 4) { x0 = add_wide(x0, x1); x1 = rotl_fs<R>(x1); x1 ^= x0; return; }
  if (V == 5) { x0 = add_wide(x0, x1); x1 = rotxor_mul<R>(x1, x0); return; }
  if (V == 6) {  // B / C alternating
    if (R == 16 || R == 12 || R == 24 || R == 42) { x0 += x1; x1 = rotxor_mul<R>(x1, x0); }
    else { x0 = add_wide(x0, x1); x1 = rotl_fs<R>(x1); x1 ^= x0; }
    return;
  }
  if (V == 7) {  // A / B alternating
    if (R == 16 || R == 12 || R == 24 || R == 42) { x0 += x1; } else { x0 = add_wide(x0, x1); }
    x1 = rotl_fs<R>(x1); x1 ^= x0;
    return;
  }
  x0 += x1;
  if (V == 0) { x1 = rotl_plain(x1, R); x1 ^= x0; }
  else if (V == 1) { x1 = rotl_fs<R>(x1); x1 ^= x0; }
  else if (V == 2) { x1 = rotxor_mul<R>(x1, x0); }
  else { // V3: alternate by rotation constant parity of position: use mul for R in {16,12,24} else funnel
    if (R == 16 || R == 12 || R == 24 || R == 42) x1 = rotxor_mul<R>(x1, x0);
    else { x1 = rotl_fs<R>(x1); x1 ^= x0; }
  }
}

template <int V, int N>
__device__ __forceinline__ void threefry_xN(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t k0, uint64_t w[N]) {
  const uint64_t ks0 = k0, ks2 = 0x1BD11BDAA9FC1A22ULL ^ k0;
  uint64_t x0[N], x1[N];
#pragma unroll
  for (int j = 0; j < N; ++j) { x0[j] = ctr_lo + (uint64_t)j + ks0; x1[j] = ctr_hi; }
#define RND4(a, b, c, d)                                   \
  _Pragma("unroll") for (int j = 0; j < N; ++j) {          \
    tf_round<V, a>(x0[j], x1[j]); tf_round<V, b>(x0[j], x1[j]); \
    tf_round<V, c>(x0[j], x1[j]); tf_round<V, d>(x0[j], x1[j]); }
  RND4(16, 42, 12, 31)
#pragma unroll
  for (int j = 0; j < N; ++j) x1[j] += ks2 + 1;
  RND4(16, 32, 24, 21)
#pragma unroll
  for (int j = 0; j < N; ++j) { x0[j] += ks2; x1[j] += ks0 + 2; }
  RND4(16, 42, 12, 31)
#pragma unroll
  for (int j = 0; j < N; ++j) { x0[j] += ks0; x1[j] += 3; }
  RND4(16, 32, 24, 21)
#pragma unroll
  for (int j = 0; j < N; ++j) x1[j] += ks2 + 4;
  RND4(16, 42, 12, 31)
#pragma unroll
  for (int j = 0; j < N; ++j) w[j] = x0[j] + ks2;
#undef RND4
}

template <int V, int N>
__global__ void __launch_bounds__(128) k_bench(uint64_t *out, int iters, uint64_t seed) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t ctr = 0, acc = 0;
  const uint64_t stream = tid * 7919 + seed;
  for (int it = 0; it < iters; ++it) {
    uint64_t w[N];
    threefry_xN<V, N>(ctr, 777ull << 32, stream, w);
#pragma unroll
    for (int j = 0; j < N; ++j) acc ^= w[j];
    ctr += N + (acc & 1);   // data-dependent counter: the next batch depends on this one (like the photon loop)
  }
  out[tid] = acc;
}

template <int V, int N>
static void run(const char *name, int blocks_per_sm, uint64_t *d_out, uint64_t *h_ref) {
  const int iters = 4096 / N;
  const int blocks = 148 * blocks_per_sm;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_bench<V, N><<<blocks, 128>>>(d_out, iters, 1);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) k_bench<V, N><<<blocks, 128>>>(d_out, iters, 1);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  uint64_t h[4];
  cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
  bool ok = true;
  if (V == 0 && N == 1 && blocks_per_sm == 4) for (int i = 0; i < 4; ++i) h_ref[i] = h[i];
  if (N == 1) for (int i = 0; i < 4; ++i) ok = ok && (h[i] == h_ref[i]);
  const double draws = 5.0 * blocks * 128.0 * iters * N;
  const double per_sm_clk = draws / (ms * 1e-3) / 148.0 / 1.965e9;
  printf("%-28s N=%d blocks/SM=%d  %8.3f Gdraws/s  %.4f draws/clk/SM  => %.1f clk/draw/SMSP-warp %s\n", name, N, blocks_per_sm,
         draws / ms * 1e-6, per_sm_clk, 128.0 / per_sm_clk / 4.0 / 1.0, N == 1 ? (ok ? "match" : "MISMATCH") : "");
}

int main() {
  { uint32_t one = 1; cudaMemcpyToSymbol(g_one, &one, sizeof one); }
  { uint32_t pw[32]; for (int i = 0; i < 32; ++i) pw[i] = 1u << i; cudaMemcpyToSymbol(c_pow, pw, sizeof pw); }
  uint64_t *d_out; cudaMalloc(&d_out, 148 * 16 * 128 * sizeof(uint64_t));
  uint64_t ref[4] = {0, 0, 0, 0};
  for (int bps : {4, 8}) {
    run<0, 1>("V0 plain", bps, d_out, ref); run<1, 1>("V1 funnel", bps, d_out, ref);
    run<2, 1>("V2 mulwide", bps, d_out, ref); run<3, 1>("V3 mixed", bps, d_out, ref);
    run<0, 4>("V0 plain", bps, d_out, ref); run<1, 4>("V1 funnel", bps, d_out, ref);
    run<2, 4>("V2 mulwide", bps, d_out, ref); run<3, 4>("V3 mixed", bps, d_out, ref);
    run<1, 5>("V1 funnel", bps, d_out, ref); run<3, 5>("V3 mixed", bps, d_out, ref);
    run<1, 8>("V1 funnel", bps, d_out, ref); run<3, 8>("V3 mixed", bps, d_out, ref);
    run<4, 1>("V4 wideadd+funnel", bps, d_out, ref); run<4, 4>("V4 wideadd+funnel", bps, d_out, ref);
    run<5, 1>("V5 wideadd+mulrot", bps, d_out, ref); run<5, 4>("V5 wideadd+mulrot", bps, d_out, ref);
    run<6, 1>("V6 B/C alt", bps, d_out, ref); run<6, 4>("V6 B/C alt", bps, d_out, ref);
    run<7, 1>("V7 A/B alt", bps, d_out, ref); run<7, 4>("V7 A/B alt", bps, d_out, ref);
    run<8, 1>("V8 one half imad", bps, d_out, ref); run<8, 4>("V8 one half imad", bps, d_out, ref);
    run<9, 1>("V9 both halves imad", bps, d_out, ref); run<9, 4>("V9 both halves imad", bps, d_out, ref);
  }
  return 0;
}
