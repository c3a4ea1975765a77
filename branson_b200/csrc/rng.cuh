// rng.cuh -- Random123 Threefry2x64-20 counter RNG, re-derived for registers.
//
// Follows the published Threefry2x64 algorithm as the reference uses it:
//   rotation constants  reference src/random123/threefry.h:86-93   {16,42,12,31,16,32,24,21}
//   key-schedule parity reference src/random123/threefry.h:170-171  0x1BD11BDAA9FC1A22
//   20 rounds, key injection after every 4th round, :196-282
//   draw -> double      reference src/RNG.h:262-285 (_ran) and :202-238 (u01fixedpt<double,uint64_t>):
//                       ((out[0] >> 11) | 1) * 2^-53, second output word discarded, counter += 1
// State per photon on the device is 16 bytes: the counter low word and the stream (key low word).  The counter high
// word (seed << 32, src/RNG.h:318-330) is one value for the whole run and the key high word is 0.
#pragma once
#include <stdint.h>

namespace bg {

// 64-bit rotate as two 32-bit funnel shifts (SHF.L.W.U32.HI); a rotation by 32 is a register swap.  The plain
// (x << n) | (x >> (64 - n)) form compiles to three alu-pipe instructions per rotation on sm_100a, and the alu pipe is
// what bounds Threefry (tools/ubench/threefry_variants.cu: +10 % draws/s with this form).
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int n) {
  const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  const uint32_t a = (n & 32) ? hi : lo, b = (n & 32) ? lo : hi;  // after the optional swap: value = b:a
  const uint32_t s = (uint32_t)n & 31u;
  if (s == 0) return ((uint64_t)b << 32) | a;
  return ((uint64_t)__funnelshift_l(a, b, s) << 32) | __funnelshift_l(b, a, s);
}

#define BG_TF_ROUND(R)   \
  x0 += x1;              \
  x1 = rotl64(x1, (R));  \
  x1 ^= x0;

// key = {stream, 0}; ctr = {ctr_lo, ctr_hi}.  Returns output word 0.
__device__ __forceinline__ uint64_t threefry2x64_20_w0(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t k0) {
  const uint64_t ks0 = k0;
  const uint64_t ks2 = 0x1BD11BDAA9FC1A22ULL ^ k0;  // ks1 == 0
  uint64_t x0 = ctr_lo + ks0;
  uint64_t x1 = ctr_hi;  // + ks1 (0)
  BG_TF_ROUND(16) BG_TF_ROUND(42) BG_TF_ROUND(12) BG_TF_ROUND(31)
  x1 += ks2 + 1;  // injection 1: x0 += ks1 (0), x1 += ks2 + 1
  BG_TF_ROUND(16) BG_TF_ROUND(32) BG_TF_ROUND(24) BG_TF_ROUND(21)
  x0 += ks2;      // injection 2: x0 += ks2, x1 += ks0 + 2
  x1 += ks0 + 2;
  BG_TF_ROUND(16) BG_TF_ROUND(42) BG_TF_ROUND(12) BG_TF_ROUND(31)
  x0 += ks0;      // injection 3: x0 += ks0, x1 += ks1 + 3
  x1 += 3;
  BG_TF_ROUND(16) BG_TF_ROUND(32) BG_TF_ROUND(24) BG_TF_ROUND(21)
  x1 += ks2 + 4;  // injection 4: x0 += ks1 (0), x1 += ks2 + 4
  BG_TF_ROUND(16) BG_TF_ROUND(42) BG_TF_ROUND(12) BG_TF_ROUND(31)
  x0 += ks2;      // injection 5: x0 += ks2, x1 += ks0 + 5
  // x1 += ks0 + 5;  (second output word is discarded by the reference)
  return x0;
}

// N counters of one stream at once: ctr_lo + OFF[j].  The Threefry evaluations are independent, so the fully unrolled
// rounds interleave into N dependency chains (the single-draw form is one serial chain of ~135 integer instructions
// and leaves the ALU pipe waiting on its own latency).
template <int N>
__device__ __forceinline__ void threefry2x64_20_w0_multi(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t k0,
                                                         const int (&OFF)[N], uint64_t (&w)[N]) {
  const uint64_t ks0 = k0;
  const uint64_t ks2 = 0x1BD11BDAA9FC1A22ULL ^ k0;
  uint64_t x0[N], x1[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    x0[j] = ctr_lo + (uint64_t)OFF[j] + ks0;
    x1[j] = ctr_hi;
  }
#define BG_TF_ROUNDN(R)            \
  _Pragma("unroll") for (int j = 0; j < N; ++j) { \
    x0[j] += x1[j];                \
    x1[j] = rotl64(x1[j], (R));    \
    x1[j] ^= x0[j];                \
  }
  BG_TF_ROUNDN(16) BG_TF_ROUNDN(42) BG_TF_ROUNDN(12) BG_TF_ROUNDN(31)
#pragma unroll
  for (int j = 0; j < N; ++j) x1[j] += ks2 + 1;
  BG_TF_ROUNDN(16) BG_TF_ROUNDN(32) BG_TF_ROUNDN(24) BG_TF_ROUNDN(21)
#pragma unroll
  for (int j = 0; j < N; ++j) { x0[j] += ks2; x1[j] += ks0 + 2; }
  BG_TF_ROUNDN(16) BG_TF_ROUNDN(42) BG_TF_ROUNDN(12) BG_TF_ROUNDN(31)
#pragma unroll
  for (int j = 0; j < N; ++j) { x0[j] += ks0; x1[j] += 3; }
  BG_TF_ROUNDN(16) BG_TF_ROUNDN(32) BG_TF_ROUNDN(24) BG_TF_ROUNDN(21)
#pragma unroll
  for (int j = 0; j < N; ++j) x1[j] += ks2 + 4;
  BG_TF_ROUNDN(16) BG_TF_ROUNDN(42) BG_TF_ROUNDN(12) BG_TF_ROUNDN(31)
#pragma unroll
  for (int j = 0; j < N; ++j) w[j] = x0[j] + ks2;
#undef BG_TF_ROUNDN
}

// Four consecutive counters: ctr_lo + 0..3.
__device__ __forceinline__ void threefry2x64_20_w0_x4(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t k0, uint64_t w[4]) {
  const int off[4] = {0, 1, 2, 3};
  uint64_t v[4];
  threefry2x64_20_w0_multi<4>(ctr_lo, ctr_hi, k0, off, v);
#pragma unroll
  for (int j = 0; j < 4; ++j) w[j] = v[j];
}

// General form (arbitrary key high word), used by the known-answer test kernel.
__device__ __forceinline__ void threefry2x64_20(const uint64_t ctr[2], const uint64_t key[2], uint64_t out[2]) {
  const int R[8] = {16, 42, 12, 31, 16, 32, 24, 21};
  uint64_t ks[3] = {key[0], key[1], 0x1BD11BDAA9FC1A22ULL ^ key[0] ^ key[1]};
  uint64_t x0 = ctr[0] + ks[0], x1 = ctr[1] + ks[1];
#pragma unroll
  for (int r = 0; r < 20; ++r) {
    x0 += x1;
    x1 = rotl64(x1, R[r & 7]);
    x1 ^= x0;
    if ((r & 3) == 3) {
      const int j = (r >> 2) + 1;
      x0 += ks[j % 3];
      x1 += ks[(j + 1) % 3] + (uint64_t)j;
    }
  }
  out[0] = x0;
  out[1] = x1;
}

__device__ __forceinline__ double u01_from_bits(uint64_t w) {
  return __ull2double_rn((w >> 11) | 1ULL) * (1.0 / 9007199254740992.0);
}

// One draw: advances the counter low word (the 128-bit carry into ctr_hi needs 2^64 draws of one photon).
__device__ __forceinline__ double rng_next(uint64_t &ctr_lo, const uint64_t ctr_hi, const uint64_t stream) {
  const uint64_t w = threefry2x64_20_w0(ctr_lo, ctr_hi, stream);
  ctr_lo += 1;
  return u01_from_bits(w);
}

}  // namespace bg
