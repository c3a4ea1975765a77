// census.cuh -- post-transport census compaction and the scan primitives shared with sourcing.
//
// Replaces post_process_photons (reference src/post_process_functions.h:33-59): in photon order, KILLED photons are
// counted, EXIT photons add their energy to exit_E, CENSUS photons get life_dx = c * next_dt and are appended to the
// census list (stable order), their energy summed into census_E.  On the device this is flag -> scan -> scatter:
//   k_census_tiles    per 4096-photon tile: census count, fixed-order partial sums of census / exit energy
//   k_scan_partials   one CTA: exclusive scan of tile counts, in-order sum of the tile energies
//   k_census_scatter  per tile: rank within the tile + tile offset -> copy the six 16-byte streams
// The energy sums use a fixed summation tree, so they are reproducible run to run (the reference's strictly serial
// order is available through the deterministic validation mode, which sums on the host).
#pragma once
#include "common.cuh"

namespace bg {

constexpr int CT_THREADS = 256;
constexpr int CT_ITEMS = 16;  // descriptors per thread: one 128-bit load
constexpr int CT_TILE = CT_THREADS * CT_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, unsigned lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= (unsigned)o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread (CT_THREADS threads); returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *s_warp, uint32_t *total) {
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const uint32_t inc = warp_incl_scan(v, lane);
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  if (w == 0) {
    const uint32_t x = (lane < (CT_THREADS >> 5)) ? s_warp[lane] : 0u;
    const uint32_t xi = warp_incl_scan(x, lane);
    if (lane < (CT_THREADS >> 5)) s_warp[lane] = xi - x;
    if (lane == (CT_THREADS >> 5) - 1) s_warp[CT_THREADS >> 5] = xi;
  }
  __syncthreads();
  const uint32_t r = s_warp[w] + inc - v;
  *total = s_warp[CT_THREADS >> 5];
  __syncthreads();
  return r;
}

// fixed-tree block sum of doubles (same tree for every launch -> reproducible)
__device__ __forceinline__ double block_sum(double v, double *s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  if (lane == 0) s_red[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < (CT_THREADS >> 5); ++i) r += s_red[i];
  }
  __syncthreads();
  return r;  // valid on thread 0
}

struct TilePartials {
  uint32_t *n_census;  // per tile
  uint32_t *n_killed;
  uint32_t *n_exit;
  double *census_E;
  double *exit_E;
  uint64_t *tile_off;  // exclusive scan of n_census, n_tiles + 1
};

__device__ __forceinline__ void load_desc16(const uint8_t *desc, uint64_t base, uint64_t n, uint8_t d[CT_ITEMS]) {
  if (base + CT_ITEMS <= n) {
    const uint4 v = *reinterpret_cast<const uint4 *>(desc + base);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < CT_ITEMS; ++i) d[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
  } else {
#pragma unroll
    for (int i = 0; i < CT_ITEMS; ++i) d[i] = (base + i < n) ? desc[base + i] : (uint8_t)EV_PASS;
  }
}

__global__ void __launch_bounds__(CT_THREADS) k_census_tiles(const uint8_t *__restrict__ desc,
                                                             const double2 *__restrict__ ee, uint64_t n,
                                                             TilePartials T) {
  __shared__ uint32_t s_warp[(CT_THREADS >> 5) + 1];
  __shared__ double s_red[CT_THREADS >> 5];
  const uint64_t base = (uint64_t)blockIdx.x * CT_TILE + (uint64_t)threadIdx.x * CT_ITEMS;
  uint8_t d[CT_ITEMS];
  load_desc16(desc, base, n, d);
  uint32_t nc = 0, nk = 0, ne = 0;
  double ce = 0.0, xe = 0.0;
#pragma unroll
  for (int i = 0; i < CT_ITEMS; ++i) {
    if (base + i < n) {
      if (d[i] == EV_CENSUS) { ++nc; ce += ee[base + i].x; }
      else if (d[i] == EV_EXIT) { ++ne; xe += ee[base + i].x; }
      else if (d[i] == EV_KILLED) ++nk;
    }
  }
  uint32_t tot;
  block_excl_scan(nc, s_warp, &tot);
  if (threadIdx.x == 0) T.n_census[blockIdx.x] = tot;
  block_excl_scan(nk, s_warp, &tot);
  if (threadIdx.x == 0) T.n_killed[blockIdx.x] = tot;
  block_excl_scan(ne, s_warp, &tot);
  if (threadIdx.x == 0) T.n_exit[blockIdx.x] = tot;
  const double bce = block_sum(ce, s_red);
  if (threadIdx.x == 0) T.census_E[blockIdx.x] = bce;
  const double bxe = block_sum(xe, s_red);
  if (threadIdx.x == 0) T.exit_E[blockIdx.x] = bxe;
}

// exclusive scan of n u32 values into n+1 u64 values by ONE CTA of 1024 threads (n = number of tiles: small)
__device__ __forceinline__ void single_cta_scan(const uint32_t *__restrict__ in, uint32_t n, uint64_t *__restrict__ out) {
  __shared__ uint64_t s_tot[32];
  __shared__ uint64_t s_carry;
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t start = 0; start < n; start += 1024) {
    const uint32_t i = start + threadIdx.x;
    const uint64_t v = (i < n) ? (uint64_t)in[i] : 0ull;
    uint64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) s_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
      const uint64_t x = s_tot[lane];
      uint64_t xi = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, xi, o);
        if (lane >= (unsigned)o) xi += t;
      }
      s_tot[lane] = xi - x;
    }
    __syncthreads();
    const uint64_t excl = s_carry + s_tot[w] + inc - v;
    if (i < n) out[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = s_carry;
}

__global__ void __launch_bounds__(1024) k_scan_single(const uint32_t *__restrict__ in, uint32_t n,
                                                      uint64_t *__restrict__ out) {
  single_cta_scan(in, n, out);
}

// results: [0] census_E, [1] exit_E ; counts -> stats[ST_N_CENSUS..]
// The tile partials are summed by a fixed tree (thread t takes tiles t, t + 1024, ... in order, then a fixed warp /
// block tree), so the result depends on the number of tiles only -- reproducible run to run.
__global__ void __launch_bounds__(1024) k_scan_partials(uint32_t n_tiles, TilePartials T, double *results,
                                                        unsigned long long *stats) {
  single_cta_scan(T.n_census, n_tiles, T.tile_off);
  __shared__ double s_ce[32], s_xe[32];
  __shared__ unsigned long long s_nk[32], s_ne[32];
  double ce = 0.0, xe = 0.0;
  unsigned long long nk = 0, ne = 0;
  for (uint32_t i = threadIdx.x; i < n_tiles; i += 1024) {
    ce += T.census_E[i];
    xe += T.exit_E[i];
    nk += T.n_killed[i];
    ne += T.n_exit[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ce += __shfl_down_sync(0xffffffffu, ce, o);
    xe += __shfl_down_sync(0xffffffffu, xe, o);
    nk += __shfl_down_sync(0xffffffffu, nk, o);
    ne += __shfl_down_sync(0xffffffffu, ne, o);
  }
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  if (lane == 0) { s_ce[w] = ce; s_xe[w] = xe; s_nk[w] = nk; s_ne[w] = ne; }
  __syncthreads();
  if (threadIdx.x == 0) {
    ce = 0.0; xe = 0.0; nk = 0; ne = 0;
    for (int i = 0; i < 32; ++i) { ce += s_ce[i]; xe += s_xe[i]; nk += s_nk[i]; ne += s_ne[i]; }
    results[0] = ce;
    results[1] = xe;
    stats[ST_N_CENSUS] = T.tile_off[n_tiles];
    stats[ST_N_KILLED] = nk;
    stats[ST_N_EXIT] = ne;
  }
}

// Stable compaction of one tile: the ranks of the tile's CENSUS photons come from the same block scan as before; their
// source indices are staged in shared memory in rank order, and the six 16-byte streams are then copied with one
// thread per census photon, so the stores are fully coalesced (consecutive ranks -> consecutive addresses) and the
// loads touch each surviving photon's sectors once.
__global__ void __launch_bounds__(CT_THREADS) k_census_scatter(const uint8_t *__restrict__ desc, PhotonSoA src,
                                                               uint64_t n, PhotonSoA dst, uint64_t dst_offset,
                                                               const uint64_t *__restrict__ tile_off,
                                                               double census_life_dx) {
  __shared__ uint32_t s_warp[(CT_THREADS >> 5) + 1];
  __shared__ uint16_t s_idx[CT_TILE];  // offset inside the tile of the r-th census photon
  const uint64_t tile_base = (uint64_t)blockIdx.x * CT_TILE;
  const uint32_t tbase = threadIdx.x * CT_ITEMS;
  const uint64_t base = tile_base + tbase;
  uint8_t d[CT_ITEMS];
  load_desc16(desc, base, n, d);
  uint32_t nc = 0;
#pragma unroll
  for (int i = 0; i < CT_ITEMS; ++i) nc += (base + i < n && d[i] == EV_CENSUS) ? 1u : 0u;
  uint32_t tot;
  uint32_t r = block_excl_scan(nc, s_warp, &tot);
  if (tot == 0) return;
#pragma unroll
  for (int i = 0; i < CT_ITEMS; ++i)
    if (base + i < n && d[i] == EV_CENSUS) s_idx[r++] = (uint16_t)(tbase + i);
  __syncthreads();
  const uint64_t o0 = dst_offset + tile_off[blockIdx.x];
  const unsigned long long life_bits = (unsigned long long)__double_as_longlong(census_life_dx);
  for (uint32_t j = threadIdx.x; j < tot; j += CT_THREADS) {
    const uint64_t sidx = tile_base + s_idx[j], o = o0 + j;
    const double2 xy = src.xy[sidx], za = src.za[sidx], bc = src.bc[sidx], ee = src.ee[sidx];
    const ulonglong2 lc = src.lc[sidx], sg = src.sg[sidx];
    dst.xy[o] = xy;
    dst.za[o] = za;
    dst.bc[o] = bc;
    dst.ee[o] = ee;
    dst.lc[o] = make_ulonglong2(life_bits, lc.y);  // life_dx = c * next_dt (src/post_process_functions.h:49)
    dst.sg[o] = sg;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// generic exclusive scan u32 -> u64 (counts of the source entries, deposits per photon)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CT_THREADS) k_scan_tile_sums(const uint32_t *__restrict__ in, uint64_t n,
                                                               uint32_t *__restrict__ tile_sum) {
  __shared__ uint32_t s_warp[(CT_THREADS >> 5) + 1];
  const uint64_t base = (uint64_t)blockIdx.x * CT_TILE + (uint64_t)threadIdx.x * CT_ITEMS;
  uint32_t v = 0;
#pragma unroll
  for (int i = 0; i < CT_ITEMS; ++i) v += (base + i < n) ? in[base + i] : 0u;
  uint32_t tot;
  block_excl_scan(v, s_warp, &tot);
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(CT_THREADS) k_scan_apply(const uint32_t *__restrict__ in, uint64_t n,
                                                           const uint64_t *__restrict__ tile_off,
                                                           uint64_t *__restrict__ out) {
  __shared__ uint32_t s_warp[(CT_THREADS >> 5) + 1];
  const uint64_t base = (uint64_t)blockIdx.x * CT_TILE + (uint64_t)threadIdx.x * CT_ITEMS;
  uint32_t x[CT_ITEMS];
  uint32_t v = 0;
#pragma unroll
  for (int i = 0; i < CT_ITEMS; ++i) {
    x[i] = (base + i < n) ? in[base + i] : 0u;
    v += x[i];
  }
  uint32_t tot;
  const uint32_t r = block_excl_scan(v, s_warp, &tot);
  uint64_t run = tile_off[blockIdx.x] + r;
#pragma unroll
  for (int i = 0; i < CT_ITEMS; ++i) {
    if (base + i < n) out[base + i] = run;
    run += x[i];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == CT_THREADS - 1) out[n] = tile_off[gridDim.x];
}

// Census ordered by cell: sort key and identity permutation ...
__global__ void k_cell_keys(const ulonglong2 *__restrict__ sg, uint64_t n, uint32_t *__restrict__ key,
                            uint32_t *__restrict__ index) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  key[i] = (uint32_t)sg[i].y;
  index[i] = (uint32_t)i;
}

// ... and dst[o] = src[order[o]]: gathered 16-byte loads, coalesced stores.
__global__ void k_permute_soa(PhotonSoA src, uint64_t n, const uint32_t *__restrict__ order, PhotonSoA dst) {
  const uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  const uint32_t i = order[o];
  dst.xy[o] = src.xy[i];
  dst.za[o] = src.za[i];
  dst.bc[o] = src.bc[i];
  dst.ee[o] = src.ee[i];
  dst.lc[o] = src.lc[i];
  dst.sg[o] = src.sg[i];
}

}  // namespace bg
