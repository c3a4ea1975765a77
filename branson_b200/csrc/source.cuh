// source.cuh -- photon creation on the device.
//
// Replaces the serial host loops make_photons (reference src/source.h:212-366) and make_initial_census_photons
// (:138-204).  The reference walks cells in mesh order and gives the k-th photon it creates the RNG stream
// `cycle_offset + rank_offset + k` (:221-222,:267,:283; initial census: rank_offset + k, :144,:171).  Here:
//   1. k_source_count   per cell: n = max(1, int(n_user * E_cell / total_E)) for E_cell > 0 (:230-233), the same
//                       IEEE double expression, so counts are identical;
//   2. exclusive scan   over the cell-major [emission, source] count pairs -> first photon index of every entry;
//   3. k_source_windows the entry of every 256th photon (one binary search per block boundary), then
//      k_source_sample  one thread per photon: binary search of its entry inside its block's window, then exactly the
//                       reference's draw order
//                       get_emission_photon (:84-97) pos x,y,z -> angle (2) -> life_dx -> group = 7 draws,
//                       get_boundary_source_photon (:100-114) face pos (2) -> cosine-law angle (2) -> life_dx -> group,
//                       get_initial_census_photon (:117-131) pos (3) -> angle (2) -> group, life_dx = c*dt.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"
#include "rng.cuh"

namespace bg {

__device__ __forceinline__ uint32_t photons_for_cell(uint64_t n_user, double E, double total_E) {
  // uint32_t t = int(n_user_photons * E / total_E); if (t == 0) t = 1;   (src/source.h:230-233)
  uint32_t n = (uint32_t)(int)((double)n_user * E / total_E);
  return n == 0 ? 1u : n;
}

// entries: [2*cell] = emission photons, [2*cell+1] = boundary-source photons (kinds == 2)
//          [cell]   = initial census photons (kinds == 1)
__global__ void k_source_count(uint32_t n_cells, int kinds, const double *__restrict__ E0,
                               const double *__restrict__ E1, uint64_t n_user, double total_E,
                               uint32_t *__restrict__ counts) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cells) return;
  const double e0 = E0[i];
  const uint32_t c0 = (e0 > 0.0) ? photons_for_cell(n_user, e0, total_E) : 0u;
  if (kinds == 1) {
    counts[i] = c0;
  } else {
    const double e1 = E1[i];
    const uint32_t c1 = (e1 > 0.0) ? photons_for_cell(n_user, e1, total_E) : 0u;
    reinterpret_cast<uint2 *>(counts)[i] = make_uint2(c0, c1);
  }
}

struct SourceParams {
  PhotonSoA ph;
  uint64_t dst_offset;   // first slot in ph to write
  uint64_t n;            // photons to make
  const uint64_t *offsets;  // exclusive scan of counts, n_entries + 1 values
  const uint32_t *first_entry;  // k_source_windows: entry of the first photon of every SOURCE_BLOCK photons (+ the last)
  uint32_t n_entries;
  int kinds;             // 2: emission + boundary source, 1: initial census
  const double *E0, *E1;
  MeshDev mesh;
  uint64_t ctr_hi;
  uint64_t stream_base;  // cycle offset + rank offset
  double dt;
};

// The (at most seven) draws of one new photon use the consecutive counters 0..6 of its own stream: they are evaluated
// up front as a batch of four and a batch of three interleaved Threefry chains (rng.cuh; w[7] is never handed out) and
// handed out in the reference's draw order.
struct Draws {
  uint64_t w[8];
  uint32_t used;
  // `used` only ever holds values the compiler can enumerate, and the select chain keeps w[] in registers (a
  // dynamically indexed array would live in local memory)
  __device__ __forceinline__ double next() {
    const uint32_t u = used++;
    const uint64_t v = (u == 0) ? w[0] : (u == 1) ? w[1] : (u == 2) ? w[2] : (u == 3) ? w[3] : (u == 4) ? w[4]
                     : (u == 5) ? w[5] : (u == 6) ? w[6] : w[7];
    return u01_from_bits(v);
  }
};

__device__ __forceinline__ void uniform_angle(Draws &D, double &ax, double &ay, double &az) {
  // src/sampling_functions.h:57-70
  const double mu = D.next() * 2.0 - 1.0;
  const double phi = D.next() * 2.0 * K_PI;
  // the same routines as the scatter of the history loop (transport.cuh scatter_direction): mu in (-1, 1), phi in [0, 2 pi)
  const double sin_theta = fm_sqrt(1.0 - mu * mu);
  double sp, cp;
  fm_sincos(phi, &sp, &cp);
  ax = sin_theta * cp;
  ay = sin_theta * sp;
  az = mu;
}

// lower-bound search: entry e with offsets[e] <= k < offsets[e+1], inside [lo, hi)
__device__ __forceinline__ uint32_t find_entry(const uint64_t *__restrict__ offsets, uint32_t lo, uint32_t hi, uint64_t k) {
  while (hi - lo > 1) {  // invariant: offsets[lo] <= k < offsets[hi]
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(&offsets[mid]) <= k) lo = mid; else hi = mid;
  }
  return lo;
}

// entry of the first photon of every block of k_source_sample (and, in the last slot, of the last photon): a block's
// photons are consecutive, so their entries lie inside [first[b], first[b + 1]] and every thread searches only that
// window.  One thread per block boundary, a plain binary search over the whole table (a few microseconds in total): the
// sampling kernel itself then needs no block-wide search and no barrier (the barrier behind the in-kernel window search
// was 32 % of its stall samples, profiles/source_r02_v18_ncu.md).
__global__ void k_source_windows(const uint64_t *__restrict__ offsets, uint32_t n_entries, uint64_t n, uint32_t block,
                                 uint32_t n_blocks, uint32_t *__restrict__ first) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_blocks) return;
  uint64_t k = (uint64_t)b * block;
  if (k > n - 1) k = n - 1;
  first[b] = find_entry(offsets, 0, n_entries, k);
}

constexpr int SOURCE_BLOCK = 256;

__global__ void __launch_bounds__(SOURCE_BLOCK) k_source_sample(const SourceParams P) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.n) return;
  const uint32_t lo = find_entry(P.offsets, __ldg(&P.first_entry[blockIdx.x]), __ldg(&P.first_entry[blockIdx.x + 1]) + 1, k);
  const uint32_t entry = lo;
  const uint32_t cell = (P.kinds == 2) ? (entry >> 1) : entry;
  const bool boundary_source = (P.kinds == 2) && (entry & 1u);
  const uint32_t n_in_entry = (uint32_t)(__ldg(&P.offsets[entry + 1]) - __ldg(&P.offsets[entry]));
  const double E_cell = boundary_source ? P.E1[cell] : P.E0[cell];
  const double E = E_cell / n_in_entry;

  const uint32_t nx = P.mesh.nx, ny = P.mesh.ny;
  const uint32_t sxy = nx * ny;
  const uint32_t kk = cell / sxy, rem = cell - kk * sxy, jj = rem / nx, ii = rem - jj * nx;
  const double *fx = P.mesh.faces, *fy = fx + (nx + 1), *fz = fy + (ny + 1);
  const double x0 = __ldg(&fx[ii]), x1 = __ldg(&fx[ii + 1]);
  const double y0 = __ldg(&fy[jj]), y1 = __ldg(&fy[jj + 1]);
  const double z0 = __ldg(&fz[kk]), z1 = __ldg(&fz[kk + 1]);

  const uint64_t stream = P.stream_base + k;
  Draws D;
  threefry2x64_20_w0_x4(0, P.ctr_hi, stream, D.w);
  {
    const int off[3] = {4, 5, 6};
    uint64_t v[3];
    threefry2x64_20_w0_multi<3>(0, P.ctr_hi, stream, off, v);
    D.w[4] = v[0]; D.w[5] = v[1]; D.w[6] = v[2]; D.w[7] = 0;
  }
  D.used = 0;
  double x, y, z, ax, ay, az, life;
  if (!boundary_source) {
    // get_uniform_position_in_cell (src/sampling_functions.h:22-29)
    x = x0 + D.next() * (x1 - x0);
    y = y0 + D.next() * (y1 - y0);
    z = z0 + D.next() * (z1 - z0);
    uniform_angle(D, ax, ay, az);
  } else {
    // Cell::get_source_face (src/cell.h:69-76): first face whose bc is SOURCE
    int face = -1;
    {
      const bool on[6] = {ii == 0, ii == nx - 1, jj == 0, jj == ny - 1, kk == 0, kk == P.mesh.nz - 1};
#pragma unroll
      for (int s = 5; s >= 0; --s)
        if (on[s] && P.mesh.bc[s] == BC_SOURCE) face = s;
    }
    // get_uniform_position_on_face (src/sampling_functions.h:32-52)
    if (face == 0 || face == 1) {
      x = (face == 0) ? x0 : x1;
      y = y0 + D.next() * (y1 - y0);
      z = z0 + D.next() * (z1 - z0);
    } else if (face == 2 || face == 3) {
      x = x0 + D.next() * (x1 - x0);
      y = (face == 2) ? y0 : y1;
      z = z0 + D.next() * (z1 - z0);
    } else {
      x = x0 + D.next() * (x1 - x0);
      y = y0 + D.next() * (y1 - y0);
      z = (face == 4) ? z0 : z1;
    }
    // get_source_angle_on_face (src/sampling_functions.h:94-121), signs exactly as the reference has them
    const double theta = acos(sqrt(D.next()));
    const double phi = D.next() * 2.0 * K_PI;
    const double sign = (face % 2) ? -1.0 : 1.0;
    double st, ct, sp, cp;
    sincos(theta, &st, &ct);
    sincos(phi, &sp, &cp);
    if (face == 0 || face == 1) { ax = ct * sign; ay = st * sp; az = st * cp; }
    else if (face == 2 || face == 3) { ax = st * sp; ay = ct; az = st * cp; }
    else { ax = st * cp; ay = st * sp; az = ct; }
  }
  if (P.kinds == 2) life = D.next() * K_C * P.dt;
  else life = K_C * P.dt;
  const uint32_t group = (uint32_t)floor(D.next() * (double)P.mesh.G);

  const uint64_t o = P.dst_offset + k;
  P.ph.xy[o] = make_double2(x, y);
  P.ph.za[o] = make_double2(z, ax);
  P.ph.bc[o] = make_double2(ay, az);
  P.ph.ee[o] = make_double2(E, E);
  P.ph.lc[o] = make_ulonglong2((unsigned long long)__double_as_longlong(life), (unsigned long long)D.used);
  P.ph.sg[o] = make_ulonglong2(stream, (unsigned long long)cell | ((unsigned long long)group << 32));
}

}  // namespace bg
