// mesh_dev.cuh -- the per-cycle mesh physics on the device.
//
// Mesh::calculate_photon_energy (src/mesh.h:237-323) and Mesh::update_temperature (src/mesh.h:327-362) are O(n_cells)
// per-cell formulas plus a handful of sums.  Done on the host they cost five H2D / D2H array copies per cycle and a
// serial section that every rank of a replicated run repeats; here the cell state (T_e, T_r, opacities, Fleck factor,
// emission / source / census energies, tallies) never leaves HBM and only a few scalars cross PCIe per cycle.
//
// Arithmetic: every expression keeps the reference's operand order (the library is built with -fmad=false), so the
// per-cell doubles equal the host path's wherever only + - * / are involved.  Two things cannot be bit-identical to
// the reference's host run and are held to the same "last bit of libm" level as the transport kernel's log / exp:
//   * pow():  glibc's pow is correctly rounded in all but ~1e-3 of cases; CUDA's is a 2-ulp routine.  The integer
//     powers the decks use (T^3, T^4, T^-3 ...) are formed here in double-double arithmetic and rounded once
//     (correctly rounded up to double-double accuracy, 2^-100), x^0.25 as sqrt(sqrt(x)) plus one double-double Newton
//     step; only a non-integer opacity exponent falls back to CUDA's pow.
//   * the running sums: the reference adds cell by cell; here a fixed tree (1024-cell tiles, then the tiles in order).
//     The result is reproducible, identical on every rank, and differs from the serial sum only in its last bits.
// tests/test_gpu_mesh.py pins the device path against the host path (mesh.h, bit-identical to the reference) at 1e-13.
#pragma once
#include "census.cuh"
#include "common.cuh"

namespace bg {

constexpr double K_A = 0.01372;  // radiation constant, src/constants.h:20

struct RegionDev {  // src/region.h
  double opac_A, opac_B, opac_C, opac_S, cV, rho;
};

enum : int { MS_PRE_MAT = 0, MS_EMISSION, MS_CENSUS, MS_SOURCE, MS_TOTAL, MS_ABS, MS_POST_MAT, MS_N = 8 };
constexpr int MESH_TILE_THREADS = CT_THREADS;
constexpr int MESH_TILE_ITEMS = 4;
constexpr int MESH_TILE = MESH_TILE_THREADS * MESH_TILE_ITEMS;

struct MeshPhysParams {
  MeshDev mesh;
  const RegionDev *regions;
  const uint32_t *region_of_cell;
  double *T_e;          // [n_cells] material temperature (updated in place)
  const double *T_r0;   // [n_cells] initial radiation temperature (step 1 census)
  const double *T_s;    // [n_cells] source temperature (0 without a SOURCE face)
  double *T_r;          // [n_cells] radiation temperature diagnostic
  double *f;            // [n_cells]
  double *op_a, *op_s;  // [n_cells] gray values (expanded to groups by k_expand_groups)
  double *E_emission, *E_source, *E_census;  // [n_cells] this rank's share
  double *E_emission_global;                 // [n_cells] what the reference's Allreduce of m_emission_E yields
  double2 *tally;       // [n_cells] {abs_E, track_E}, already summed over ranks
  double *tile_sums;    // [MS_N][n_tiles]
  uint32_t n_tiles;
  double dt, replicated_factor, global_source_E;
  uint64_t n_user;
  uint32_t step;
  int rank, n_ranks;
};

// ---- double-double helpers -------------------------------------------------------------------------------------
struct dd {
  double h, l;
};
__device__ __forceinline__ dd dd_mul_d(dd a, double b) {
  const double p = a.h * b;
  double e = fma(a.h, b, -p);
  e = fma(a.l, b, e);
  const double s = p + e;
  return dd{s, e - (s - p)};
}
// x^n for an integer n, |n| <= 64, rounded once from a double-double product chain
__device__ __forceinline__ double pow_int_dd(double x, int n) {
  if (n == 0) return 1.0;
  const int m = n < 0 ? -n : n;
  dd r{x, 0.0};
  for (int i = 1; i < m; ++i) r = dd_mul_d(r, x);
  if (n > 0) return r.h + r.l;
  const double q = 1.0 / r.h;
  double e = fma(-r.h, q, 1.0);
  e = fma(-r.l, q, e);
  return fma(q, e, q);
}
// the reference's std::pow(x, y) for the arguments the mesh physics passes (x >= 0)
__device__ __forceinline__ double pow_like_host(double x, double y) {
  const double yi = rint(y);
  if (yi == y && fabs(y) <= 64.0 && isfinite(x) && (x != 0.0 || y > 0.0)) return pow_int_dd(x, (int)yi);
  return pow(x, y);
}
// x^(1/4), x >= 0
__device__ __forceinline__ double fourth_root(double x) {
  if (!(x > 0.0) || !isfinite(x)) return pow(x, 0.25);
  const double y = sqrt(sqrt(x));
  // one Newton step on y^4 - x with y^4 in double-double
  const double p = y * y, pe = fma(y, y, -p);          // y^2 = p + pe
  const double h = p * p;
  const double l = fma(p, p, -h) + 2.0 * p * pe;       // y^4 = h + l
  const double resid = (h - x) + l;
  return y - resid / (4.0 * y * p);
}

// cell geometry from the per-axis faces: Cell::get_volume (src/cell.h:188-191), get_source_face (:69-76),
// source face area (:83-100, -1.0 without a SOURCE face)
__device__ __forceinline__ void cell_geometry(const MeshDev &m, uint32_t cell, double &vol, double &source_area) {
  const uint32_t sxy = m.nx * m.ny;
  const uint32_t k = cell / sxy, rem = cell - k * sxy, j = rem / m.nx, i = rem - j * m.nx;
  const double *fx = m.faces, *fy = fx + (m.nx + 1), *fz = fy + (m.ny + 1);
  const double dx = fx[i + 1] - fx[i], dy = fy[j + 1] - fy[j], dz = fz[k + 1] - fz[k];
  vol = dx * dy * dz;
  const bool on[6] = {i == 0, i == m.nx - 1, j == 0, j == m.ny - 1, k == 0, k == m.nz - 1};
  int face = -1;
#pragma unroll
  for (int s = 5; s >= 0; --s)
    if (on[s] && m.bc[s] == BC_SOURCE) face = s;
  source_area = -1.0;
  if (face == 0 || face == 1) source_area = dy * dz;
  else if (face == 2 || face == 3) source_area = dx * dz;
  else if (face == 4 || face == 5) source_area = dx * dy;
}

__device__ __forceinline__ void tile_store(double v, double *s_red, double *dst) {
  const double b = block_sum(v, s_red);
  if (threadIdx.x == 0) *dst = b;
}

// src/mesh.h:253-287: opacities, Fleck factor, emission / census / source energies, material energy; tile sums of
// the five running totals.  One thread handles MESH_TILE_ITEMS consecutive cells.
__global__ void __launch_bounds__(MESH_TILE_THREADS) k_mesh_energy(const MeshPhysParams P) {
  __shared__ double s_red[MESH_TILE_THREADS >> 5];
  const uint32_t nc = P.mesh.n_cells;
  const uint32_t base = blockIdx.x * MESH_TILE + threadIdx.x * MESH_TILE_ITEMS;
  double s_mat = 0.0, s_em = 0.0, s_cen = 0.0, s_src = 0.0, s_tot = 0.0;
#pragma unroll 1
  for (int it = 0; it < MESH_TILE_ITEMS; ++it) {
    const uint32_t i = base + it;
    if (i >= nc) break;
    double vol, area;
    cell_geometry(P.mesh, i, vol, area);
    const RegionDev r = P.regions[P.region_of_cell[i]];
    const double T = P.T_e[i], Tr = P.T_r0[i], Ts = P.T_s[i];
    const double opa = r.opac_A + r.opac_B * pow_like_host(T, r.opac_C);  // src/region.h:44-46
    const double ops = r.opac_S;
    const double fleck = 1.0 / (1.0 + P.dt * opa * K_C * (4.0 * K_A * pow_int_dd(T, 3) / (r.cV * r.rho)));
    const double em = P.replicated_factor * P.dt * vol * fleck * opa * K_A * K_C * pow_int_dd(T, 4);
    const double cen = (P.step > 1) ? 0.0 : P.replicated_factor * vol * K_A * pow_int_dd(Tr, 4);
    const double src = P.replicated_factor * 0.25 * K_A * K_C * area * pow_int_dd(Ts, 4) * P.dt;
    P.op_a[i] = opa;
    P.op_s[i] = ops;
    P.f[i] = fleck;
    P.E_emission[i] = em;
    P.E_census[i] = cen;
    P.E_source[i] = src;
    s_mat += T * r.cV * vol * r.rho;
    s_em += em;
    s_cen += cen;
    s_src += src;
    s_tot += src + cen + em;
  }
  const uint32_t nt = P.n_tiles;
  tile_store(s_mat, s_red, &P.tile_sums[MS_PRE_MAT * nt + blockIdx.x]);
  tile_store(s_em, s_red, &P.tile_sums[MS_EMISSION * nt + blockIdx.x]);
  tile_store(s_cen, s_red, &P.tile_sums[MS_CENSUS * nt + blockIdx.x]);
  tile_store(s_src, s_red, &P.tile_sums[MS_SOURCE * nt + blockIdx.x]);
  tile_store(s_tot, s_red, &P.tile_sums[MS_TOTAL * nt + blockIdx.x]);
}

// src/mesh.h:291-315: cells whose share would make no photon are handed whole to rank (i % n_ranks); the totals are
// recomputed.  Also forms what the reference's Allreduce of m_emission_E (:343-345) yields, locally: every rank holds
// the same pre-redistribution share and takes the same decision, so the rank-ordered sum is either the share added
// n_ranks times or the one un-split value (see host/mesh.h).
//
// No collective is needed for any of it.  The global source energy the rule divides by (:291-294) is the rank-ordered
// sum of n_ranks identical totals and is read from `sums_in` (k_mesh_final_sums wrote this rank's totals there); and
// because every rank takes the same decisions from the same values, this rank forms the post-redistribution totals of
// EVERY rank (tile_sums[r][q][tile]) -- what the reference gets from its second MPI_Allreduce of total_photon_E
// (src/replicated_driver.h:56-59) is then a local rank-ordered sum (k_mesh_final_sums).  Only this rank's energies are
// written back.
__global__ void __launch_bounds__(MESH_TILE_THREADS) k_mesh_redistribute(const MeshPhysParams P,
                                                                         const double *__restrict__ sums_in) {
  __shared__ double s_red[MESH_TILE_THREADS >> 5];
  const uint32_t nc = P.mesh.n_cells;
  const uint32_t base = blockIdx.x * MESH_TILE + threadIdx.x * MESH_TILE_ITEMS;
  // src/mesh.h:291-294: global_source_E = Allreduce(tot_emission + tot_census + tot_source), summed in rank order
  double gse;
  if (sums_in) {
    const double t = sums_in[MS_EMISSION] + sums_in[MS_CENSUS] + sums_in[MS_SOURCE];
    gse = t;
    for (int r = 1; r < P.n_ranks; ++r) gse += t;
  } else {
    gse = P.global_source_E;
  }
  double em0[MESH_TILE_ITEMS], cen0[MESH_TILE_ITEMS], src0[MESH_TILE_ITEMS];
  bool em_small[MESH_TILE_ITEMS], cen_small[MESH_TILE_ITEMS], src_small[MESH_TILE_ITEMS];
#pragma unroll
  for (int it = 0; it < MESH_TILE_ITEMS; ++it) {
    const uint32_t i = base + it;
    em0[it] = cen0[it] = src0[it] = 0.0;
    em_small[it] = cen_small[it] = src_small[it] = false;
    if (i >= nc) continue;
    const double em = P.E_emission[i], cen = P.E_census[i], src = P.E_source[i];
    em0[it] = em; cen0[it] = cen; src0[it] = src;
    em_small[it] = em > 0.0 && int(P.n_user * (em / gse)) == 0;
    cen_small[it] = P.step == 1 && cen > 0.0 && int(P.n_user * (cen / gse)) == 0;
    src_small[it] = src > 0.0 && int(P.n_user * (src / gse)) == 0;
    double g;
    if (em_small[it]) {
      g = em / P.replicated_factor;
      for (int r = 1; r < P.n_ranks; ++r) g = g + 0.0;
    } else {
      g = em;
      for (int r = 1; r < P.n_ranks; ++r) g = g + em;
    }
    P.E_emission_global[i] = g;
  }
  const uint32_t nt = P.n_tiles;
#pragma unroll 1
  for (int r = 0; r < P.n_ranks; ++r) {
    double s_em = 0.0, s_cen = 0.0, s_src = 0.0, s_tot = 0.0;
#pragma unroll
    for (int it = 0; it < MESH_TILE_ITEMS; ++it) {
      const uint32_t i = base + it;
      if (i >= nc) continue;
      const bool owner = (int)(i % (uint32_t)P.n_ranks) == r;
      const double em = em_small[it] ? (owner ? em0[it] / P.replicated_factor : 0.0) : em0[it];
      const double cen = cen_small[it] ? (owner ? cen0[it] / P.replicated_factor : 0.0) : cen0[it];
      const double src = src_small[it] ? (owner ? src0[it] / P.replicated_factor : 0.0) : src0[it];
      if (r == P.rank) {
        P.E_emission[i] = em;
        P.E_census[i] = cen;
        P.E_source[i] = src;
      }
      s_em += em;
      s_cen += cen;
      s_src += src;
      s_tot += src + cen + em;
    }
    double *ts = P.tile_sums + (uint64_t)r * MS_N * nt;
    tile_store(s_em, s_red, &ts[MS_EMISSION * nt + blockIdx.x]);
    tile_store(s_cen, s_red, &ts[MS_CENSUS * nt + blockIdx.x]);
    tile_store(s_src, s_red, &ts[MS_SOURCE * nt + blockIdx.x]);
    tile_store(s_tot, s_red, &ts[MS_TOTAL * nt + blockIdx.x]);
  }
}

// src/mesh.h:343-362: new material temperature from the (rank-summed) tallies, radiation temperature diagnostic,
// absorbed and post-cycle material energy.  `emission` is E_emission_global in a multi-rank run.
__global__ void __launch_bounds__(MESH_TILE_THREADS) k_mesh_update_temperature(const MeshPhysParams P,
                                                                               const double *__restrict__ emission) {
  __shared__ double s_red[MESH_TILE_THREADS >> 5];
  const uint32_t nc = P.mesh.n_cells;
  const uint32_t base = blockIdx.x * MESH_TILE + threadIdx.x * MESH_TILE_ITEMS;
  double s_abs = 0.0, s_mat = 0.0;
#pragma unroll 1
  for (int it = 0; it < MESH_TILE_ITEMS; ++it) {
    const uint32_t i = base + it;
    if (i >= nc) break;
    double vol, area;
    cell_geometry(P.mesh, i, vol, area);
    const RegionDev r = P.regions[P.region_of_cell[i]];
    const double2 t = P.tally[i];
    const double T = P.T_e[i];
    const double T_new = T + (t.x - emission[i]) / (r.cV * vol * r.rho);
    P.T_r[i] = fourth_root(t.y / (vol * P.dt * K_A * K_C));
    P.T_e[i] = T_new;
    s_abs += t.x;
    s_mat += T_new * r.cV * vol * r.rho;
  }
  const uint32_t nt = P.n_tiles;
  tile_store(s_abs, s_red, &P.tile_sums[MS_ABS * nt + blockIdx.x]);
  tile_store(s_mat, s_red, &P.tile_sums[MS_POST_MAT * nt + blockIdx.x]);
}

// out[r][q] = sum_t tile_sums[r][q][t] in a fixed order: blockIdx.x = r (the kernels that are not per-rank use block 0 =
// the rank's own block); for every requested quantity thread t adds its contiguous run of tiles serially and the 256 runs
// are added by a fixed tree -- reproducible, identical on every rank, a few microseconds for the 7 813 tiles of the 200^3
// cube where one thread per quantity took 35
__global__ void __launch_bounds__(256) k_mesh_final_sums(const double *__restrict__ tile_sums, uint32_t n_tiles,
                                                         uint32_t q_mask, double *__restrict__ out) {
  __shared__ double s[256];
  const uint32_t per = (n_tiles + 255u) / 256u;
  const uint32_t b = threadIdx.x * per, e = (b + per < n_tiles) ? b + per : n_tiles;
  for (uint32_t q = 0; q < (uint32_t)MS_N; ++q) {
    if (!((q_mask >> q) & 1u)) continue;  // (uniform over the block)
    const double *v = tile_sums + ((uint64_t)blockIdx.x * MS_N + q) * n_tiles;
    double acc = 0.0;
    for (uint32_t i = b; i < e; ++i) acc += v[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t w = 128; w > 0; w >>= 1) {
      if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x * MS_N + q] = s[0];
    __syncthreads();
  }
}

}  // namespace bg
