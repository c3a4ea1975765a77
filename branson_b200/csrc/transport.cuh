// transport.cuh -- the per-photon history loop as a persistent sm_100a kernel.
//
// Physics and decision order follow the reference's transport_photon (src/history_based_transport.h:32-141):
//   sigma lookup :56-59, distance to scatter :62-63 (a draw is consumed only if total_sigma_s > 0), distance to
//   boundary src/cell.h:116-132, distance to census :68, min :70, implicit capture :74-82 (+ Photon::move
//   src/photon.h:109-114), energy cutoff FIRST :85-91, then equality dispatch scatter :94-101 (isotropic angle
//   src/sampling_functions.h:57-70, group resample :126-138), boundary :103-131, census :133-138.
//
// What is B200-specific (none of this exists in the reference's gpu_no_accel_transport, :263-275):
//   * persistent grid (blocks = SMs x resident CTAs); a warp pulls chunks of photons from a global counter and a
//     lane that retires its photon immediately refills from the warp's chunk, so the 10x spread in history length
//     (mean 150-190, max > 2000 events) does not idle the warp;
//   * photon state lives in registers for the whole history, read and written as six coalesced 128-bit streams;
//   * geometry comes from per-axis face arrays in shared memory (cell = i + nx*(j + ny*k)), so a cell crossing
//     costs zero HBM bytes for nodes / neighbours / boundary conditions;
//   * tallies are return-less FP64 reductions (RED.E.ADD.F64) on interleaved {abs_E, track_E}, or -- in the
//     deterministic validation mode -- a per-photon deposit log that is later summed in the reference's serial order.
#pragma once
#include "common.cuh"
#include "rng.cuh"

namespace bg {

enum : int { TM_ATOMIC = 0, TM_COUNT = 1, TM_LOG = 2 };

struct TransportParams {
  PhotonSoA ph;
  uint64_t n;
  uint8_t *desc;
  uint32_t *counters;  // [n][4] or nullptr
  MeshDev mesh;
  const double *f;    // [n_cells]
  const double *opa;  // [n_cells][G]
  const double *ops;  // [n_cells][G]
  double2 *tally;     // [n_cells] {abs_E, track_E}
  uint64_t ctr_hi;    // seed << 32
  unsigned long long *work_counter;
  uint32_t chunk;
  int writeback_all;  // 1: every photon's full state is written back; 0: full state only for CENSUS photons
  // deterministic tally mode
  uint32_t *ndep;           // TM_COUNT: deposits per photon
  const uint64_t *dep_off;  // TM_LOG: exclusive scan of ndep
  uint32_t *dep_cell;
  double2 *dep_val;
  unsigned long long *stats;
};

struct Lane {
  double x, y, z, ax, ay, az, E, E0, life;
  uint64_t ctr, stream, idx;
  uint32_t cell, group;
  int i, j, k;
};

template <bool SMEM>
__device__ __forceinline__ double face_at(const double *faces, uint32_t idx) {
  return faces[idx];
}

// One warp-uniform step of work distribution: hand photon indices to lanes that want one.
// Returns the index for this lane or ~0ull.
struct WarpQueue {
  uint64_t next, end;
  bool exhausted;
};

template <int MODE, bool COUNTERS, bool SMEM>
__global__ void __launch_bounds__(128, 4) k_transport_history(const TransportParams P) {
  extern __shared__ double s_faces[];
  const double *faces;
  if (SMEM) {
    for (uint32_t t = threadIdx.x; t < P.mesh.n_faces; t += blockDim.x) s_faces[t] = P.mesh.faces[t];
    __syncthreads();
    faces = s_faces;
  } else {
    faces = P.mesh.faces;
  }
  const double *fx = faces;
  const double *fy = faces + (P.mesh.nx + 1);
  const double *fz = fy + (P.mesh.ny + 1);
  const uint32_t nx = P.mesh.nx, ny = P.mesh.ny, nz = P.mesh.nz, G = P.mesh.G;
  const uint32_t sxy = nx * ny;
  const unsigned FULL = 0xffffffffu;
  const unsigned lane_id = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane_id) - 1u;
  const uint64_t ctr_hi = P.ctr_hi;

  WarpQueue q{0, 0, false};
  bool active = false;
  Lane L;
  L.x = L.y = L.z = L.ax = L.ay = L.az = L.E = L.E0 = L.life = 0.0;
  L.ctr = L.stream = L.idx = 0;
  L.cell = L.group = 0;
  L.i = L.j = L.k = 0;
  double f = 0.0, sig_a = 0.0, sig_s = 0.0;
  double loc_abs = 0.0, loc_trk = 0.0;
  uint32_t surface = 0;  // persists across events like the reference's surface_cross (:37)
  uint32_t c_ev = 0, c_sc = 0, c_cr = 0, c_rf = 0;  // per-photon counters
  uint32_t t_ev = 0, t_sc = 0, t_cr = 0, t_rf = 0, t_dep = 0, t_lk = 0;  // per-thread totals
  uint64_t gmask = 0;  // groups touched during the current cell visit (algorithmic-bytes accounting)
  uint32_t lk_visit = 0;
  uint32_t ndep = 0;
  uint64_t dep_pos = 0;
  bool need_f = false, need_xs = false;

  auto close_visit = [&]() {
    // distinct (cell, group) opacity pairs touched in this visit, capped at G (S_cell)
    t_lk += (G <= 64) ? (uint32_t)__popcll(gmask) : min(lk_visit, G);
    gmask = 0;
    lk_visit = 0;
  };

  auto deposit = [&](uint32_t cell, double a, double t) {
    ++t_dep;
    if (MODE == TM_ATOMIC) {
      atomicAdd(&P.tally[cell].x, a);
      atomicAdd(&P.tally[cell].y, t);
    } else if (MODE == TM_COUNT) {
      ++ndep;
    } else {
      P.dep_cell[dep_pos] = cell;
      P.dep_val[dep_pos] = make_double2(a, t);
      ++dep_pos;
    }
  };

  for (;;) {
    // ---------------- refill idle lanes ----------------
    unsigned need = __ballot_sync(FULL, !active);
    if (need) {
      bool want = !active;
      while (!q.exhausted) {
        if (q.next == q.end) {
          unsigned long long base = 0;
          if (lane_id == 0) base = atomicAdd(P.work_counter, (unsigned long long)P.chunk);
          base = __shfl_sync(FULL, base, 0);
          if (base >= P.n) {
            q.exhausted = true;
            break;
          }
          q.next = base;
          q.end = (base + P.chunk < P.n) ? base + P.chunk : P.n;
        }
        const unsigned m = __ballot_sync(FULL, want);
        if (!m) break;
        const uint64_t avail = q.end - q.next;
        const unsigned r = __popc(m & lt_mask);
        if (want && r < avail) {
          const uint64_t idx = q.next + r;
          const double2 xy = P.ph.xy[idx], za = P.ph.za[idx], bc = P.ph.bc[idx], ee = P.ph.ee[idx];
          const ulonglong2 lc = P.ph.lc[idx], sg = P.ph.sg[idx];
          L.x = xy.x; L.y = xy.y; L.z = za.x; L.ax = za.y; L.ay = bc.x; L.az = bc.y;
          L.E = ee.x; L.E0 = ee.y;
          L.life = __longlong_as_double((long long)lc.x);
          L.ctr = lc.y;
          L.stream = sg.x;
          L.cell = (uint32_t)sg.y;
          L.group = (uint32_t)(sg.y >> 32);
          L.idx = idx;
          const uint32_t kk = L.cell / sxy;
          const uint32_t rem = L.cell - kk * sxy;
          const uint32_t jj = rem / nx;
          L.k = (int)kk; L.j = (int)jj; L.i = (int)(rem - jj * nx);
          loc_abs = 0.0; loc_trk = 0.0;
          c_ev = c_sc = c_cr = c_rf = 0;
          ndep = 0;
          if (MODE == TM_LOG) dep_pos = P.dep_off[idx];
          need_f = true; need_xs = true;
          gmask = 0; lk_visit = 0;
          active = true;
          want = false;
        }
        const uint64_t asked = (uint64_t)__popc(m);
        q.next += (asked < avail) ? asked : avail;
        if (asked <= avail) break;
      }
      if (!__any_sync(FULL, active)) break;
    }
    if (!active) continue;

    // ---------------- one event (one trip of the reference's while(active) loop) ----------------
    if (need_f) { f = __ldg(&P.f[L.cell]); need_f = false; }
    if (need_xs) {
      const uint64_t o = (uint64_t)L.cell * G + L.group;
      sig_a = __ldg(&P.opa[o]);
      sig_s = __ldg(&P.ops[o]);
      need_xs = false;
      if (G <= 64) gmask |= 1ull << L.group;
      ++lk_visit;
    }
    const double total_sigma_s = (1.0 - f) * sig_a + sig_s;
    double d_scat = 1.0e100;
    if (total_sigma_s > 0.0) d_scat = -log(rng_next(L.ctr, ctr_hi, L.stream)) / total_sigma_s;

    // distance to boundary: strict-minimum scan over x, y, z starting from 1e16 (src/cell.h:116-132)
    double d_bnd = 1.0e16;
    {
      const bool px = 0.0 < L.ax, py = 0.0 < L.ay, pz = 0.0 < L.az;
      const double dx = (fx[L.i + (px ? 1 : 0)] - L.x) / L.ax;
      const double dy = (fy[L.j + (py ? 1 : 0)] - L.y) / L.ay;
      const double dz = (fz[L.k + (pz ? 1 : 0)] - L.z) / L.az;
      if (dx < d_bnd) { d_bnd = dx; surface = px ? 1u : 0u; }
      if (dy < d_bnd) { d_bnd = dy; surface = 2u + (py ? 1u : 0u); }
      if (dz < d_bnd) { d_bnd = dz; surface = 4u + (pz ? 1u : 0u); }
    }
    const double d_cen = L.life;
    const double m1 = (d_cen < d_bnd) ? d_cen : d_bnd;    // std::min(boundary, census)
    const double d = (m1 < d_scat) ? m1 : d_scat;         // std::min(scatter, m1)

    const double absorbed = L.E * (1.0 - exp(-sig_a * f * d));
    loc_abs += absorbed;
    loc_trk += absorbed / (sig_a * f);
    L.E = L.E - absorbed;
    L.x += L.ax * d;
    L.y += L.ay * d;
    L.z += L.az * d;
    L.life -= d;
    ++c_ev;

    uint8_t descriptor = EV_PASS;
    bool done = false;
    if (L.E / L.E0 < K_CUTOFF) {
      loc_abs += L.E;
      deposit(L.cell, loc_abs, loc_trk);
      descriptor = EV_KILLED;
      done = true;
    } else if (d == d_scat) {
      // isotropic re-emission direction (src/sampling_functions.h:57-70)
      const double mu = rng_next(L.ctr, ctr_hi, L.stream) * 2.0 - 1.0;
      const double phi = rng_next(L.ctr, ctr_hi, L.stream) * 2.0 * K_PI;
      const double sin_theta = sqrt(1.0 - mu * mu);
      double sp, cp;
      sincos(phi, &sp, &cp);
      L.ax = sin_theta * cp;
      L.ay = sin_theta * sp;
      L.az = mu;
      // physical vs effective scatter (src/history_based_transport.h:98-100)
      if (rng_next(L.ctr, ctr_hi, L.stream) > (sig_s / ((1.0 - f) * sig_a + sig_s))) {
        // sample_emission_group (src/sampling_functions.h:126-138): sequential walk of the cell's group array
        double cdf = rng_next(L.ctr, ctr_hi, L.stream);
        const double *ag = P.opa + (uint64_t)L.cell * G;
        const double norm = 1.0 / (__ldg(&ag[0]) * (double)G);
        int g = -1;
        while (cdf > 0.0) {
          ++g;
          if (g >= (int)G) { g = (int)G - 1; break; }  // round-off guard: the reference would read past the array
          cdf -= __ldg(&ag[g]) * norm;
        }
        if ((uint32_t)g != L.group) { L.group = (uint32_t)g; need_xs = true; }
      }
      ++c_sc;
    } else if (d == d_bnd) {
      const uint32_t axis = surface >> 1;
      const bool pos_dir = surface & 1u;
      bool domain_face;
      if (axis == 0) domain_face = pos_dir ? (L.i == (int)nx - 1) : (L.i == 0);
      else if (axis == 1) domain_face = pos_dir ? (L.j == (int)ny - 1) : (L.j == 0);
      else domain_face = pos_dir ? (L.k == (int)nz - 1) : (L.k == 0);
      const int bcv = domain_face ? P.mesh.bc[surface] : BC_ELEMENT;
      if (bcv == BC_ELEMENT) {
        deposit(L.cell, loc_abs, loc_trk);
        close_visit();
        const int step = pos_dir ? 1 : -1;
        if (axis == 0) { L.i += step; L.cell += (uint32_t)step; }
        else if (axis == 1) { L.j += step; L.cell += (uint32_t)(step * (int)nx); }
        else { L.k += step; L.cell += (uint32_t)(step * (int)sxy); }
        loc_abs = 0.0;
        loc_trk = 0.0;
        need_f = true;
        need_xs = true;
        ++c_cr;
      } else if (bcv == BC_VACUUM || bcv == BC_SOURCE) {
        deposit(L.cell, loc_abs, loc_trk);
        descriptor = EV_EXIT;
        done = true;
      } else if (bcv == BC_PROCESSOR) {
        // never produced in replicated mode (every rank owns the whole mesh); kept for fidelity (:115-121)
        deposit(L.cell, loc_abs, loc_trk);
        descriptor = EV_PASS;
        done = true;
      } else {  // REFLECT (:127-130)
        if (axis == 0) L.ax = -L.ax;
        else if (axis == 1) L.ay = -L.ay;
        else L.az = -L.az;
        ++c_rf;
      }
    } else if (d == d_cen) {
      deposit(L.cell, loc_abs, loc_trk);
      descriptor = EV_CENSUS;
      done = true;
    }

    if (done) {
      close_visit();
      t_ev += c_ev; t_sc += c_sc; t_cr += c_cr; t_rf += c_rf;
      const uint64_t idx = L.idx;
      if (MODE == TM_COUNT) {
        P.ndep[idx] = ndep;
      } else {
        P.desc[idx] = descriptor;
        P.ph.ee[idx] = make_double2(L.E, L.E0);
        if (P.writeback_all || descriptor == EV_CENSUS) {
          P.ph.xy[idx] = make_double2(L.x, L.y);
          P.ph.za[idx] = make_double2(L.z, L.ax);
          P.ph.bc[idx] = make_double2(L.ay, L.az);
          P.ph.lc[idx] = make_ulonglong2((unsigned long long)__double_as_longlong(L.life), L.ctr);
          P.ph.sg[idx] = make_ulonglong2(L.stream, (unsigned long long)L.cell | ((unsigned long long)L.group << 32));
        }
        if (COUNTERS) reinterpret_cast<uint4 *>(P.counters)[idx] = make_uint4(c_ev, c_sc, c_cr, c_rf);
      }
      active = false;
    }
  }

  // ---------------- statistics: warp reduce, one atomic per warp and counter ----------------
  if (MODE != TM_COUNT) {
    unsigned long long v[6] = {t_ev, t_sc, t_cr, t_rf, t_dep, t_lk};
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      unsigned long long x = v[s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
      if (lane_id == 0 && x) atomicAdd(&P.stats[s], x);
    }
  }
}

}  // namespace bg
