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
#include "fastmath.cuh"
#include "rng.cuh"

// A/B switches for tools/ timing sweeps (make variant NAME=.. EXTRA=-DBG_FM_LOG=0); the shipped build has all on
#ifndef BG_LAZY_GROUP
#define BG_LAZY_GROUP 1  // A/B knob: 0 = sample the group at every effective scatter (see scatter_event)
#endif
#ifndef BG_FM_LOG
#define BG_FM_LOG 1
#endif
#ifndef BG_FM_EXP
#define BG_FM_EXP 1
#endif
#ifndef BG_FM_SINCOS
#define BG_FM_SINCOS 1
#endif
#ifndef BG_FM_DIV
#define BG_FM_DIV 1
#endif
#if BG_FM_DIV
#define BG_DIV(a, b) fm_div((a), (b))
#else
#define BG_DIV(a, b) ((a) / (b))
#endif

namespace bg {

enum : int { TM_ATOMIC = 0, TM_COUNT = 1, TM_LOG = 2 };

struct TransportParams {
  PhotonSoA ph;
  uint64_t n;          // < 2^32 (checked by the launcher): photon indices travel as 32-bit values
  uint8_t *desc;
  uint32_t *counters;  // [n][4] or nullptr
  MeshDev mesh;
  const double *f;    // [n_cells]
  const double *opa;  // [n_cells][G]
  const double *ops;  // [n_cells][G]
  const double2 *cellrec;  // [n_cells][2]: {f, sigma_a}, {sigma_s, 0} -- valid where uniform_groups
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
  // RESUME launches (the tail of the event-based variant, event.cuh)
  double inv_sxy, inv_nx;   // 1 / (nx ny), 1 / nx: cell -> (i, j, k) without integer division (set by make_params)
  const uint32_t *index_list;
  const double2 *carry_acc;
  const uint4 *carry_cnt;
  const uint32_t *carry_lk;
  int uniform_groups;  // every cell's abs_groups are all equal (gray decks, Cell::set_op_a): closed-form group walk
  int resume_pending_scatter;  // the photons of index_list are parked AT a scatter (else after a non-scatter event)
  // divergence / contention control of the history kernel (bgpu_set_tuning; results do not depend on either)
  uint32_t scatter_batch;  // a warp samples its parked scatters once this many lanes wait at one (1 = immediately)
  int aggregate;           // combine the same-cell deposits of a warp trip into one pair of atomics
  // Replicated tallies: warp w deposits into copy w % tally_copies (copy 0 = `tally`, copies 1.. = `tally_rep`, n_cells
  // entries each, zero at launch); k_fold_tally adds the copies into `tally` afterwards.  Same-address FP64 reductions
  // serialise in L2, so a few hot cells (hot_zone's corner, marshak's 25 cells) otherwise throttle the whole grid.
  double2 *tally_rep;
  uint32_t tally_copies;
};

// Photon state of one lane, kept in registers for the whole history.
struct PState {
  double x, y, z, ax, ay, az, E, E0, life;
  uint64_t ctr, stream;
  uint32_t cell, group;
  int i, j, k;
  double f, sig_a, sig_s;        // cell / group data of the current visit (loaded when the cell or the group changes)
  double loc_abs, loc_trk;       // thread-local tallies of the current cell visit (reference :45-46)
  double p_grp;                  // abs_groups[g] * norm of the current cell when all its groups are equal; 0 = not yet
  uint32_t surface;              // persists across events like the reference's surface_cross (:37)
  uint32_t c_sc, c_cr, c_rf, c_lk;  // per-photon counters: scatters, crossings, reflections, lookups
  uint32_t ev_entry;             // c_sc + c_cr + c_rf when the current cell visit began (algorithmic-bytes accounting)
  // lazily sampled group (scatter_event): cell of the history's latest effective scatter (~0u: none pending) and the low
  // 32 bits of the counter of its group-CDF draw
  uint32_t grp_cell, grp_ctr32;
};

struct PCtx {
  const double *fx, *fy, *fz;    // per-axis faces (shared memory when they fit)
  uint32_t nx, ny, nz, G, sxy;
  const double *f, *opa, *ops;
  const double2 *cellrec;        // packed per-cell records (k_fill_cellrec), read by the PACKED instantiations
  uint64_t ctr_hi;
  bool uniform_groups;
  double inv_sxy, inv_nx;
};

// n / d for n < 2^32 through the double reciprocal inv = fl(1 / d): the product is within 2^-20 / d of the true
// quotient, closer than any non-integer quotient is to an integer, so truncation can only be off (by -1) when d divides
// n exactly -- one fix-up.  (The SASS of an integer division is ~20 instructions at 1-2 active lanes per refill.)
__device__ __forceinline__ uint32_t div_by_inv(uint32_t n, uint32_t d, double inv, uint32_t &rem) {
  uint32_t q = __double2uint_rz(__uint2double_rn(n) * inv);
  rem = n - q * d;
  if (rem >= d) { ++q; rem -= d; }
  return q;
}

enum : int { R_CONTINUE = 0, R_DONE = 1, R_SCATTER = 2 };

// Every trip of the reference's loop ends in exactly one of: scatter, cell crossing, reflection, or the event that
// retires the photon -- so the trip count of a finished history needs no counter of its own.
__device__ __forceinline__ uint32_t events_of_finished(const PState &S) { return S.c_sc + S.c_cr + S.c_rf + 1u; }

// (sigma_a, sigma_s) of the current (cell, group): src/history_based_transport.h:56-57
__device__ __forceinline__ void load_xs(PState &S, const PCtx &C) {
  const uint64_t o = (uint64_t)S.cell * C.G + S.group;
  S.sig_a = __ldg(&C.opa[o]);
  S.sig_s = __ldg(&C.ops[o]);
}

// entering a cell: Fleck factor (:58) and the opacities of the photon's group
// PACKED (history kernel on decks whose cells carry the same opacities in every group, P.uniform_groups): the visit's
// three values come from the cell's 32-byte record {f, sigma_a, sigma_s, 0} (k_fill_cellrec) -- one address, one sector,
// and a table that stays in L2 (19 MB for the hohlraum's 591 500 cells; the [cell][G] arrays are 289 MB, and a photon
// entering a cell found their three sectors evicted: 9.2 GB of DRAM reads per hohlraum launch)
template <bool PACKED = false>
__device__ __forceinline__ void enter_cell(PState &S, const PCtx &C) {
  S.p_grp = 0.0;
  if (PACKED) {
    const double2 *r = C.cellrec + 2 * (uint64_t)S.cell;
    const double2 fa = __ldg(r);
    S.sig_s = __ldg(reinterpret_cast<const double *>(r + 1));
    S.f = fa.x; S.sig_a = fa.y;
  } else {
    S.f = __ldg(&C.f[S.cell]);
    load_xs(S, C);
  }
}

template <bool PACKED = false>
__device__ __forceinline__ void pstate_load(PState &S, const PhotonSoA &ph, uint64_t idx, const PCtx &C) {
  const double2 xy = ph.xy[idx], za = ph.za[idx], bc = ph.bc[idx], ee = ph.ee[idx];
  const ulonglong2 lc = ph.lc[idx], sg = ph.sg[idx];
  S.x = xy.x; S.y = xy.y; S.z = za.x; S.ax = za.y; S.ay = bc.x; S.az = bc.y;
  S.E = ee.x; S.E0 = ee.y;
  S.life = __longlong_as_double((long long)lc.x);
  S.ctr = lc.y;
  S.stream = sg.x;
  S.cell = (uint32_t)sg.y;
  S.group = (uint32_t)(sg.y >> 32);
  uint32_t rem, ii;
  const uint32_t kk = div_by_inv(S.cell, C.sxy, C.inv_sxy, rem);
  const uint32_t jj = div_by_inv(rem, C.nx, C.inv_nx, ii);
  S.k = (int)kk; S.j = (int)jj; S.i = (int)ii;
  S.loc_abs = 0.0; S.loc_trk = 0.0;
  S.c_sc = S.c_cr = S.c_rf = S.c_lk = 0;
  S.ev_entry = 0;
  S.grp_cell = ~0u;
  S.grp_ctr32 = 0u;
  enter_cell<PACKED>(S, C);
}

__device__ __forceinline__ void pstate_store_full(const PState &S, const PhotonSoA &ph, uint64_t idx) {
  ph.xy[idx] = make_double2(S.x, S.y);
  ph.za[idx] = make_double2(S.z, S.ax);
  ph.bc[idx] = make_double2(S.ay, S.az);
  ph.lc[idx] = make_ulonglong2((unsigned long long)__double_as_longlong(S.life), S.ctr);
  ph.sg[idx] = make_ulonglong2(S.stream, (unsigned long long)S.cell | ((unsigned long long)S.group << 32));
}

__device__ __forceinline__ void close_visit(PState &S, const PCtx &C, const uint32_t closing_event_uncounted) {
  // (sigma_a, sigma_s) pairs the reference algorithm fetches during this visit, SURVEY section 8d: S_cell counts one pair
  // per event of the visit (every trip of the reference's loop re-reads its group's pair, :56-57), at most the G distinct
  // pairs the cell has.  Counted from the event counters, not from what this kernel happens to load (with lazily
  // sampled groups it loads one pair per visit), so that the algorithmic bytes are a property of the workload.
  const uint32_t ev_now = S.c_sc + S.c_cr + S.c_rf;
  const uint32_t n = ev_now - S.ev_entry + closing_event_uncounted;
  S.c_lk += (n < C.G) ? n : C.G;
  S.ev_entry = ev_now;
}

// The six domain boundary conditions packed three bits each (bc_type values 0..4), indexed by face: a register
// instead of a divergent constant-bank lookup.
__host__ __device__ __forceinline__ uint32_t pack_bc(const int *bc) {
  uint32_t p = 0;
  for (int s = 0; s < 6; ++s) p |= ((uint32_t)bc[s] & 7u) << (3 * s);
  return p;
}

// One trip of the reference's while(active) loop (src/history_based_transport.h:55-139) up to the event dispatch.
// Returns R_SCATTER with the scatter not yet sampled (the caller runs scatter_event), R_DONE with `descriptor` set, or
// R_CONTINUE after a cell crossing / reflection.  `bcpack` = pack_bc(domain boundary conditions).
//
// Divergence control: the boundary branch is straight-line code (axis / direction / boundary type resolved with
// selects), so lanes crossing different faces run it together -- ncu showed the branchy form executing the crossing
// path six times per trip at 4 of 32 lanes -- and every deposit of the trip (cell left, photon killed / escaped /
// reaching census) is issued from ONE converged site: deposit(do_it, cell, abs, trk, lanes), called by every lane of
// `lanes` (the lanes that entered this function together), which lets the caller combine same-cell deposits.
template <bool PACKED = false, class Deposit>
__device__ __forceinline__ int advance_event(PState &S, const PCtx &C, const uint32_t bcpack, Deposit &&deposit,
                                             uint8_t &descriptor, const unsigned lanes) {
  const double total_sigma_s = (1.0 - S.f) * S.sig_a + S.sig_s;
  // Distance to the next collision (:62-63).  The reference draws only if total_sigma_s > 0; here the Threefry block is
  // evaluated unconditionally and the draw merely *consumed* under that condition (the counter decides what the stream
  // yields next, not the evaluation), which keeps this long dependent integer chain in one basic block with the
  // boundary distances below, so the FP64 divisions fill its latency.  (total_sigma_s > 0 in every reference deck.)
  double d_scat;
  {
    const uint64_t w = threefry2x64_20_w0(S.ctr, C.ctr_hi, S.stream);
    const bool collide = total_sigma_s > 0.0;
    S.ctr += collide ? 1u : 0u;
#if BG_FM_LOG
    // -log(u) / sigma with u = ((w >> 11) | 1) 2^-53 (rng.cuh u01_from_bits): the 2^-53 goes into the exponent
    const double dd = BG_DIV(-fm_log_pos_scaled(__ull2double_rn((w >> 11) | 1ULL), -53), total_sigma_s);
#else
    const double dd = -log(u01_from_bits(w)) / total_sigma_s;
#endif
    d_scat = collide ? dd : 1.0e100;
  }

  // distance to boundary: strict-minimum scan over x, y, z starting from 1e16 (src/cell.h:116-132)
  double d_bnd = 1.0e16;
  {
    const bool px = 0.0 < S.ax, py = 0.0 < S.ay, pz = 0.0 < S.az;
    const double dx = BG_DIV(C.fx[S.i + (px ? 1 : 0)] - S.x, S.ax);
    const double dy = BG_DIV(C.fy[S.j + (py ? 1 : 0)] - S.y, S.ay);
    const double dz = BG_DIV(C.fz[S.k + (pz ? 1 : 0)] - S.z, S.az);
    if (dx < d_bnd) { d_bnd = dx; S.surface = px ? 1u : 0u; }
    if (dy < d_bnd) { d_bnd = dy; S.surface = 2u + (py ? 1u : 0u); }
    if (dz < d_bnd) { d_bnd = dz; S.surface = 4u + (pz ? 1u : 0u); }
  }
  const double d_cen = S.life;
  const double m1 = (d_cen < d_bnd) ? d_cen : d_bnd;  // std::min(boundary, census)
  const double d = (m1 < d_scat) ? m1 : d_scat;       // std::min(scatter, m1)

#if BG_FM_EXP
  const double absorbed = S.E * (1.0 - fm_exp_flush(-S.sig_a * S.f * d));
#else
  const double absorbed = S.E * (1.0 - exp(-S.sig_a * S.f * d));
#endif
  S.loc_abs += absorbed;
  S.loc_trk += BG_DIV(absorbed, S.sig_a * S.f);
  S.E = S.E - absorbed;
  S.x += S.ax * d;
  S.y += S.ay * d;
  S.z += S.az * d;
  S.life -= d;

  int r = R_CONTINUE;
  bool dep = false, crossed = false;
  const uint32_t dep_cell = S.cell;
  if (BG_DIV(S.E, S.E0) < K_CUTOFF) {  // energy cutoff first (:85-91)
    S.loc_abs += S.E;
    dep = true;
    descriptor = EV_KILLED;
    r = R_DONE;
  } else if (d == d_scat) {  // (:94-101), sampled by the caller
    r = R_SCATTER;
  } else if (d == d_bnd) {  // (:103-131)
    const uint32_t axis = S.surface >> 1;
    const bool pos_dir = (S.surface & 1u) != 0u;
    const int pos = (axis == 0u) ? S.i : ((axis == 1u) ? S.j : S.k);
    const int lim = (int)((axis == 0u) ? C.nx : ((axis == 1u) ? C.ny : C.nz)) - 1;
    const bool domain_face = pos == (pos_dir ? lim : 0);
    const uint32_t bcv = domain_face ? ((bcpack >> (3u * S.surface)) & 7u) : (uint32_t)BC_ELEMENT;
    const bool cross = bcv == (uint32_t)BC_ELEMENT;   // (:104-114)
    const bool reflect = bcv == (uint32_t)BC_REFLECT;  // (:127-130)
    // ELEMENT: step into the neighbour (cell = i + nx (j + ny k)); every other boundary type leaves i, j, k alone
    const int step = cross ? (pos_dir ? 1 : -1) : 0;
    S.i += (axis == 0u) ? step : 0;
    S.j += (axis == 1u) ? step : 0;
    S.k += (axis == 2u) ? step : 0;
    S.cell += (uint32_t)(step * (int)((axis == 0u) ? 1u : ((axis == 1u) ? C.nx : C.sxy)));
    S.c_cr += cross ? 1u : 0u;
    // REFLECT: flip the direction component normal to the face
    S.ax = (reflect && axis == 0u) ? -S.ax : S.ax;
    S.ay = (reflect && axis == 1u) ? -S.ay : S.ay;
    S.az = (reflect && axis == 2u) ? -S.az : S.az;
    S.c_rf += reflect ? 1u : 0u;
    // VACUUM or SOURCE: the photon escapes (:122-126); PROCESSOR (never produced in replicated mode, every rank owns
    // the whole mesh; kept for fidelity, :115-121): it would be passed on
    const bool leave = !cross && !reflect;
    descriptor = leave ? ((bcv == (uint32_t)BC_PROCESSOR) ? EV_PASS : EV_EXIT) : descriptor;
    r = leave ? R_DONE : R_CONTINUE;
    dep = !reflect;
    crossed = cross;
  } else if (d == d_cen) {  // (:133-138)
    dep = true;
    descriptor = EV_CENSUS;
    r = R_DONE;
  }
  // (no branch taken: only reachable through NaN distances -- the reference loops as well)
  if (crossed) {  // the loads of the new cell are in flight before the tally traffic of the old one is issued
    close_visit(S, C, 0u);  // (the crossing is in c_cr already)
    enter_cell<PACKED>(S, C);
  }
  deposit(dep, dep_cell, S.loc_abs, S.loc_trk, lanes);
  if (dep) {
    S.loc_abs = 0.0;
    S.loc_trk = 0.0;
  }
  return r;
}

// sample_emission_group (src/sampling_functions.h:126-138) for the CDF draw `cdf` in cell `cell`, whose first group
// opacity is a0 (p_grp: the cell's cached abs_groups[g] * norm when all its groups are equal, 0 = not formed yet).
template <bool CLOSED_FORM = true>
__device__ __forceinline__ int sample_group(const PCtx &C, uint32_t cell, double a0, double &p_grp, double cdf) {
  const uint32_t G = C.G;
  int g = -1;
  if (CLOSED_FORM && C.uniform_groups && G <= 512) {
    // All groups of the cell hold the same opacity, so every step of the reference's walk subtracts the same
    // p = abs_groups[g] * norm and the walk stops at g = min{k : c_(k+1) <= 0}, c_(k+1) = fl(c_k - p).  The rounding
    // error accumulated over k <= G steps is below G * 2^-54, so when the residuals of the candidate k0 = floor(cdf*G)
    // clear zero by a margin far above that bound, the sequential result is provably k0 -- no loads, no dependent
    // chain, no divergence over the walk length.  Otherwise (probability ~1e-11 per scatter) fall through to the walk.
    if (p_grp == 0.0) p_grp = a0 * BG_DIV(1.0, a0 * (double)G);
    const int k0 = (int)(cdf * (double)G);
    const double before = fma(-(double)k0, p_grp, cdf);  // c_k0 up to rounding
    const double after = before - p_grp;                  // c_(k0+1)
    if (before > 1.0e-13 && after < -1.0e-13 && k0 < (int)G) g = k0;
  }
  if (g < 0) {
    // the sequential walk of the cell's group array, same arithmetic, loads batched by four
    const double *ag = C.opa + (uint64_t)cell * G;
    double a4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) a4[q] = (q < (int)G) ? __ldg(&ag[q]) : 0.0;
    const double norm = 1.0 / (a4[0] * (double)G);
    for (uint32_t base = 0;;) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (cdf > 0.0 && base + q < G) {
          g = (int)(base + q);
          cdf -= a4[q] * norm;
        }
      }
      base += 4;
      if (!(cdf > 0.0) || base >= G) break;  // g == G-1 here if the walk ran off the end (round-off guard)
#pragma unroll
      for (int q = 0; q < 4; ++q) a4[q] = (base + q < G) ? __ldg(&ag[base + q]) : 0.0;
    }
  }
  return g;
}

// get_uniform_angle (src/sampling_functions.h:57-70) from the Threefry words of its two draws.
__device__ __forceinline__ void scatter_direction(PState &S, const uint64_t w_mu, const uint64_t w_phi) {
  const double mu = u01_from_bits(w_mu) * 2.0 - 1.0;
  const double phi = u01_from_bits(w_phi) * 2.0 * K_PI;
#if BG_FM_DIV
  const double sin_theta = fm_sqrt(1.0 - mu * mu);  // mu in (-1, 1): the argument is a normal number
#else
  const double sin_theta = sqrt(1.0 - mu * mu);
#endif
  double sp, cp;
#if BG_FM_SINCOS
  fm_sincos(phi, &sp, &cp);
#else
  sincos(phi, &sp, &cp);
#endif
  S.ax = sin_theta * cp;
  S.ay = sin_theta * sp;
  S.az = mu;
}

// The scatter event (:94-101).  Its draws sit at consecutive counters -- mu, phi (get_uniform_angle,
// src/sampling_functions.h:57-70), the physical-vs-effective test (:98) and, if effective, the group CDF value
// (sample_emission_group, src/sampling_functions.h:126-138) -- and are evaluated as interleaved Threefry chains.  Every
// draw is CONSUMED exactly as in the reference (the counter is what aligns the stream); what is skipped is the
// evaluation of draws whose value provably cannot matter:
//   * with sigma_s == 0 the test is `u > 0`, true for every u01 value (u >= 2^-53): not evaluated when no lane of the
//     warp has sigma_s != 0;
//   * with one group the walk returns g = 0 for every c in (0,1): abs[0] * fl(1 / abs[0]) is 1 or 1 - 2^-53 and
//     u01 <= 1 - 2^-53, so the first subtraction already ends the loop;
//   * LAZY (history kernel): where all groups of every cell carry the same opacities (Cell::set_op_a fills them all,
//     src/cell.h:260-275 -- every reference deck), the group does not enter the physics, only the photon's final
//     record.  Only the LAST group draw of a history can be observed, so the scatter merely notes where it happened
//     (cell, counter) and finalize_group evaluates that one draw when the history ends.
// Every reference deck has sigma_s == 0 and faux-multigroup cells: an effective scatter then costs 2 Threefry
// evaluations instead of 4, with bit-identical photons, counters and tallies (tests/test_gpu_parity.py: post/group).
template <bool LAZY_OK>
__device__ __forceinline__ void scatter_event(PState &S, const PCtx &C, const unsigned lanes) {
  const bool test_void = !__any_sync(lanes, S.sig_s != 0.0);
  const bool cdf_void = C.G == 1u;
  const bool lazy = LAZY_OK && BG_LAZY_GROUP && C.uniform_groups && C.G <= 512u;
  uint64_t w_mu, w_phi, w_test = ~0ull, w_cdf = 0ull;
  if (test_void && (cdf_void || lazy)) {
    const int off[2] = {0, 1};
    uint64_t w[2];
    threefry2x64_20_w0_multi<2>(S.ctr, C.ctr_hi, S.stream, off, w);
    w_mu = w[0]; w_phi = w[1];
  } else if (test_void) {
    const int off[3] = {0, 1, 3};
    uint64_t w[3];
    threefry2x64_20_w0_multi<3>(S.ctr, C.ctr_hi, S.stream, off, w);
    w_mu = w[0]; w_phi = w[1]; w_cdf = w[2];
  } else {
    const int off[4] = {0, 1, 2, 3};
    uint64_t w[4];
    threefry2x64_20_w0_multi<4>(S.ctr, C.ctr_hi, S.stream, off, w);
    w_mu = w[0]; w_phi = w[1]; w_test = w[2]; w_cdf = w[3];
  }
  S.ctr += 3;
  ++S.c_sc;
  scatter_direction(S, w_mu, w_phi);
  // physical vs effective scatter (src/history_based_transport.h:98-100)
  // sigma_s == 0 (every reference deck): 0 / x is exactly +0, no division needed
  if (!test_void) {
    const double p_phys = (S.sig_s == 0.0) ? 0.0 : S.sig_s / ((1.0 - S.f) * S.sig_a + S.sig_s);
    if (!(u01_from_bits(w_test) > p_phys)) return;
  }
  S.ctr += 1;  // the group draw sits at counter S.ctr - 1
  if (cdf_void) return;  // g = 0 = the photon's group: nothing changes
  if (lazy) {
    S.grp_cell = S.cell;
    S.grp_ctr32 = (uint32_t)S.ctr - 1u;
    return;
  }
  // (with LAZY_OK the closed form's cells have returned above: S.p_grp is then dead state and costs no registers)
  const int g = sample_group<!(LAZY_OK && BG_LAZY_GROUP)>(C, S.cell, S.sig_a, S.p_grp, u01_from_bits(w_cdf));
  if ((uint32_t)g != S.group) {
    S.group = (uint32_t)g;
    if (!C.uniform_groups) load_xs(S, C);  // (same opacities in every group of the cell: nothing to fetch)
  }
}

// LAZY group sampling: the one group draw of the history that can be observed (see scatter_event).
__device__ __forceinline__ void finalize_group(PState &S, const PCtx &C) {
  if (S.grp_cell == ~0u) return;
  const uint64_t ctr = S.ctr - (uint64_t)((uint32_t)S.ctr - S.grp_ctr32);  // the pending draw's counter (< 2^32 back)
  const double cdf = u01_from_bits(threefry2x64_20_w0(ctr, C.ctr_hi, S.stream));
  double p = 0.0;
  S.group = (uint32_t)sample_group(C, S.grp_cell, __ldg(&C.opa[(uint64_t)S.grp_cell * C.G]), p, cdf);
  S.grp_cell = ~0u;
}

// Per-CTA statistics of a finished history.  Shared-memory 64-bit atomics are CAS loops on sm_100a, so the counters are
// kept as {low, high} 32-bit words with native 32-bit adds and an explicit carry.
__device__ __forceinline__ void stat_add32(uint32_t *s_stats, int which, uint32_t v) {
  const uint32_t old = atomicAdd(&s_stats[2 * which], v);
  if (old + v < old) atomicAdd(&s_stats[2 * which + 1], 1u);
}
// Running totals of the histories a lane has finished, kept in registers (the per-history update is five integer adds
// at the 1-2 lanes that retire per trip, where shared-memory atomics cost 12 instructions and a short-scoreboard wait)
// and flushed to the CTA's shared counters at the end of the kernel, or early if a 32-bit total is about to wrap.
struct LaneStats {
  uint32_t n, sc, cr, rf, lk;
};
__device__ __forceinline__ void lane_stats_flush(uint32_t *s_stats, LaneStats &L) {
  stat_add32(s_stats, ST_SCATTERS, L.sc);
  stat_add32(s_stats, ST_CROSSINGS, L.cr);
  stat_add32(s_stats, ST_REFLECTIONS, L.rf);
  stat_add32(s_stats, ST_LOOKUPS, L.lk);
  stat_add32(s_stats, ST_DEPOSITS, L.n);   // + crossings: added by stats_flush (one deposit per cell left + the last)
  stat_add32(s_stats, ST_EVENTS, L.n);     // + scatters + crossings + reflections: likewise (events_of_finished)
  L.n = L.sc = L.cr = L.rf = L.lk = 0u;
}
__device__ __forceinline__ void lane_stats_add(uint32_t *s_stats, LaneStats &L, const PState &S) {
  L.n += 1u; L.sc += S.c_sc; L.cr += S.c_cr; L.rf += S.c_rf; L.lk += S.c_lk;
  if ((L.sc | L.cr | L.rf | L.lk) & 0x80000000u) lane_stats_flush(s_stats, L);
}
// stats_flush for kernels that used LaneStats: ST_EVENTS / ST_DEPOSITS hold only the history count so far
__device__ __forceinline__ void stats_flush_derived(const uint32_t *s_stats, unsigned long long *g_stats) {
  if (threadIdx.x < 6) {
    auto val = [&](int w) { return ((unsigned long long)s_stats[2 * w + 1] << 32) | s_stats[2 * w]; };
    unsigned long long v = val((int)threadIdx.x);
    if (threadIdx.x == ST_EVENTS) v += val(ST_SCATTERS) + val(ST_CROSSINGS) + val(ST_REFLECTIONS);
    if (threadIdx.x == ST_DEPOSITS) v += val(ST_CROSSINGS);
    if (v) atomicAdd(&g_stats[threadIdx.x], v);
  }
}
__device__ __forceinline__ void stats_add(uint32_t *s_stats, const PState &S) {
  stat_add32(s_stats, ST_EVENTS, events_of_finished(S));
  stat_add32(s_stats, ST_SCATTERS, S.c_sc);
  stat_add32(s_stats, ST_CROSSINGS, S.c_cr);
  stat_add32(s_stats, ST_REFLECTIONS, S.c_rf);
  stat_add32(s_stats, ST_DEPOSITS, S.c_cr + 1u);  // one per cell left + the final one
  stat_add32(s_stats, ST_LOOKUPS, S.c_lk);
}
__device__ __forceinline__ void stats_flush(const uint32_t *s_stats, unsigned long long *g_stats) {
  if (threadIdx.x < 6) {
    const unsigned long long v = ((unsigned long long)s_stats[2 * threadIdx.x + 1] << 32) | s_stats[2 * threadIdx.x];
    if (v) atomicAdd(&g_stats[threadIdx.x], v);
  }
}

#ifndef BG_MIN_BLOCKS
#define BG_MIN_BLOCKS 5
#endif

// RESUME: the launch continues histories that the event-based passes (event.cuh) parked at a pending scatter: photon
// indices come from P.index_list, the thread-local tallies and counters from P.carry_*.
template <int MODE, bool COUNTERS, bool SMEM, bool RESUME, bool PACKED>
__global__ void __launch_bounds__(128, BG_MIN_BLOCKS) k_transport_history(const TransportParams P) {
  extern __shared__ double s_faces[];
  __shared__ uint32_t s_stats[12];
  if (threadIdx.x < 12) s_stats[threadIdx.x] = 0u;
  const double *faces;
  if (SMEM) {
    for (uint32_t t = threadIdx.x; t < P.mesh.n_faces; t += blockDim.x) s_faces[t] = P.mesh.faces[t];
    faces = s_faces;
  } else {
    faces = P.mesh.faces;
  }
  __syncthreads();
  PCtx C;
  C.fx = faces;
  C.fy = faces + (P.mesh.nx + 1);
  C.fz = C.fy + (P.mesh.ny + 1);
  C.nx = P.mesh.nx; C.ny = P.mesh.ny; C.nz = P.mesh.nz; C.G = P.mesh.G;
  C.sxy = C.nx * C.ny;
  C.f = P.f; C.opa = P.opa; C.ops = P.ops; C.cellrec = P.cellrec;
  // the counter's high word is seed << 32 (src/RNG.h:318-330): rebuilt from its upper half so that the compiler knows the
  // low 32 bits are zero and folds them out of the first Threefry round
  C.ctr_hi = (uint64_t)(uint32_t)(P.ctr_hi >> 32) << 32;
  C.uniform_groups = P.uniform_groups != 0;
  C.inv_sxy = P.inv_sxy; C.inv_nx = P.inv_nx;
  const unsigned FULL = 0xffffffffu;
  const unsigned lane_id = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane_id) - 1u;
  const uint32_t n_total = (uint32_t)P.n;

  // the warp's chunk of the work list [q_next, q_end); both are warp-uniform
  uint32_t q_next = 0, q_end = 0;
  bool exhausted = false;
  bool active = false;
  bool pending_scatter = false;
  PState S;
  S.x = S.y = S.z = S.ax = S.ay = S.az = S.E = S.E0 = S.life = 0.0;
  S.ctr = S.stream = 0;
  S.cell = S.group = 0;
  S.i = S.j = S.k = 0;
  S.f = S.sig_a = S.sig_s = S.loc_abs = S.loc_trk = 0.0;
  S.surface = 0;
  S.c_sc = S.c_cr = S.c_rf = S.c_lk = 0;
  S.ev_entry = 0;
  S.grp_cell = ~0u;
  S.grp_ctr32 = 0u;
  S.p_grp = 0.0;
  LaneStats LS{0u, 0u, 0u, 0u, 0u};
  uint32_t my_idx = 0;
  uint32_t ndep = 0;
  uint64_t dep_pos = 0;

  const uint32_t bcpack = pack_bc(P.mesh.bc);
  double2 *my_tally = P.tally;
  if (MODE == TM_ATOMIC && P.tally_copies > 1u) {
    const uint32_t copy = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % P.tally_copies;
    if (copy) my_tally = P.tally_rep + (size_t)(copy - 1u) * P.mesh.n_cells;
  }
  const uint32_t scatter_batch = P.scatter_batch;
  const bool aggregate = P.aggregate != 0;

  // Called by every lane of `lanes` together; `dep` says whether this lane has a deposit.
  auto deposit = [&](bool dep, uint32_t cell, double a, double t, unsigned lanes) {
    if (MODE == TM_ATOMIC) {
      if (aggregate) {
        // Warp-aggregated tallies: lanes depositing into the same cell are combined by a segmented tree reduction over
        // their peer set and the group's first lane issues the one pair of atomics.  Pays off where many photons sit
        // in few cells (marshak: 25 cells; hot_zone: a 5 x 5 hot corner), where same-address reductions serialise in L2.
        const unsigned dm = __ballot_sync(lanes, dep);
        if (dep) {
          const unsigned peers = __match_any_sync(dm, cell);
          const unsigned below = peers & lt_mask;
          if (__any_sync(dm, (peers & (peers - 1u)) != 0u)) {
            unsigned rel = __popc(below);                       // rank inside the group
            unsigned higher = peers & ~(lt_mask | (1u << lane_id));
            while (__any_sync(dm, higher != 0u)) {
              const int next = __ffs(higher);                   // nearest remaining peer above this lane (1-based)
              const int src = next ? next - 1 : (int)lane_id;
              const double ra = __shfl_sync(dm, a, src), rt = __shfl_sync(dm, t, src);
              if (next) { a += ra; t += rt; }
              higher &= ~__ballot_sync(dm, (rel & 1u) != 0u);   // odd ranks have been absorbed by their left peer
              rel >>= 1;
            }
          }
          if (below == 0u) {
            atomicAdd(&my_tally[cell].x, a);
            atomicAdd(&my_tally[cell].y, t);
          }
        }
      } else if (dep) {
        atomicAdd(&my_tally[cell].x, a);
        atomicAdd(&my_tally[cell].y, t);
      }
    } else if (MODE == TM_COUNT) {
      if (dep) ++ndep;
    } else if (dep) {
      P.dep_cell[dep_pos] = cell;
      P.dep_val[dep_pos] = make_double2(a, t);
      ++dep_pos;
    }
  };

  for (;;) {
    // ---------------- refill idle lanes ----------------
    // One attempt per trip and no inner loop: if the warp's chunk holds fewer photons than there are idle lanes, the
    // lanes left over stay idle for this trip and are served from the next chunk on the next one (once per chunk).
    const unsigned idle = __ballot_sync(FULL, !active);
    if (idle) {
      if (!exhausted && q_next == q_end) {
        unsigned long long base = 0;
        if (lane_id == 0) base = atomicAdd(P.work_counter, (unsigned long long)P.chunk);
        base = __shfl_sync(FULL, base, 0);
        if (base >= (unsigned long long)n_total) {
          exhausted = true;
        } else {
          q_next = (uint32_t)base;
          q_end = ((unsigned long long)n_total - base < P.chunk) ? n_total : (uint32_t)base + P.chunk;
        }
      }
      if (!exhausted) {
        const uint32_t avail = q_end - q_next;
        const uint32_t r = __popc(idle & lt_mask);
        if (!active && r < avail) {
          const uint32_t slot = q_next + r;
          const uint32_t idx = RESUME ? P.index_list[slot] : slot;
          pstate_load<PACKED>(S, P.ph, idx, C);
          my_idx = idx;
          if (RESUME) {
            const double2 acc = P.carry_acc[idx];
            const uint4 cn = P.carry_cnt[idx];
            S.loc_abs = acc.x; S.loc_trk = acc.y;
            S.c_sc = cn.y; S.c_cr = cn.z; S.c_rf = cn.w;
            S.c_lk = P.carry_lk[idx];
            S.ev_entry = cn.x;
            pending_scatter = P.resume_pending_scatter != 0;  // (pstate_load has fetched this visit's f, sigma_a, sigma_s)
          }
          ndep = 0;
          if (MODE == TM_LOG) dep_pos = P.dep_off[idx];
          active = true;
        }
        const uint32_t asked = __popc(idle);
        q_next += (asked < avail) ? asked : avail;
      } else if (idle == FULL) {
        break;  // nothing left to fetch and nobody is working
      }
    }

    // ---------------- one event for every lane that is not parked at a scatter ----------------
    const bool go = active && !pending_scatter;
    const unsigned adv = __ballot_sync(FULL, go);
    if (go) {
      uint8_t descriptor = EV_PASS;
      const int r = advance_event<PACKED>(S, C, bcpack, deposit, descriptor, adv);
      if (r == R_SCATTER) pending_scatter = true;
      if (r == R_DONE) {
        close_visit(S, C, 1u);  // (the retiring event is in no counter, events_of_finished)
        // (the group is observable only where the photon's record is: census photons, or everything in validation runs)
        if (!RESUME && (P.writeback_all || descriptor == EV_CENSUS)) finalize_group(S, C);
        if (MODE != TM_COUNT) lane_stats_add(s_stats, LS, S);
        const uint32_t idx = my_idx;
        if (MODE == TM_COUNT) {
          P.ndep[idx] = ndep;
        } else {
          P.desc[idx] = descriptor;
          P.ph.ee[idx] = make_double2(S.E, S.E0);
          if (P.writeback_all || descriptor == EV_CENSUS) pstate_store_full(S, P.ph, idx);
          if (COUNTERS)
            reinterpret_cast<uint4 *>(P.counters)[idx] = make_uint4(events_of_finished(S), S.c_sc, S.c_cr, S.c_rf);
        }
        active = false;
      }
    }

    // ---------------- scatters, sampled by the warp's parked lanes together ----------------
    // The scatter is the expensive event (4 of its 5 Threefry evaluations, sincos, sqrt).  A lane that reaches one
    // parks; the warp samples the parked scatters once `scatter_batch` lanes wait, or when no lane is left that could
    // advance instead.  In scattering-dominated cycles nearly every lane parks on every trip and this is the old
    // one-event-per-trip loop; in mixed cycles (big_cube: 1 event in 6 is a scatter) the scatter code runs at
    // >= scatter_batch lanes instead of ~5.  Per-photon results cannot change: a history depends on nothing but its own
    // state.
    const unsigned parked = __ballot_sync(FULL, pending_scatter);
    if (parked) {
      const unsigned movable = __ballot_sync(FULL, active && !pending_scatter);
      if ((uint32_t)__popc(parked) >= scatter_batch || movable == 0u) {
        if (pending_scatter) {
          scatter_event<!RESUME>(S, C, parked);
          pending_scatter = false;
        }
      }
    }
  }

  // ---------------- statistics: one global atomic per CTA and counter ----------------
  if (MODE != TM_COUNT) lane_stats_flush(s_stats, LS);
  __syncthreads();
  if (MODE != TM_COUNT) stats_flush_derived(s_stats, P.stats);
}

// tally[i] += rep[0][i] + rep[1][i] + ... in copy order (a fixed order: the fold adds no run-to-run noise of its own),
// and the copies are zeroed for the next launch.
__global__ void k_fold_tally(double2 *tally, double2 *rep, uint32_t n_cells, uint32_t n_rep) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cells) return;
  double2 acc = tally[i];
  for (uint32_t r = 0; r < n_rep; ++r) {
    const double2 v = rep[(size_t)r * n_cells + i];
    acc.x += v.x;
    acc.y += v.y;
    rep[(size_t)r * n_cells + i] = make_double2(0.0, 0.0);
  }
  tally[i] = acc;
}

}  // namespace bg
