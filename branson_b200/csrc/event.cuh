// event.cuh -- event-based transport variant (BGPU_EVENT).
//
// The reference's event_based_transport.h (CPU only, :341-427) regroups the history loop by event type over an active
// list so that each loop is uniform work.  The idea maps to the GPU as warp-divergence control: the expensive event is
// the (effective) scatter -- 4 of the ~5 Threefry draws of an event, sincos, sqrt and the sequential group walk -- so
// the photons are regrouped at scatters:
//
//   pass k:   two active lists come out of pass k-1: photons parked AT a scatter, and photons parked after a cell
//             crossing / reflection.  The scatter list is processed by a launch in which every lane samples its scatter
//             (100% lane efficiency in that phase) and then advances; the other list by a launch that only advances.
//             "Advance" = at most EV_MAX_ADVANCE trips of the reference's loop (distance sampling, implicit capture,
//             boundary handling), so no lane waits for a neighbour that streams through many cells.  Finished
//             histories are written back; the rest is appended (warp-aggregated) to the two lists of pass k+1.
//
// The kernels carry no refill logic and no divergent scatter branch; the price is that photon state streams through
// HBM once per pass (six 128-bit streams + 36 bytes of carried thread-local tallies and counters).  When the active
// lists no longer fill the machine the remaining histories are finished by the persistent history kernel in RESUME
// mode.  Measured on the 30-group hohlraum the HISTORY kernel stays the faster of the two.
//
// Per-photon results are identical to the HISTORY variant by construction (the same advance_event / scatter_event
// device functions run in the same per-photon order; only the tally summation order differs) -- unlike the reference,
// whose EVENT path is not photon-identical to its HISTORY path (note E).
#pragma once
#include <string>

#include "transport.cuh"

namespace bg {

#ifndef EV_MAX_ADVANCE
#define EV_MAX_ADVANCE 2
#endif

struct EventParams {
  TransportParams T;
  const uint32_t *list_in;   // nullptr on the first pass: identity
  uint64_t n_in;
  uint32_t *scatter_out;     // photons parked at a scatter
  uint32_t *cont_out;        // photons parked after EV_MAX_ADVANCE non-scatter events
  unsigned long long *n_out; // [0] scatter list size, [1] continue list size
  double2 *acc;              // carried {loc_abs, loc_trk}
  uint4 *cnt;                // carried counters
  uint32_t *lk;
  int first;                 // no carried state yet
  int pending_scatter;       // list_in holds photons parked at a scatter
};

template <bool COUNTERS, bool SMEM>
__global__ void __launch_bounds__(128, 4) k_event_pass(const EventParams E) {
  extern __shared__ double s_faces[];
  __shared__ uint32_t s_stats[12];
  const TransportParams &P = E.T;
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
  C.ctr_hi = P.ctr_hi;
  C.uniform_groups = P.uniform_groups != 0;
  C.inv_sxy = P.inv_sxy; C.inv_nx = P.inv_nx;
  const unsigned lane_id = threadIdx.x & 31u;

  const uint32_t bcpack = pack_bc(P.mesh.bc);
  auto deposit = [&](bool dep, uint32_t cell, double a, double t, unsigned) {
    if (dep) {
      atomicAdd(&P.tally[cell].x, a);
      atomicAdd(&P.tally[cell].y, t);
    }
  };

  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t n_round = (E.n_in + 31) & ~31ull;  // whole warps stay converged for the ballots below
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += stride) {
    const bool valid = t < E.n_in;
    int park = 0;  // 1: at a scatter, 2: after a non-scatter event
    uint64_t idx = 0;
    if (valid) {
      idx = E.list_in ? (uint64_t)E.list_in[t] : t;
      PState S;
      S.surface = 0;
      S.p_grp = 0.0;
      pstate_load(S, P.ph, idx, C);
      if (!E.first) {
        const double2 acc = E.acc[idx];
        const uint4 cn = E.cnt[idx];
        S.loc_abs = acc.x; S.loc_trk = acc.y;
        S.c_sc = cn.y; S.c_cr = cn.z; S.c_rf = cn.w;
        S.c_lk = E.lk[idx];
        S.ev_entry = cn.x;
      }
      // the parked scatter: every lane of the warp is here together (pstate_load has fetched f, sigma_a, sigma_s)
      if (E.pending_scatter) scatter_event<false>(S, C, __activemask());
      uint8_t descriptor = EV_PASS;
      int r = R_CONTINUE;
#pragma unroll 1
      for (int step = 0; step < EV_MAX_ADVANCE && r == R_CONTINUE; ++step)
        r = advance_event(S, C, bcpack, deposit, descriptor, 0u);
      if (r == R_DONE) {
        close_visit(S, C, 1u);
        stats_add(s_stats, S);
        P.desc[idx] = descriptor;
        P.ph.ee[idx] = make_double2(S.E, S.E0);
        if (P.writeback_all || descriptor == EV_CENSUS) pstate_store_full(S, P.ph, idx);
        if (COUNTERS)
          reinterpret_cast<uint4 *>(P.counters)[idx] = make_uint4(events_of_finished(S), S.c_sc, S.c_cr, S.c_rf);
      } else {  // parked
        P.ph.ee[idx] = make_double2(S.E, S.E0);
        pstate_store_full(S, P.ph, idx);
        E.acc[idx] = make_double2(S.loc_abs, S.loc_trk);
        E.cnt[idx] = make_uint4(S.ev_entry, S.c_sc, S.c_cr, S.c_rf);  // .x: the visit goes on in the next pass
        E.lk[idx] = S.c_lk;
        park = (r == R_SCATTER) ? 1 : 2;
      }
    }
    // warp-aggregated append of the survivors to the two active lists of the next pass
#pragma unroll
    for (int kind = 1; kind <= 2; ++kind) {
      const unsigned m = __ballot_sync(0xffffffffu, park == kind);
      if (m) {
        unsigned long long base = 0;
        const int leader = __ffs(m) - 1;
        if ((int)lane_id == leader) base = atomicAdd(&E.n_out[kind - 1], (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (park == kind)
          (kind == 1 ? E.scatter_out : E.cont_out)[base + __popc(m & ((1u << lane_id) - 1u))] = (uint32_t)idx;
      }
    }
  }
  __syncthreads();
  stats_flush(s_stats, P.stats);
}

}  // namespace bg
