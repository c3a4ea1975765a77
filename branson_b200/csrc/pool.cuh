// pool.cuh -- event-based transport with the event queues in SHARED MEMORY (BGPU_EVENT).
//
// The reference's event_based_transport.h (:341-427) regroups the history loop by event type over an active list, so
// that each inner loop does uniform work.  On the GPU the same idea is warp-divergence control: in k_transport_history a
// warp's 32 lanes sit at different points of their histories -- some advancing, some waiting at a scatter, some between
// photons -- and on decks that mix event types (big_cube, hot_zone: one event in six is a scatter) a third of the issued
// lanes are idle (ncu: 22 of 32).  The first event-based variant of this library (event.cuh) regroups through HBM: every
// pass streams the photon state out and back, which costs 8-10x on scattering decks.  Here the regrouping costs shared
// memory traffic only:
//
//   * every lane owns POOL_ROWS photon slots in shared memory (176 bytes each, eleven 16-byte vectors laid out
//     [vector][lane] so that a warp's access to one vector of one row is a conflict-free 512-byte LDS.128 / STS.128);
//   * a slot is EMPTY, ADVANCE (next: one trip of the reference's loop), SCATTER (parked at a scatter) or DONE (history
//     finished, record not written yet);
//   * each trip the warp votes on ONE event type -- the one most lanes can serve from either of their slots -- and runs
//     that block for those lanes: A (advance_event: distance sampling, implicit capture, boundary handling), S
//     (scatter_event: the angle draws and the group), R (write the finished record, fetch the next photon of the work
//     list into the slot).  With two slots per lane a lane almost always holds a photon of the elected type, so each
//     block runs at ~30 lanes instead of 27 / 14 / 7.
//
// The blocks are the SAME device functions the history kernel runs (transport.cuh), per photon in the same order, so the
// per-photon results are identical by construction (reference SURVEY note N5: a history depends on nothing but its own
// state); only the order in which the atomic tallies are summed differs.
#pragma once
#include "transport.cuh"

namespace bg {

#ifndef POOL_ROWS
#define POOL_ROWS 2
#endif
#ifndef POOL_MIN_BLOCKS
#define POOL_MIN_BLOCKS 4
#endif

enum : uint32_t { PM_EMPTY = 0u, PM_ADV = 1u, PM_SCAT = 2u, PM_DONE = 3u };
// vectors of a slot (16 bytes each)
enum : int {
  PV_XY = 0,    // x, y
  PV_ZL,        // z, life_dx
  PV_AXY,       // angle.x, angle.y
  PV_AZC,       // angle.z, RNG counter (bits)
  PV_EE,        // E, E0
  PV_SG,        // RNG stream, cell | group << 32
  PV_FA,        // f, sigma_a
  PV_SK,        // sigma_s, {i | j << 16, k | descriptor << 24}
  PV_LOC,       // loc_abs, loc_trk
  PV_CNT,       // c_cr, c_rf, c_lk, ev_entry      (written by the A block)
  PV_SCT,       // c_sc, grp_cell, grp_ctr32, idx  (written by the S block)
  PV_N
};
constexpr size_t POOL_BYTES_PER_WARP = (size_t)POOL_ROWS * PV_N * 32 * 16;

struct PoolSlot {
  char *base;  // this lane's column of one row: vector v at base + v * 512
  __device__ __forceinline__ double2 ld2(int v) const { return *reinterpret_cast<const double2 *>(base + v * 512); }
  __device__ __forceinline__ uint4 ld4(int v) const { return *reinterpret_cast<const uint4 *>(base + v * 512); }
  __device__ __forceinline__ void st2(int v, double a, double b) const {
    *reinterpret_cast<double2 *>(base + v * 512) = make_double2(a, b);
  }
  __device__ __forceinline__ void st4(int v, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const {
    *reinterpret_cast<uint4 *>(base + v * 512) = make_uint4(a, b, c, d);
  }
};

__device__ __forceinline__ double u2d(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }
__device__ __forceinline__ uint64_t d2u(double d) { return (uint64_t)__double_as_longlong(d); }

// everything an A block needs (the whole record)
__device__ __forceinline__ void pool_load_all(PState &S, const PoolSlot &sl, uint32_t &idx, uint8_t &desc) {
  const double2 xy = sl.ld2(PV_XY), zl = sl.ld2(PV_ZL), axy = sl.ld2(PV_AXY), azc = sl.ld2(PV_AZC), ee = sl.ld2(PV_EE);
  const double2 sg = sl.ld2(PV_SG), fa = sl.ld2(PV_FA), loc = sl.ld2(PV_LOC);
  const uint4 sk = sl.ld4(PV_SK), cnt = sl.ld4(PV_CNT), sct = sl.ld4(PV_SCT);
  S.x = xy.x; S.y = xy.y; S.z = zl.x; S.life = zl.y;
  S.ax = axy.x; S.ay = axy.y; S.az = azc.x; S.ctr = d2u(azc.y);
  S.E = ee.x; S.E0 = ee.y;
  S.stream = d2u(sg.x);
  const uint64_t cg = d2u(sg.y);
  S.cell = (uint32_t)cg; S.group = (uint32_t)(cg >> 32);
  S.f = fa.x; S.sig_a = fa.y;
  S.sig_s = u2d(sk.x, sk.y);
  S.i = (int)(sk.z & 0xffffu); S.j = (int)(sk.z >> 16); S.k = (int)(sk.w & 0xffffffu);
  desc = (uint8_t)(sk.w >> 24);
  S.loc_abs = loc.x; S.loc_trk = loc.y;
  S.c_cr = cnt.x; S.c_rf = cnt.y; S.c_lk = cnt.z; S.ev_entry = cnt.w;
  S.c_sc = sct.x; S.grp_cell = sct.y; S.grp_ctr32 = sct.z; idx = sct.w;
  S.surface = 0u;
  S.p_grp = 0.0;
}

__device__ __forceinline__ void pool_store_cell(const PState &S, const PoolSlot &sl, uint8_t desc) {
  sl.st2(PV_SG, __longlong_as_double((long long)S.stream),
         __longlong_as_double((long long)((uint64_t)S.cell | ((uint64_t)S.group << 32))));
  sl.st2(PV_FA, S.f, S.sig_a);
  const uint64_t ss = d2u(S.sig_s);
  sl.st4(PV_SK, (uint32_t)ss, (uint32_t)(ss >> 32), (uint32_t)S.i | ((uint32_t)S.j << 16),
         (uint32_t)S.k | ((uint32_t)desc << 24));
}

struct PoolParams {
  TransportParams T;
  uint32_t batch_scatter;  // lanes that must hold a parked scatter before the S block is elected over a fuller A block
  uint32_t batch_refill;   // ... a finished / empty slot, before the R block is
};

template <bool COUNTERS, bool SMEM, bool PACKED>
__global__ void __launch_bounds__(128, POOL_MIN_BLOCKS) k_transport_pool(const PoolParams Q) {
  extern __shared__ __align__(16) char s_dyn[];  // [pool: 4 warps][faces]
  __shared__ uint32_t s_stats[12];
  const TransportParams &P = Q.T;
  if (threadIdx.x < 12) s_stats[threadIdx.x] = 0u;
  double *s_faces = reinterpret_cast<double *>(s_dyn + 4 * POOL_BYTES_PER_WARP);
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
  C.ctr_hi = (uint64_t)(uint32_t)(P.ctr_hi >> 32) << 32;
  C.uniform_groups = P.uniform_groups != 0;
  C.inv_sxy = P.inv_sxy; C.inv_nx = P.inv_nx;
  const unsigned FULL = 0xffffffffu;
  const unsigned lane_id = threadIdx.x & 31u, warp_id = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane_id) - 1u;
  const uint32_t n_total = (uint32_t)P.n;
  const uint32_t bcpack = pack_bc(P.mesh.bc);
  char *const my_pool = s_dyn + warp_id * POOL_BYTES_PER_WARP + lane_id * 16;
  auto slot_of = [&](uint32_t row) { return PoolSlot{my_pool + (size_t)row * (PV_N * 512)}; };

  double2 *my_tally = P.tally;
  if (P.tally_copies > 1u) {
    const uint32_t copy = (blockIdx.x * (blockDim.x >> 5) + warp_id) % P.tally_copies;
    if (copy) my_tally = P.tally_rep + (size_t)(copy - 1u) * P.mesh.n_cells;
  }
  auto deposit = [&](bool dep, uint32_t cell, double a, double t, unsigned) {
    if (dep) {
      atomicAdd(&my_tally[cell].x, a);
      atomicAdd(&my_tally[cell].y, t);
    }
  };

  uint32_t mode[POOL_ROWS];
#pragma unroll
  for (int r = 0; r < POOL_ROWS; ++r) mode[r] = PM_EMPTY;
  uint32_t q_next = 0, q_end = 0;  // the warp's chunk of the work list (warp-uniform)
  bool exhausted = false;
  LaneStats LS{0u, 0u, 0u, 0u, 0u};
  const uint32_t T_S = Q.batch_scatter, T_R = Q.batch_refill;

  for (;;) {
    bool hasA = false, hasS = false, hasD = false, hasE = false;
#pragma unroll
    for (int r = 0; r < POOL_ROWS; ++r) {
      hasA = hasA || mode[r] == PM_ADV;
      hasS = hasS || mode[r] == PM_SCAT;
      hasD = hasD || mode[r] == PM_DONE;
      hasE = hasE || mode[r] == PM_EMPTY;
    }
    const bool hasR = hasD || (hasE && !exhausted);
    const unsigned bA = __ballot_sync(FULL, hasA), bS = __ballot_sync(FULL, hasS), bR = __ballot_sync(FULL, hasR);
    if ((bA | bS | bR) == 0u) break;
    const uint32_t nA = __popc(bA), nS = __popc(bS), nR = __popc(bR);
    // the election: the block that serves the most lanes, with a head start for A (the common event)
    int block;  // 0 A, 1 S, 2 R
    if (nR && (nR >= T_R || (nR >= nA && nR >= nS))) block = 2;
    else if (nS && (nS >= T_S || nS >= nA)) block = 1;
    else block = 0;

    if (block == 0) {
      // ---------------- A: one trip of the reference's loop up to the event dispatch ----------------
      if (hasA) {
        uint32_t row = 0;
#pragma unroll
        for (int r = POOL_ROWS - 1; r >= 0; --r) row = (mode[r] == PM_ADV) ? (uint32_t)r : row;
        const PoolSlot sl = slot_of(row);
        PState S;
        uint32_t idx;
        uint8_t descriptor;
        pool_load_all(S, sl, idx, descriptor);
        const uint32_t cell0 = S.cell;
        const uint32_t rf0 = S.c_rf;
        const int res = advance_event<PACKED>(S, C, bcpack, deposit, descriptor, bA);
        uint32_t m = PM_ADV;
        if (res == R_SCATTER) m = PM_SCAT;
        if (res == R_DONE) {
          close_visit(S, C, 1u);
          m = PM_DONE;
        }
#pragma unroll
        for (int r = 0; r < POOL_ROWS; ++r) mode[r] = ((uint32_t)r == row) ? m : mode[r];
        sl.st2(PV_XY, S.x, S.y);
        sl.st2(PV_ZL, S.z, S.life);
        sl.st2(PV_AZC, S.az, __longlong_as_double((long long)S.ctr));
        sl.st2(PV_EE, S.E, S.E0);
        sl.st2(PV_LOC, S.loc_abs, S.loc_trk);
        if (S.cell != cell0 || S.c_rf != rf0 || res == R_DONE) {  // crossed, reflected or finished
          sl.st2(PV_AXY, S.ax, S.ay);
          pool_store_cell(S, sl, descriptor);
          sl.st4(PV_CNT, S.c_cr, S.c_rf, S.c_lk, S.ev_entry);
        }
      }
    } else if (block == 1) {
      // ---------------- S: the parked scatters, sampled together ----------------
      if (hasS) {
        uint32_t row = 0;
#pragma unroll
        for (int r = POOL_ROWS - 1; r >= 0; --r) row = (mode[r] == PM_SCAT) ? (uint32_t)r : row;
        const PoolSlot sl = slot_of(row);
        PState S;
        const double2 azc = sl.ld2(PV_AZC), sg = sl.ld2(PV_SG), fa = sl.ld2(PV_FA);
        const uint4 sk = sl.ld4(PV_SK), sct = sl.ld4(PV_SCT);
        S.ctr = d2u(azc.y);
        S.stream = d2u(sg.x);
        const uint64_t cg = d2u(sg.y);
        S.cell = (uint32_t)cg; S.group = (uint32_t)(cg >> 32);
        S.f = fa.x; S.sig_a = fa.y; S.sig_s = u2d(sk.x, sk.y);
        S.c_sc = sct.x; S.grp_cell = sct.y; S.grp_ctr32 = sct.z;
        S.p_grp = 0.0;
        const uint32_t group0 = S.group;
        scatter_event<true>(S, C, bS);
#pragma unroll
        for (int r = 0; r < POOL_ROWS; ++r) mode[r] = ((uint32_t)r == row) ? PM_ADV : mode[r];
        sl.st2(PV_AXY, S.ax, S.ay);
        sl.st2(PV_AZC, S.az, __longlong_as_double((long long)S.ctr));
        sl.st4(PV_SCT, S.c_sc, S.grp_cell, S.grp_ctr32, sct.w);
        if (S.group != group0) {  // (general multigroup decks only: the lazy path never changes the group here)
          sl.st2(PV_SG, sg.x, __longlong_as_double((long long)((uint64_t)S.cell | ((uint64_t)S.group << 32))));
          sl.st2(PV_FA, S.f, S.sig_a);
          const uint64_t ss = d2u(S.sig_s);
          sl.st4(PV_SK, (uint32_t)ss, (uint32_t)(ss >> 32), sk.z, sk.w);
        }
      }
    } else {
      // ---------------- R: write finished records, fetch the next photons ----------------
      // one slot per lane and trip: a finished one first, else an empty one
      uint32_t row = 0;
      bool retire = false, fill = false;
      if (hasR) {
#pragma unroll
        for (int r = POOL_ROWS - 1; r >= 0; --r) row = (mode[r] == PM_EMPTY) ? (uint32_t)r : row;
#pragma unroll
        for (int r = POOL_ROWS - 1; r >= 0; --r) row = (mode[r] == PM_DONE) ? (uint32_t)r : row;
        retire = hasD;
        fill = !exhausted;
      }
      const PoolSlot sl = slot_of(row);
      if (retire) {
        PState S;
        uint32_t idx;
        uint8_t descriptor;
        pool_load_all(S, sl, idx, descriptor);
        // (the group is observable only where the photon's record is: census photons, or everything in validation runs)
        const bool full = P.writeback_all || descriptor == EV_CENSUS;
        if (full) finalize_group(S, C);
        lane_stats_add(s_stats, LS, S);
        P.desc[idx] = descriptor;
        P.ph.ee[idx] = make_double2(S.E, S.E0);
        if (full) pstate_store_full(S, P.ph, idx);
        if (COUNTERS)
          reinterpret_cast<uint4 *>(P.counters)[idx] = make_uint4(events_of_finished(S), S.c_sc, S.c_cr, S.c_rf);
#pragma unroll
        for (int r = 0; r < POOL_ROWS; ++r) mode[r] = ((uint32_t)r == row) ? PM_EMPTY : mode[r];
      }
      // refill (warp-uniform chunk bookkeeping, as in k_transport_history)
      const unsigned want = __ballot_sync(FULL, fill);
      if (want) {
        if (q_next == q_end) {
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
          const uint32_t rnk = __popc(want & lt_mask);
          if (fill && rnk < avail) {
            const uint32_t idx = q_next + rnk;
            PState S;
            S.surface = 0u;
            pstate_load<PACKED>(S, P.ph, idx, C);
            sl.st2(PV_XY, S.x, S.y);
            sl.st2(PV_ZL, S.z, S.life);
            sl.st2(PV_AXY, S.ax, S.ay);
            sl.st2(PV_AZC, S.az, __longlong_as_double((long long)S.ctr));
            sl.st2(PV_EE, S.E, S.E0);
            pool_store_cell(S, sl, EV_PASS);
            sl.st2(PV_LOC, 0.0, 0.0);
            sl.st4(PV_CNT, 0u, 0u, 0u, 0u);
            sl.st4(PV_SCT, 0u, ~0u, 0u, idx);
#pragma unroll
            for (int r = 0; r < POOL_ROWS; ++r) mode[r] = ((uint32_t)r == row) ? PM_ADV : mode[r];
          }
          const uint32_t asked = __popc(want);
          q_next += (asked < avail) ? asked : avail;
        }
      }
    }
    __syncwarp();
  }

  lane_stats_flush(s_stats, LS);
  __syncthreads();
  stats_flush_derived(s_stats, P.stats);
}

}  // namespace bg
