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
//   * every warp keeps 8 * POOL_K = 56 photon slots in shared memory (176 bytes each, eleven 16-byte vectors laid out
//     [vector][slot], so vector v of slot s sits in the 16-byte bank group s % 8); the slots are shared by "classes" of
//     four lanes -- lanes l, l + 8, l + 16, l + 24 own the slots s with s % 8 == l % 8 -- so the eight lanes of a quarter warp
//     always touch eight different bank groups: every LDS.128 / STS.128 is conflict-free whichever slots are picked;
//   * a slot is EMPTY, ADVANCE (next: one trip of the reference's loop), SCATTER (parked at a scatter) or DONE (history
//     finished, record not written yet); a class's eight slot states live in one register (four 8-bit masks) that its
//     four lanes keep in step with two shuffles per trip;
//   * each trip the warp votes on ONE event type -- the one most lanes can serve -- and runs that block alone: A
//     (advance_event: distance sampling, implicit capture, boundary handling), S (scatter_event: the angle draws and
//     the group), R (write the finished record, fetch the next photon of the work list into the slot); the k-th lane of
//     a class takes the class's k-th slot of the elected type.  With seven slots behind four lanes a lane almost
//     always finds one, so every block runs near full width instead of 27 / 14 / 7 lanes.
//
// The blocks are the SAME device functions the history kernel runs (transport.cuh), per photon in the same order, so the
// per-photon results are identical by construction (reference SURVEY note N5: a history depends on nothing but its own
// state); only the order in which the atomic tallies are summed differs.
#pragma once
#include "transport.cuh"

namespace bg {

// slots per class of four lanes (<= 8: one byte of the mask register per state).  Seven slots (1.75 per lane) leave
// room for five CTAs per SM instead of four; A/B in profiles/pool_variants_r02.txt
#ifndef POOL_K
#define POOL_K 7
#endif
#ifndef POOL_MIN_BLOCKS
#define POOL_MIN_BLOCKS 5
#endif
#define POOL_ROWS 2  // (work-list photons per resident thread the launcher assumes when it trims the grid)
#ifndef POOL_STICKY_LANES
#define POOL_STICKY_LANES 28u  // an advance block re-runs without an election while it serves at least this many lanes
#endif

// slot states = byte index of the class's mask register
enum : uint32_t { PM_ADV = 0u, PM_SCAT = 1u, PM_DONE = 2u, PM_EMPTY = 3u };
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
constexpr int POOL_SLOTS = 8 * POOL_K;  // per warp (slot s = 8 j + class, j < POOL_K)
constexpr size_t POOL_BYTES_PER_WARP = (size_t)POOL_SLOTS * PV_N * 16;
constexpr int PV_STRIDE = POOL_SLOTS * 16;  // bytes between consecutive vectors of a slot

struct PoolSlot {
  char *base;  // vector v of this slot at base + v * PV_STRIDE
  __device__ __forceinline__ double2 ld2(int v) const { return *reinterpret_cast<const double2 *>(base + v * PV_STRIDE); }
  __device__ __forceinline__ uint4 ld4(int v) const { return *reinterpret_cast<const uint4 *>(base + v * PV_STRIDE); }
  __device__ __forceinline__ void st2(int v, double a, double b) const {
    *reinterpret_cast<double2 *>(base + v * PV_STRIDE) = make_double2(a, b);
  }
  __device__ __forceinline__ void st4(int v, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const {
    *reinterpret_cast<uint4 *>(base + v * PV_STRIDE) = make_uint4(a, b, c, d);
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
  __shared__ uint8_t s_nth[(1 << POOL_K) * 4];
  const TransportParams &P = Q.T;
  if (threadIdx.x < 12) s_stats[threadIdx.x] = 0u;
  for (uint32_t e = threadIdx.x; e < (4u << POOL_K); e += blockDim.x) {  // s_nth[m][r] = index of the r-th set bit of m (0 if none)
    uint32_t m = e >> 2;
    for (uint32_t i = 0; i < (e & 3u); ++i) m &= m - 1u;
    s_nth[e] = (uint8_t)(m ? __ffs((int)m) - 1 : 0);
  }
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
  // this lane's class (l % 8) and its rank inside the class (l / 8): the rank-th slot of the elected state is its slot
  const uint32_t cls = lane_id & 7u, rank = lane_id >> 3;
  char *const cls_pool = s_dyn + warp_id * POOL_BYTES_PER_WARP + cls * 16;
  auto slot_of = [&](int j) { return PoolSlot{cls_pool + (size_t)j * 128}; };  // slot s = 8 j + cls: byte offset 16 s
  // the j of the class's rank-th slot whose bit is set in the 8-bit mask m (the caller knows there is one): a 512-byte
  // table in shared memory, [mask][rank] -> bit index, filled once per CTA
  const uint8_t *const my_nth = s_nth + rank;
  auto pick = [&](uint32_t m) { return (int)my_nth[(m & 0xffu) * 4u]; };

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

  // the class's slot states: byte PM_x = mask of the slots in state x (identical in the four lanes of the class)
  uint32_t masks = ((1u << POOL_K) - 1u) << (8 * PM_EMPTY);
  uint32_t q_next = 0, q_end = 0;  // the warp's chunk of the work list (warp-uniform)
  bool exhausted = false;
  LaneStats LS{0u, 0u, 0u, 0u, 0u};
  const uint32_t T_S = Q.batch_scatter, T_R = Q.batch_refill;

  bool last_was_A = false;
  for (;;) {
    const uint32_t mA = masks >> (8 * PM_ADV);
    const bool hasA = rank < (uint32_t)__popc(mA & 0xffu);
    const unsigned bA = __ballot_sync(FULL, hasA);
    const uint32_t nA = __popc(bA);
    uint32_t mS = 0u;
    bool hasS = false, hasR = false;
    unsigned bS = 0u;
    int block = 0;  // 0 A, 1 S, 2 R
    // An advance block that still fills the warp follows the previous one without an election (the other two counts
    // are not even formed); the scatters and finished records it leaves waiting are served, by more lanes at once,
    // as soon as it does not.
    if (!(last_was_A && nA >= POOL_STICKY_LANES)) {
      mS = masks >> (8 * PM_SCAT);
      const uint32_t mR = (masks >> (8 * PM_DONE)) | (exhausted ? 0u : (masks >> (8 * PM_EMPTY)));
      hasS = rank < (uint32_t)__popc(mS & 0xffu);
      hasR = rank < (uint32_t)__popc(mR & 0xffu);
      bS = __ballot_sync(FULL, hasS);
      const unsigned bR = __ballot_sync(FULL, hasR);
      if ((bA | bS | bR) == 0u) break;
      const uint32_t nS = __popc(bS), nR = __popc(bR);
      // the election: the block that serves the most lanes, with a head start for A (the common event)
      if (nR && (nR >= T_R || (nR >= nA && nR >= nS))) block = 2;
      else if (nS && (nS >= T_S || nS >= nA)) block = 1;
    }
    last_was_A = block == 0;
    uint32_t moved = 0u;  // this lane's slot in its new state's byte (the class ORs these together below)

    if (block == 0) {
      // ---------------- A: one trip of the reference's loop up to the event dispatch ----------------
      if (hasA) {
        const int j = pick(mA);
        const PoolSlot sl = slot_of(j);
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
        moved = (1u << j) << (8 * m);
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
        const int j = pick(mS);
        const PoolSlot sl = slot_of(j);
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
        moved = (1u << j) << (8 * PM_ADV);
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
      // the class's finished slots first, then its empty ones
      int j = -1;
      bool retire = false;
      if (hasR) {
        const uint32_t mD = (masks >> (8 * PM_DONE)) & 0xffu;
        // rank-th slot of (finished, then empty): the finished ones occupy the first ranks
        const uint32_t nD = (uint32_t)__popc(mD);
        if (rank < nD) {
          j = pick(mD);
          retire = true;
        } else {  // (hasR: the class has more than rank - nD empty slots)
          j = (int)s_nth[((masks >> (8 * PM_EMPTY)) & 0xffu) * 4u + (rank - nD)];
        }
      }
      const PoolSlot sl = slot_of(j < 0 ? 0 : j);
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
        moved = (1u << j) << (8 * PM_EMPTY);
      }
      // refill (warp-uniform chunk bookkeeping, as in k_transport_history)
      const bool fill = j >= 0 && !exhausted;
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
            moved = (1u << j) << (8 * PM_ADV);
          }
          const uint32_t asked = __popc(want);
          q_next += (asked < avail) ? asked : avail;
        }
      }
    }
    // the class's four lanes merge what they moved: a moved slot leaves whatever state it was in
    moved |= __shfl_xor_sync(FULL, moved, 8);
    moved |= __shfl_xor_sync(FULL, moved, 16);
    uint32_t touched = moved | (moved >> 16);
    touched = (touched | (touched >> 8)) & 0xffu;
    masks = (masks & ~(touched * 0x01010101u)) | moved;
    // the slots are shared by the lanes of a class from one trip to the next: order this trip's shared-memory stores
    // before the next trip's loads (the shuffles above converge the warp but are not memory barriers)
    __syncwarp();
  }

  lane_stats_flush(s_stats, LS);
  __syncthreads();
  stats_flush_derived(s_stats, P.stats);
}

}  // namespace bg
