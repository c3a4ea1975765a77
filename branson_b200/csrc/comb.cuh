// comb.cuh -- population control of the census on the device.
//
// Follows the reference's comb_photons (src/census_functions.h:48-93; defined there but without a call site in this
// snapshot -- SURVEY section 8f item 2).  In census-list order every photon draws once from ONE generator and survives
// with probability E / comb_photon_E, comb_photon_E = global census energy / max_census_photons (:65,:74-78); a cell's
// last photon survives regardless (:78: `cell_census_count[cell] == 1`); survivors get comb_photon_E plus an equal
// share of their cell's energy defect, so every cell's census energy is conserved exactly (:86-92).
//
// What makes this parallel without changing a bit of the result:
//   * the generator is counter based, so the i-th photon's draw is Threefry(counter = i) -- no serial RNG state;
//   * `count == 1` can only be met by the LAST photon of a cell, and only if every earlier photon of that cell was
//     killed (the count drops by one per kill and a kept photon leaves it unchanged), so the rule is "if nothing else
//     of the cell survives, its last photon does" -- a per-cell reduction;
//   * the per-cell sums the reference accumulates in list order (cell energy :69, corrected energy :80) are formed by
//     one thread per cell walking the cell's photons in list order (stable sort by cell), so they round identically.
// Kernels: k_comb_draw (per photon) -> stable radix sort of photon indices by cell (cub) -> k_seg_bounds ->
// k_comb_cells (per cell) -> scan of the keep flags -> k_comb_gather (stable compaction into the other photon buffer).
#pragma once
#include "common.cuh"
#include "rng.cuh"

namespace bg {

// keep_nat[i] = 1 if the i-th census photon survives its own draw (:74-78 without the last-photon rule)
__global__ void k_comb_draw(const double2 *__restrict__ ee, const ulonglong2 *__restrict__ sg, uint64_t n,
                            double comb_photon_E, uint64_t ctr_hi, uint64_t stream, uint32_t *__restrict__ cell_key,
                            uint32_t *__restrict__ index, uint32_t *__restrict__ keep) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double E = ee[i].x;
  const double p_kill = 1.0 - E / comb_photon_E;
  const double rand_check = u01_from_bits(threefry2x64_20_w0(i, ctr_hi, stream));
  keep[i] = (rand_check > p_kill) ? 1u : 0u;
  cell_key[i] = (uint32_t)sg[i].y;
  index[i] = (uint32_t)i;
}

// One thread per cell: its photons in list order are order[seg_start .. seg_end).
__global__ void k_comb_cells(uint32_t n_cells, const uint64_t *__restrict__ seg_start,
                             const uint64_t *__restrict__ seg_end, const uint32_t *__restrict__ order,
                             const double2 *__restrict__ ee, double comb_photon_E, uint32_t *__restrict__ keep,
                             double *__restrict__ new_E) {
  const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  const uint64_t b = seg_start[cell], e = seg_end[cell];
  if (e <= b) return;
  double cell_E = 0.0, corrected = 0.0;
  uint32_t kept = 0;
  for (uint64_t t = b; t < e; ++t) {
    const uint32_t i = order[t];
    cell_E += ee[i].x;  // (:69) in list order
    uint32_t k = keep[i];
    if (t == e - 1 && kept == 0 && !k) {  // the cell's last photon and nothing kept so far: count == 1 (:78)
      k = 1;
      keep[i] = 1;
    }
    if (k) {
      corrected += comb_photon_E;  // (:80) accumulated once per survivor, like the reference
      ++kept;
    }
  }
  new_E[cell] = comb_photon_E + (cell_E - corrected) / (double)kept;  // (:88-91)
}

// stable compaction of the survivors, energy replaced by their cell's corrected value
__global__ void k_comb_gather(PhotonSoA src, uint64_t n, const uint32_t *__restrict__ keep,
                              const uint64_t *__restrict__ offset, const double *__restrict__ new_E, PhotonSoA dst) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  const uint64_t o = offset[i];
  const ulonglong2 sg = src.sg[i];
  dst.xy[o] = src.xy[i];
  dst.za[o] = src.za[i];
  dst.bc[o] = src.bc[i];
  dst.ee[o] = make_double2(new_E[(uint32_t)sg.y], src.ee[i].y);
  dst.lc[o] = src.lc[i];
  dst.sg[o] = sg;
}

}  // namespace bg
