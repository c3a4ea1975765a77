// common.cuh -- shared types of the B200 IMC hot path (device side).
//
// Data layout in HBM:
//   * photons: structure of arrays made of six 16-byte streams so that every warp access is a coalesced 128-bit
//     load/store (reference layout: 120-byte AoS Photon, src/photon.h:171-182);
//   * cells: the mesh is a tensor-product grid, so geometry is three per-axis face arrays (staged in shared
//     memory) and the per-cell HBM record shrinks to f + abs_groups[G] + sct_groups[G]
//     (reference: 192..656-byte Cell, src/cell.h:318-337);
//   * tallies: interleaved {abs_E, track_E} == the reference's Cell_Tally (src/cell_tally.h:53-54).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bg {

// reference src/constants.h:16-24
constexpr double K_PI = 3.1415926535897932384626433832795;
constexpr double K_C = 299.792458;
constexpr double K_CUTOFF = 0.01;

enum : int { BC_REFLECT = 0, BC_VACUUM = 1, BC_ELEMENT = 2, BC_SOURCE = 3, BC_PROCESSOR = 4 };
enum : uint8_t { EV_EXIT = 0, EV_PASS = 1, EV_CENSUS = 2, EV_SCATTER = 3, EV_KILLED = 4, EV_BOUND = 5 };

// Six 16-byte streams, `cap` entries each, carved from one allocation.
struct PhotonSoA {
  double2 *xy;     // pos.x, pos.y
  double2 *za;     // pos.z, angle.x
  double2 *bc;     // angle.y, angle.z
  double2 *ee;     // E, E0
  ulonglong2 *lc;  // life_dx (bits), RNG counter low word
  ulonglong2 *sg;  // RNG stream (key low word), cell | group << 32
  uint64_t cap;
  void *base;
};

struct MeshDev {
  uint32_t nx, ny, nz, G;
  uint32_t n_cells;
  uint32_t n_faces;     // nx+1 + ny+1 + nz+1
  const double *faces;  // x faces, y faces, z faces concatenated
  int bc[6];
};

// indices into the device statistics block (unsigned long long each)
enum : int {
  ST_EVENTS = 0, ST_SCATTERS, ST_CROSSINGS, ST_REFLECTIONS, ST_DEPOSITS, ST_LOOKUPS,
  ST_N_CENSUS, ST_N_KILLED, ST_N_EXIT, ST_BAD_RNG, ST_COUNT = 16
};

}  // namespace bg
