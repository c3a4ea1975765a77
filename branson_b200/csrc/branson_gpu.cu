// branson_gpu.cu -- context + C ABI (include/branson_gpu.h) of the B200 IMC hot path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (see csrc/Makefile).  -fmad=false keeps the
// reference-stated arithmetic un-contracted, like the reference's own CPU build (src/CMakeLists.txt:93 has no -march,
// so gcc emits no FMA): pos += angle*d, (1-f)*sigma_a + sigma_s, ... round exactly as on the host.
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/branson_gpu.h"
#include "census.cuh"
#include "comb.cuh"
#include "comm_native.cuh"
#include "common.cuh"
#include "event.cuh"
#include "mesh_dev.cuh"
#include "pool.cuh"
#include "source.cuh"
#include "transport.cuh"

using namespace bg;

namespace {

std::string g_create_error;

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
};

}  // namespace

constexpr int WORK_COUNTERS = 64;  // [0]: every ordinary launch; [j]: slice j of the pipelined AoS drop-in

struct bgpu_ctx {
  int device = 0;
  int n_sm = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the pipelined AoS drop-in (created on first use)
  cudaStream_t s_k2 = nullptr;                   // its second compute stream
  cudaEvent_t ev[6] = {};
  std::string err;

  MeshDev mesh{};
  double *d_faces = nullptr;
  uint32_t seed = 0;
  uint64_t n_user = 0;
  int rank = 0, n_ranks = 1;
  uint64_t ctr_hi = 0;

  // cell data
  double *d_f = nullptr, *d_opa = nullptr, *d_ops = nullptr;  // f [n_cells]; opa/ops [n_cells*G]
  double *d_cellrec = nullptr;  // {f, sigma_a, sigma_s, 0} per cell (32 B): what a visit needs where groups are uniform
  double *d_cell_stage = nullptr;                               // 3*n_cells staging (gray values / E arrays)
  bool have_cell_data = false;
  bool closed_form_walk = true;
  bool uniform_groups = false;  // all groups of every cell equal (always true for bgpu_set_cell_data)

  // photons
  PhotonSoA work{}, census{};
  PhotonSoA comb_scratch{};  // target of the comb's compaction, swapped with `census` afterwards
  uint64_t n_work = 0, n_new = 0, n_census = 0;
  uint8_t *d_desc = nullptr;
  uint64_t desc_cap = 0;
  uint32_t *d_counters = nullptr;
  uint64_t counters_cap = 0;
  bool counters_on = false;

  // tallies: interleaved {abs_E, track_E}[n_cells] + extra doubles for the packed all-reduce
  double *d_tally = nullptr;
  uint64_t tally_extra = 0;

  unsigned long long *d_stats = nullptr;  // ST_COUNT counters
  unsigned long long *d_work_counter = nullptr;
  double *d_results = nullptr;  // [0] census_E [1] exit_E

  // scratch (grown on demand)
  DevBuf scr_counts, scr_offsets, scr_tile_sum, scr_tile_off, scr_tiles, scr_ndep, scr_dep_off, scr_dep_cell,
      scr_dep_val, scr_sort, scr_keys_out, scr_vals_in, scr_vals_out, scr_seg, scr_aos, scr_event, scr_tally_rep, scr_comb,
      scr_src_win;
  void *h_pinned = nullptr;
  size_t h_pinned_bytes = 0;
  void *h_stage = nullptr;  // pinned staging buffers of the AoS drop-in's copy threads
  size_t h_stage_bytes = 0;

  // launch config
  int block_threads = 128;
  int blocks_per_sm = 0;  // 0: occupancy query
  uint32_t chunk = 64;   // (v16 sweep: 64 is 0.5 % ahead of 128 on the hohlraum, 256 is 3 % behind; profiles/sweep_r01_v15.txt)
  bool chunk_auto = true;      // shrink the chunk when the work list is too short to give every warp several
  uint32_t scatter_batch = 12;  // history kernel: parked scatters a warp waits for before sampling them together
  bool scatter_batch_auto = true;   // ... 6 instead, when the previous launch's histories were short (make_params)
  double prev_events_per_history = 0.0;  // of the previous transport launch (0: none yet)
  int aggregate = 0;            // history kernel: combine same-cell deposits of a warp trip (off: see contended_mesh)
  int tally_copies = 0;         // replicated tallies of the history kernel (0: auto from the mesh size, 1: off)
  uint32_t tally_copies_live = 0;  // copies the zeroed scr_tally_rep currently holds (+1 for the main tally)
  uint64_t event_tail = 0;   // active-list size below which BGPU_EVENT hands over to the history kernel (0: auto)
  uint32_t event_passes = 0;
  // BGPU_EVENT: 0 = event queues in shared memory (pool.cuh, default), 1 = lockstep passes through HBM (event.cuh)
  int event_hbm = 0;
  // (swept with the v21 kernel, profiles/pool_thresholds_r02.txt: 20 / 8 is 2 % ahead of the first version's 24 / 16)
  uint32_t pool_batch_scatter = 20, pool_batch_refill = 8;
  // BGPU_HISTORY: which kernel runs the histories (per-photon results are identical): 0 = auto -- the event-queue kernel
  // on decks that MIX event types (previous launch: >= 16 events per history, 8..45 % of them scatters: big_cube,
  // hot_zone), the history kernel elsewhere; 1 = always the history kernel; 2 = always the event queues
  int kernel_choice = 0;
  double prev_scatter_fraction = 0.0;
  uint32_t kernel_used = 0;  // of the last transport: 0 history, 1 event queues, 2 event passes through HBM

  // device-resident mesh physics (bgpu_mesh_*, mesh_dev.cuh); allocated by bgpu_mesh_init
  bool mesh_ready = false;
  RegionDev *d_regions = nullptr;
  uint32_t *d_region_of_cell = nullptr;
  double *d_mesh = nullptr;  // n_cells doubles each: T_e, T_r0, T_s, T_r, op_a, op_s, E_emission, E_source, E_census, E_emission_global
  double *d_tile_sums = nullptr, *d_mesh_sums = nullptr;
  uint32_t mesh_tiles = 0;
  double mesh_dt = 0.0;
  uint32_t mesh_step = 0;
  bool mesh_redistributed = false;

  // replicated-mode collectives (comm_native.cuh): the back end this rank's all-reduces go through, the pinned staging
  // of the rank's scalars and of the tail coming back, a device scratch for host-buffer reductions
  CommHandle comm;
  double *h_comm = nullptr;       // pinned: [0, RANK_SCALARS) this rank's row; then the reduced tail + the mesh sums
  double *d_comm_scratch = nullptr;
  static constexpr uint64_t COMM_SCRATCH_DOUBLES = 4096;

  uint64_t launches = 0;  // kernels launched through this ctx since creation
  bgpu_cycle_stats stats{};
  double pre_census_E = 0.0, new_photon_E = 0.0;
  bool stats_valid = false;
};

namespace {

int fail(bgpu_ctx *c, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  else g_create_error = buf;
  return 1;
}

#define CU(c, call)                                                                                     \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return fail((c), "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, #call); \
  } while (0)

int ensure(bgpu_ctx *c, DevBuf &b, size_t bytes) {
  if (bytes <= b.bytes) return 0;
  if (b.p) CU(c, cudaFree(b.p));
  b.p = nullptr;
  b.bytes = 0;
  size_t want = bytes + bytes / 8 + 256;
  CU(c, cudaMalloc(&b.p, want));
  b.bytes = want;
  return 0;
}

int ensure_pinned(bgpu_ctx *c, size_t bytes) {
  if (bytes <= c->h_pinned_bytes) return 0;
  if (c->h_pinned) CU(c, cudaFreeHost(c->h_pinned));
  c->h_pinned = nullptr;
  c->h_pinned_bytes = 0;
  CU(c, cudaHostAlloc(&c->h_pinned, bytes + bytes / 8 + 256, cudaHostAllocDefault));
  c->h_pinned_bytes = bytes + bytes / 8 + 256;
  return 0;
}

void carve(PhotonSoA &s, void *base, uint64_t cap) {
  char *p = (char *)base;
  s.base = base;
  s.cap = cap;
  s.xy = (double2 *)p;                 p += 16 * cap;
  s.za = (double2 *)p;                 p += 16 * cap;
  s.bc = (double2 *)p;                 p += 16 * cap;
  s.ee = (double2 *)p;                 p += 16 * cap;
  s.lc = (ulonglong2 *)p;              p += 16 * cap;
  s.sg = (ulonglong2 *)p;
}

// grow a photon list to hold at least n photons, keeping the first `keep` photons
int ensure_soa(bgpu_ctx *c, PhotonSoA &s, uint64_t n, uint64_t keep) {
  if (n <= s.cap) return 0;
  uint64_t cap = std::max<uint64_t>(n + n / 4, 1024);
  cap = (cap + 15) & ~15ull;
  void *base = nullptr;
  CU(c, cudaMalloc(&base, 96 * cap));
  PhotonSoA t{};
  carve(t, base, cap);
  if (keep && s.base) {
    CU(c, cudaMemcpyAsync(t.xy, s.xy, 16 * keep, cudaMemcpyDeviceToDevice, c->stream));
    CU(c, cudaMemcpyAsync(t.za, s.za, 16 * keep, cudaMemcpyDeviceToDevice, c->stream));
    CU(c, cudaMemcpyAsync(t.bc, s.bc, 16 * keep, cudaMemcpyDeviceToDevice, c->stream));
    CU(c, cudaMemcpyAsync(t.ee, s.ee, 16 * keep, cudaMemcpyDeviceToDevice, c->stream));
    CU(c, cudaMemcpyAsync(t.lc, s.lc, 16 * keep, cudaMemcpyDeviceToDevice, c->stream));
    CU(c, cudaMemcpyAsync(t.sg, s.sg, 16 * keep, cudaMemcpyDeviceToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  if (s.base) CU(c, cudaFree(s.base));
  s = t;
  return 0;
}

int ensure_work(bgpu_ctx *c, uint64_t n, uint64_t keep) {
  if (ensure_soa(c, c->work, n, keep)) return 1;
  if (c->desc_cap < c->work.cap) {
    if (c->d_desc) CU(c, cudaFree(c->d_desc));
    c->d_desc = nullptr;
    CU(c, cudaMalloc((void **)&c->d_desc, c->work.cap + 64));
    c->desc_cap = c->work.cap;
  }
  if (c->counters_on && c->counters_cap < c->work.cap) {
    if (c->d_counters) CU(c, cudaFree(c->d_counters));
    c->d_counters = nullptr;
    CU(c, cudaMalloc((void **)&c->d_counters, 16 * c->work.cap));
    c->counters_cap = c->work.cap;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_expand_groups(uint32_t n_cells, uint32_t G, const double *__restrict__ a,
                                const double *__restrict__ s, double *__restrict__ opa, double *__restrict__ ops) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint64_t)n_cells * G) return;
  const uint32_t cell = (uint32_t)(t / G);
  opa[t] = a[cell];
  ops[t] = s[cell];
}

// cells whose groups do not all carry the same (sigma_a, sigma_s) pair: with none of them the photon's group never enters
// the physics (closed-form group walk, no reload on a group change, lazily sampled groups: transport.cuh)
__global__ void k_count_nonuniform_cells(uint32_t n_cells, uint32_t G, const double *__restrict__ opa,
                                         const double *__restrict__ ops, unsigned long long *count) {
  const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  const double *a = opa + (uint64_t)cell * G, *sc = ops + (uint64_t)cell * G;
  bool same = true;
  for (uint32_t g = 1; g < G; ++g) same = same && (a[g] == a[0]) && (sc[g] == sc[0]);
  if (!same) atomicAdd(count, 1ull);
}

// One 32-byte record per cell for decks whose groups all carry the same opacities: a cell visit then costs one sector
// instead of three (f, abs_groups[g], sct_groups[g] live in three arrays, the latter two 8 G bytes per cell), and the
// whole table (19 MB for the hohlraum's 591 500 cells, against 289 MB) stays in L2 next to the tallies.
__global__ void k_fill_cellrec(uint32_t n_cells, uint32_t G, const double *__restrict__ f, const double *__restrict__ opa,
                               const double *__restrict__ ops, double *__restrict__ rec) {
  const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  reinterpret_cast<double2 *>(rec)[2 * (uint64_t)cell] = make_double2(f[cell], opa[(uint64_t)cell * G]);
  reinterpret_cast<double2 *>(rec)[2 * (uint64_t)cell + 1] = make_double2(ops[(uint64_t)cell * G], 0.0);
}

__global__ void k_copy_soa(PhotonSoA src, uint64_t src_off, PhotonSoA dst, uint64_t dst_off, uint64_t n) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  dst.xy[dst_off + t] = src.xy[src_off + t];
  dst.za[dst_off + t] = src.za[src_off + t];
  dst.bc[dst_off + t] = src.bc[src_off + t];
  dst.ee[dst_off + t] = src.ee[src_off + t];
  dst.lc[dst_off + t] = src.lc[src_off + t];
  dst.sg[dst_off + t] = src.sg[src_off + t];
}

// in-order (tile-tree) sum of E over a photon list: get_photon_list_E (src/census_functions.h:31-46)
__global__ void __launch_bounds__(CT_THREADS) k_list_E_tiles(const double2 *__restrict__ ee, uint64_t n,
                                                             double *__restrict__ tile_E, int use_E0) {
  __shared__ double s_red[CT_THREADS >> 5];
  const uint64_t base = (uint64_t)blockIdx.x * CT_TILE + (uint64_t)threadIdx.x * CT_ITEMS;
  double e = 0.0;
#pragma unroll
  for (int i = 0; i < CT_ITEMS; ++i)
    if (base + i < n) e += use_E0 ? ee[base + i].y : ee[base + i].x;
  const double b = block_sum(e, s_red);
  if (threadIdx.x == 0) tile_E[blockIdx.x] = b;
}
// the tile sums in a fixed order: thread t adds its contiguous run of tiles serially, the 1024 runs are then added in
// thread order by a fixed tree -- reproducible for a given n, and a few microseconds where one thread walking 8 000 tiles
// took 90
__global__ void __launch_bounds__(1024) k_sum_ordered(const double *__restrict__ v, uint32_t n, double *out) {
  __shared__ double s[1024];
  const uint32_t per = (n + 1023u) / 1024u;
  const uint32_t b = threadIdx.x * per, e = (b + per < n) ? b + per : n;
  double acc = 0.0;
  for (uint32_t i = b; i < e; ++i) acc += v[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t w = 512; w > 0; w >>= 1) {
    if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = s[0];
}

// reference AoS Photon (120 bytes, src/photon.h:171-182) <-> device SoA
__global__ void k_aos_to_soa(const uint64_t *__restrict__ aos, uint64_t n, PhotonSoA ph, uint64_t ctr_hi,
                             unsigned long long *stats) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t *w = aos + 15 * t;
  const uint64_t w0 = w[0];
  ph.xy[t] = make_double2(__longlong_as_double(w[2]), __longlong_as_double(w[3]));
  ph.za[t] = make_double2(__longlong_as_double(w[4]), __longlong_as_double(w[5]));
  ph.bc[t] = make_double2(__longlong_as_double(w[6]), __longlong_as_double(w[7]));
  ph.ee[t] = make_double2(__longlong_as_double(w[8]), __longlong_as_double(w[9]));
  ph.lc[t] = make_ulonglong2(w[10], w[11]);
  ph.sg[t] = make_ulonglong2(w[13], w0);  // cell | group << 32 is exactly word 0
  if (w[12] != ctr_hi || w[14] != 0ull) atomicAdd(&stats[ST_BAD_RNG], 1ull);
}
__global__ void k_soa_to_aos(uint64_t *__restrict__ aos, uint64_t n, PhotonSoA ph, const uint8_t *__restrict__ desc) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint64_t *w = aos + 15 * t;
  const double2 xy = ph.xy[t], za = ph.za[t], bc = ph.bc[t], ee = ph.ee[t];
  const ulonglong2 lc = ph.lc[t], sg = ph.sg[t];
  w[0] = sg.y;
  // descriptors[0] = event (Photon::set_descriptor, src/photon.h); keep source_type and the padding bytes
  w[1] = (w[1] & 0xffffff00ffffffffull) | ((uint64_t)desc[t] << 32);
  w[2] = __double_as_longlong(xy.x); w[3] = __double_as_longlong(xy.y); w[4] = __double_as_longlong(za.x);
  w[5] = __double_as_longlong(za.y); w[6] = __double_as_longlong(bc.x); w[7] = __double_as_longlong(bc.y);
  w[8] = __double_as_longlong(ee.x);
  w[10] = lc.x;
  w[11] = lc.y;
}

// deterministic tally: segment bounds of the cell-sorted deposit log, then in-order sums
__global__ void k_seg_bounds(const uint32_t *__restrict__ keys, uint64_t n, uint64_t *__restrict__ seg_start,
                             uint64_t *__restrict__ seg_end) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t k = keys[t];
  if (t == 0 || keys[t - 1] != k) seg_start[k] = t;
  if (t == n - 1 || keys[t + 1] != k) seg_end[k] = t + 1;
}
__global__ void k_seg_sum(uint32_t n_cells, const uint64_t *__restrict__ seg_start,
                          const uint64_t *__restrict__ seg_end, const uint32_t *__restrict__ order,
                          const double2 *__restrict__ dep_val, double2 *__restrict__ tally) {
  const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  const uint64_t b = seg_start[cell], e = seg_end[cell];
  if (e <= b) return;
  double2 acc = tally[cell];
  for (uint64_t p = b; p < e; ++p) {
    const double2 v = dep_val[order[p]];
    acc.x += v.x;
    acc.y += v.y;
  }
  tally[cell] = acc;
}
__global__ void k_iota(uint32_t *v, uint64_t n) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) v[t] = (uint32_t)t;
}

inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// exclusive scan of n u32 -> n+1 u64 on the ctx stream
int device_scan(bgpu_ctx *c, const uint32_t *in, uint64_t n, uint64_t *out) {
  if (n == 0) {
    CU(c, cudaMemsetAsync(out, 0, 8, c->stream));
    return 0;
  }
  const uint32_t tiles = (uint32_t)((n + CT_TILE - 1) / CT_TILE);
  if (ensure(c, c->scr_tile_sum, 4ull * tiles)) return 1;
  if (ensure(c, c->scr_tile_off, 8ull * (tiles + 1))) return 1;
  ++c->launches;
  k_scan_tile_sums<<<tiles, CT_THREADS, 0, c->stream>>>(in, n, (uint32_t *)c->scr_tile_sum.p);
  ++c->launches;
  k_scan_single<<<1, 1024, 0, c->stream>>>((const uint32_t *)c->scr_tile_sum.p, tiles, (uint64_t *)c->scr_tile_off.p);
  ++c->launches;
  k_scan_apply<<<tiles, CT_THREADS, 0, c->stream>>>(in, n, (const uint64_t *)c->scr_tile_off.p, out);
  CU(c, cudaGetLastError());
  return 0;
}

template <int MODE, bool RESUME = false>
int launch_history(bgpu_ctx *c, const TransportParams &P, cudaStream_t stream = nullptr) {
  if (!stream) stream = c->stream;
  const size_t smem = (size_t)P.mesh.n_faces * 8;
  const bool use_smem = smem <= 160 * 1024;
  const bool ctrs = P.counters != nullptr;
  // PACKED: the cell's record instead of the three cell-data arrays (transport.cuh enter_cell); the RESUME launches of the
  // event variant keep the arrays (they share event.cuh's device functions)
  const bool packed = !RESUME && P.uniform_groups != 0;
  void (*kern)(const TransportParams) = nullptr;
  if (packed) {
    if (use_smem) kern = ctrs ? k_transport_history<MODE, true, true, false, true> : k_transport_history<MODE, false, true, false, true>;
    else kern = ctrs ? k_transport_history<MODE, true, false, false, true> : k_transport_history<MODE, false, false, false, true>;
  } else {
    if (use_smem) kern = ctrs ? k_transport_history<MODE, true, true, RESUME, false> : k_transport_history<MODE, false, true, RESUME, false>;
    else kern = ctrs ? k_transport_history<MODE, true, false, RESUME, false> : k_transport_history<MODE, false, false, RESUME, false>;
  }
  if (use_smem && smem > 48 * 1024)
    CU(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = c->blocks_per_sm;
  if (per_sm <= 0) {
    CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, c->block_threads, use_smem ? smem : 0));
    if (per_sm < 1) per_sm = 1;
  }
  uint64_t blocks = (uint64_t)c->n_sm * per_sm;
  const uint64_t max_useful = (P.n + c->block_threads - 1) / c->block_threads;
  if (blocks > max_useful) blocks = std::max<uint64_t>(max_useful, 1);
  ++c->launches;
  kern<<<(unsigned)blocks, c->block_threads, use_smem ? smem : 0, stream>>>(P);
  CU(c, cudaGetLastError());
  return 0;
}

// BGPU_EVENT: lockstep passes over two active lists (event.cuh), tail finished by the history kernel in RESUME mode
int run_event(bgpu_ctx *c, TransportParams P) {
  const uint64_t n = P.n;
  const size_t bytes = 16 * n + 16 * n + 4 * n + 4 * 4 * n + 64;
  if (ensure(c, c->scr_event, bytes)) return 1;
  char *p = (char *)c->scr_event.p;
  double2 *acc = (double2 *)p;                 p += 16 * n;
  uint4 *cnt = (uint4 *)p;                     p += 16 * n;
  uint32_t *lk = (uint32_t *)p;                p += 4 * n;
  uint32_t *lists[2][2];                       // [ping-pong][0 scatter, 1 continue]
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) { lists[a][b] = (uint32_t *)p; p += 4 * n; }
  unsigned long long *n_out = (unsigned long long *)(((uintptr_t)p + 15) & ~(uintptr_t)15);
  const size_t smem = (size_t)P.mesh.n_faces * 8;
  const bool use_smem = smem <= 160 * 1024;
  const bool ctrs = P.counters != nullptr;
  void (*kern)(const EventParams) = nullptr;
  if (use_smem) kern = ctrs ? k_event_pass<true, true> : k_event_pass<false, true>;
  else kern = ctrs ? k_event_pass<true, false> : k_event_pass<false, false>;
  if (use_smem && smem > 48 * 1024)
    CU(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, use_smem ? smem : 0));
  if (per_sm < 1) per_sm = 1;
  const uint64_t machine = (uint64_t)c->n_sm * per_sm * 128;  // lanes resident at once
  const uint64_t tail = c->event_tail ? c->event_tail : 2 * machine;
  EventParams E{};
  E.T = P;
  E.acc = acc; E.cnt = cnt; E.lk = lk;
  E.n_out = n_out;
  uint64_t n_in[2] = {0, n};  // first pass: everything is on the "continue" list (identity order)
  int cur = 0;
  bool first = true;
  c->event_passes = 0;
  while (n_in[0] + n_in[1] > 0) {
    if (!first && n_in[0] + n_in[1] <= tail) {
      for (int kind = 0; kind < 2; ++kind) {
        if (!n_in[kind]) continue;
        TransportParams R = P;
        R.n = n_in[kind];
        R.index_list = lists[cur][kind];
        R.carry_acc = acc; R.carry_cnt = cnt; R.carry_lk = lk;
        R.resume_pending_scatter = kind == 0 ? 1 : 0;
        CU(c, cudaMemsetAsync(c->d_work_counter, 0, 8, c->stream));
        if (launch_history<TM_ATOMIC, true>(c, R)) return 1;
      }
      break;
    }
    CU(c, cudaMemsetAsync(n_out, 0, 16, c->stream));
    for (int kind = 0; kind < 2; ++kind) {
      if (!n_in[kind]) continue;
      E.list_in = first ? nullptr : lists[cur][kind];
      E.n_in = n_in[kind];
      E.scatter_out = lists[cur ^ 1][0];
      E.cont_out = lists[cur ^ 1][1];
      E.first = first ? 1 : 0;
      E.pending_scatter = (!first && kind == 0) ? 1 : 0;
      const uint64_t blocks = std::min<uint64_t>((n_in[kind] + 127) / 128, (uint64_t)c->n_sm * per_sm * 4);
      ++c->launches;
      kern<<<(unsigned)blocks, 128, use_smem ? smem : 0, c->stream>>>(E);
      CU(c, cudaGetLastError());
    }
    unsigned long long h[2] = {0, 0};
    CU(c, cudaMemcpyAsync(h, n_out, 16, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    n_in[0] = h[0];
    n_in[1] = h[1];
    cur ^= 1;
    first = false;
    ++c->event_passes;
  }
  return 0;
}

// the slot record of the queue kernel holds i and j in 16 bits each and k in 24 (pool.cuh, PV_SK)
bool pool_mesh_ok(const bgpu_ctx *c) { return c->mesh.nx < 65536u && c->mesh.ny < 65536u && c->mesh.nz < (1u << 24); }

// BGPU_EVENT, default form: event queues in shared memory (pool.cuh); tallies and statistics as in the history launch
int launch_pool(bgpu_ctx *c, const TransportParams &P) {
  const size_t face_bytes = (size_t)P.mesh.n_faces * 8;
  const bool use_smem = face_bytes <= 100 * 1024;
  const size_t smem = 4 * POOL_BYTES_PER_WARP + (use_smem ? face_bytes : 0);
  const bool ctrs = P.counters != nullptr;
  const bool packed = P.uniform_groups != 0;
  PoolParams Q{};
  Q.T = P;
  Q.batch_scatter = c->pool_batch_scatter;
  Q.batch_refill = c->pool_batch_refill;
  void (*kern)(const PoolParams) = nullptr;
  if (packed) {
    if (use_smem) kern = ctrs ? k_transport_pool<true, true, true> : k_transport_pool<false, true, true>;
    else kern = ctrs ? k_transport_pool<true, false, true> : k_transport_pool<false, false, true>;
  } else {
    if (use_smem) kern = ctrs ? k_transport_pool<true, true, false> : k_transport_pool<false, true, false>;
    else kern = ctrs ? k_transport_pool<true, false, false> : k_transport_pool<false, false, false>;
  }
  CU(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
  if (per_sm < 1) return fail(c, "event-queue kernel: %zu bytes of shared memory per CTA do not fit", smem);
  if (c->blocks_per_sm > 0 && c->blocks_per_sm < per_sm) per_sm = c->blocks_per_sm;
  uint64_t blocks = (uint64_t)c->n_sm * per_sm;
  const uint64_t max_useful = (P.n + 128 * POOL_ROWS - 1) / (128 * POOL_ROWS);
  if (blocks > max_useful) blocks = std::max<uint64_t>(max_useful, 1);
  ++c->launches;
  kern<<<(unsigned)blocks, 128, smem, c->stream>>>(Q);
  CU(c, cudaGetLastError());
  return 0;
}

// Replicated tallies (transport.cuh, TransportParams::tally_rep): as many copies as fit 64 MB (L2-sized), at most 64;
// big meshes (8e6 cells) get none -- their deposits are spread over so many addresses that nothing serialises.  The
// copies are zero between launches (k_fold_tally re-zeroes them).
// Tally contention: where many photons deposit into few cells -- marshak's 25 cells, hot_zone's 5 x 5 hot corner of
// 40 000 -- same-address FP64 reductions serialise in L2, and replicated tallies (up to 64 copies, folded after the
// launch) are worth 3x.  On the 591 500-cell hohlraum and the 8e6-cell cube they only cost L2 space and a fold: they are
// on for meshes below 2^17 cells unless the host says otherwise (bgpu_set_tally_copies).  The second measure, combining
// a warp trip's same-cell deposits by match_any + a shuffle tree (bgpu_set_divergence), is off by default: with the
// replicated tallies in place its matching costs more than the atomics it saves, on every deck (hot_zone +6...10 %,
// marshak +7 %, big_cube +7 %, hohlraum streaming cycle +9 %, multi-node share +12 % without it;
// profiles/sweep_r01_v15.txt).
bool contended_mesh(const bgpu_ctx *c) { return c->mesh.n_cells < (1u << 17); }

int prepare_tally_copies(bgpu_ctx *c) {
  uint32_t copies = c->tally_copies > 0 ? (uint32_t)c->tally_copies
                    : !contended_mesh(c) ? 1u
                                         : (uint32_t)std::min<uint64_t>(64, (64ull << 20) / (16ull * c->mesh.n_cells));
  if (copies < 1) copies = 1;
  if (c->tally_copies_live == copies) return 0;
  if (copies > 1) {
    const size_t bytes = 16ull * c->mesh.n_cells * (copies - 1);
    if (ensure(c, c->scr_tally_rep, bytes)) return 1;
    CU(c, cudaMemsetAsync(c->scr_tally_rep.p, 0, bytes, c->stream));
  }
  c->tally_copies_live = copies;
  return 0;
}

// photons a warp takes from the work list at a time
uint32_t chunk_for(const bgpu_ctx *c, uint64_t n) {
  if (!c->chunk_auto) return c->chunk;
  // at least ~8 chunks per resident warp, so the last chunks do not leave most of the machine idle (marshak /
  // hot_zone: 1e6 photons over 2960 warps)
  const uint64_t warps = (uint64_t)c->n_sm * 5 * 4;
  uint64_t ch = n / (warps * 8);
  ch = std::max<uint64_t>(32, std::min<uint64_t>(c->chunk, ch & ~31ull));
  return (uint32_t)ch;
}

TransportParams make_params(bgpu_ctx *c, bool writeback_all) {
  TransportParams P{};
  P.ph = c->work;
  P.n = c->n_work;
  P.desc = c->d_desc;
  P.counters = c->counters_on ? c->d_counters : nullptr;
  P.mesh = c->mesh;
  P.f = c->d_f;
  P.opa = c->d_opa;
  P.ops = c->d_ops;
  P.cellrec = reinterpret_cast<const double2 *>(c->d_cellrec);
  P.tally = (double2 *)c->d_tally;
  P.ctr_hi = c->ctr_hi;
  P.work_counter = c->d_work_counter;
  P.chunk = chunk_for(c, c->n_work);
  P.inv_sxy = 1.0 / ((double)c->mesh.nx * (double)c->mesh.ny);
  P.inv_nx = 1.0 / (double)c->mesh.nx;
  P.scatter_batch = c->scatter_batch;
  // In a history of a handful of events a lane parked at a scatter idles for a large part of it while the warp collects
  // twelve: the short-history multi-node hohlraum (4.8 events per history) runs 3.5 % faster at 4-8, the decks with
  // 30+ events per history want 8-16 (profiles/sweep_r01_v15.txt).  The previous cycle's event count decides.
  if (c->scatter_batch_auto && c->prev_events_per_history > 0.0 && c->prev_events_per_history < 16.0) P.scatter_batch = 6;
  P.aggregate = c->aggregate;
  P.writeback_all = writeback_all ? 1 : 0;
  P.stats = c->d_stats;
  P.uniform_groups = (c->uniform_groups && c->closed_form_walk) ? 1 : 0;
  return P;
}

// the history loop over the device work list; tallies are ACCUMULATED into d_tally
int run_transport(bgpu_ctx *c, int algorithm, int tally_mode, bool writeback_all) {
  if (!c->have_cell_data) return fail(c, "bgpu_transport: cell data not set (call bgpu_set_cell_data first)");
  if (algorithm != BGPU_HISTORY && algorithm != BGPU_EVENT) return fail(c, "unknown transport algorithm %d", algorithm);
  if (tally_mode != BGPU_TALLY_ATOMIC && tally_mode != BGPU_TALLY_DETERMINISTIC)
    return fail(c, "unknown tally mode %d", tally_mode);
  CU(c, cudaMemsetAsync(c->d_stats, 0, 8 * ST_COUNT, c->stream));
  if (c->n_work == 0) return 0;
  if (c->n_work >= (1ull << 32)) return fail(c, "bgpu_transport: %llu photons in one work list (limit 2^32 - 1)",
                                             (unsigned long long)c->n_work);
  TransportParams P = make_params(c, writeback_all);
  c->kernel_used = 0u;
  if (algorithm == BGPU_EVENT) {
    if (tally_mode != BGPU_TALLY_ATOMIC) return fail(c, "the event-based variant supports BGPU_TALLY_ATOMIC only");
    // (the queue kernel packs a slot's i | j << 16 and k | descriptor << 24: meshes beyond that take the HBM-pass form)
    const bool hbm = c->event_hbm || !pool_mesh_ok(c);
    c->kernel_used = hbm ? 2u : 1u;
    if (hbm) return run_event(c, P);
    CU(c, cudaMemsetAsync(c->d_work_counter, 0, 8, c->stream));
    if (prepare_tally_copies(c)) return 1;
    const uint32_t copies = c->tally_copies_live;
    if (copies > 1) {
      P.tally_rep = (double2 *)c->scr_tally_rep.p;
      P.tally_copies = copies;
    }
    if (launch_pool(c, P)) return 1;
    if (copies > 1) {
      ++c->launches;
      k_fold_tally<<<grid_for(c->mesh.n_cells, 256), 256, 0, c->stream>>>((double2 *)c->d_tally, P.tally_rep,
                                                                            c->mesh.n_cells, copies - 1);
      CU(c, cudaGetLastError());
    }
    return 0;
  }
  if (tally_mode == BGPU_TALLY_ATOMIC) {
    CU(c, cudaMemsetAsync(c->d_work_counter, 0, 8, c->stream));
    if (prepare_tally_copies(c)) return 1;
    const uint32_t copies = c->tally_copies_live;
    if (copies > 1) {
      P.tally_rep = (double2 *)c->scr_tally_rep.p;
      P.tally_copies = copies;
    }
    // Mixed decks lose a third of their lanes to divergence in the history kernel; the event-queue kernel regroups them
    // (pool.cuh; big_cube 22 -> 28 lanes per instruction, +6 % histories/s; profiles/pool_variants_r02.txt).  It loses on
    // scattering-dominated decks (every lane scatters on every trip anyway) and on short histories.
    // (hot_zone at 1e6 photons: 767 against 718 M histories/s; at 1e7: 882 against 837; a list that does not even fill
    // the resident slots a few times is left to the history kernel)
    const bool mixed = c->prev_events_per_history >= 16.0 && c->prev_scatter_fraction >= 0.08 &&
                       c->prev_scatter_fraction <= 0.45 && c->n_work >= 500000ull;
    const bool queues = pool_mesh_ok(c) && (c->kernel_choice == 2 || (c->kernel_choice == 0 && mixed));
    c->kernel_used = queues ? 1u : 0u;
    if (queues ? launch_pool(c, P) : launch_history<TM_ATOMIC>(c, P)) return 1;
    if (copies > 1) {
      ++c->launches;
      k_fold_tally<<<grid_for(c->mesh.n_cells, 256), 256, 0, c->stream>>>((double2 *)c->d_tally, P.tally_rep,
                                                                            c->mesh.n_cells, copies - 1);
      CU(c, cudaGetLastError());
    }
    return 0;
  }
  // ---- deterministic: count deposits, scan, log, stable sort by cell, in-order segment sums ----
  const uint64_t n = c->n_work;
  if (ensure(c, c->scr_ndep, 4 * n)) return 1;
  if (ensure(c, c->scr_dep_off, 8 * (n + 1))) return 1;
  P.ndep = (uint32_t *)c->scr_ndep.p;
  CU(c, cudaMemsetAsync(c->d_work_counter, 0, 8, c->stream));
  if (launch_history<TM_COUNT>(c, P)) return 1;
  if (device_scan(c, P.ndep, n, (uint64_t *)c->scr_dep_off.p)) return 1;
  uint64_t n_dep = 0;
  CU(c, cudaMemcpyAsync(&n_dep, (uint64_t *)c->scr_dep_off.p + n, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (n_dep >= (1ull << 32)) return fail(c, "deterministic tally mode: %llu deposits exceed 2^32", (unsigned long long)n_dep);
  if (ensure(c, c->scr_dep_cell, 4 * n_dep)) return 1;
  if (ensure(c, c->scr_dep_val, 16 * n_dep)) return 1;
  P.dep_off = (const uint64_t *)c->scr_dep_off.p;
  P.dep_cell = (uint32_t *)c->scr_dep_cell.p;
  P.dep_val = (double2 *)c->scr_dep_val.p;
  CU(c, cudaMemsetAsync(c->d_work_counter, 0, 8, c->stream));
  CU(c, cudaMemsetAsync(c->d_stats, 0, 8 * ST_COUNT, c->stream));
  if (launch_history<TM_LOG>(c, P)) return 1;
  if (n_dep == 0) return 0;
  if (ensure(c, c->scr_keys_out, 4 * n_dep)) return 1;
  if (ensure(c, c->scr_vals_in, 4 * n_dep)) return 1;
  if (ensure(c, c->scr_vals_out, 4 * n_dep)) return 1;
  ++c->launches;
  k_iota<<<grid_for(n_dep, 256), 256, 0, c->stream>>>((uint32_t *)c->scr_vals_in.p, n_dep);
  int end_bit = 1;
  while ((1ull << end_bit) < c->mesh.n_cells) ++end_bit;
  size_t tmp_bytes = 0;
  CU(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint32_t *)c->scr_dep_cell.p,
                                        (uint32_t *)c->scr_keys_out.p, (const uint32_t *)c->scr_vals_in.p,
                                        (uint32_t *)c->scr_vals_out.p, (uint64_t)n_dep, 0, end_bit, c->stream));
  if (ensure(c, c->scr_sort, tmp_bytes)) return 1;
  CU(c, cub::DeviceRadixSort::SortPairs(c->scr_sort.p, tmp_bytes, (const uint32_t *)c->scr_dep_cell.p,
                                        (uint32_t *)c->scr_keys_out.p, (const uint32_t *)c->scr_vals_in.p,
                                        (uint32_t *)c->scr_vals_out.p, (uint64_t)n_dep, 0, end_bit, c->stream));
  if (ensure(c, c->scr_seg, 16ull * c->mesh.n_cells)) return 1;
  CU(c, cudaMemsetAsync(c->scr_seg.p, 0, 16ull * c->mesh.n_cells, c->stream));
  uint64_t *seg_start = (uint64_t *)c->scr_seg.p, *seg_end = seg_start + c->mesh.n_cells;
  ++c->launches;
  k_seg_bounds<<<grid_for(n_dep, 256), 256, 0, c->stream>>>((const uint32_t *)c->scr_keys_out.p, n_dep, seg_start,
                                                             seg_end);
  ++c->launches;
  k_seg_sum<<<grid_for(c->mesh.n_cells, 128), 128, 0, c->stream>>>(c->mesh.n_cells, seg_start, seg_end,
                                                                   (const uint32_t *)c->scr_vals_out.p,
                                                                   (const double2 *)c->scr_dep_val.p,
                                                                   (double2 *)c->d_tally);
  CU(c, cudaGetLastError());
  return 0;
}

// post_process_photons: compaction of CENSUS photons of the work list into the census list
int run_census(bgpu_ctx *c, double next_dt, bool serial_sums) {
  const uint64_t n = c->n_work;
  c->n_census = 0;
  double res[2] = {0.0, 0.0};
  unsigned long long st[ST_COUNT] = {};
  if (n) {
    const uint32_t tiles = (uint32_t)((n + CT_TILE - 1) / CT_TILE);
    const size_t per = 4ull * 3 + 8ull * 2;
    if (ensure(c, c->scr_tiles, per * tiles + 8ull * (tiles + 1) + 64)) return 1;
    TilePartials T;
    char *p = (char *)c->scr_tiles.p;
    T.census_E = (double *)p;    p += 8ull * tiles;
    T.exit_E = (double *)p;      p += 8ull * tiles;
    T.tile_off = (uint64_t *)p;  p += 8ull * (tiles + 1);
    T.n_census = (uint32_t *)p;  p += 4ull * tiles;
    T.n_killed = (uint32_t *)p;  p += 4ull * tiles;
    T.n_exit = (uint32_t *)p;
    ++c->launches;
    k_census_tiles<<<tiles, CT_THREADS, 0, c->stream>>>(c->d_desc, c->work.ee, n, T);
    ++c->launches;
    k_scan_partials<<<1, 1024, 0, c->stream>>>(tiles, T, c->d_results, c->d_stats);
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(st, c->d_stats, 8 * ST_COUNT, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(res, c->d_results, 16, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    const uint64_t nc = st[ST_N_CENSUS];
    if (ensure_soa(c, c->census, nc, 0)) return 1;
    if (nc) {
      ++c->launches;
      k_census_scatter<<<tiles, CT_THREADS, 0, c->stream>>>(c->d_desc, c->work, n, c->census, 0, T.tile_off,
                                                            K_C * next_dt);
      CU(c, cudaGetLastError());
    }
    c->n_census = nc;
    if (serial_sums) {
      // the reference's strictly serial photon-order sums (src/post_process_functions.h:41-55), on the host
      if (ensure_pinned(c, 17 * n)) return 1;
      double2 *h_ee = (double2 *)c->h_pinned;
      uint8_t *h_d = (uint8_t *)c->h_pinned + 16 * n;
      CU(c, cudaMemcpyAsync(h_ee, c->work.ee, 16 * n, cudaMemcpyDeviceToHost, c->stream));
      CU(c, cudaMemcpyAsync(h_d, c->d_desc, n, cudaMemcpyDeviceToHost, c->stream));
      CU(c, cudaStreamSynchronize(c->stream));
      double ce = 0.0, xe = 0.0;
      for (uint64_t i = 0; i < n; ++i) {
        if (h_d[i] == EV_CENSUS) ce += h_ee[i].x;
        else if (h_d[i] == EV_EXIT) xe += h_ee[i].x;
      }
      res[0] = ce;
      res[1] = xe;
    }
  }
  bgpu_cycle_stats &s = c->stats;
  s.census_E = res[0];
  s.exit_E = res[1];
  s.n_census = c->n_census;
  s.n_killed = st[ST_N_KILLED];
  s.n_exit = st[ST_N_EXIT];
  s.n_events = st[ST_EVENTS];
  s.n_scatters = st[ST_SCATTERS];
  s.n_crossings = st[ST_CROSSINGS];
  s.n_reflections = st[ST_REFLECTIONS];
  s.n_deposits = st[ST_DEPOSITS];
  s.n_group_lookups = st[ST_LOOKUPS];
  s.n_launches = c->launches;
  if (s.n_transported) c->prev_events_per_history = (double)s.n_events / (double)s.n_transported;
  if (s.n_events) c->prev_scatter_fraction = (double)s.n_scatters / (double)s.n_events;
  s.transport_kernel = c->kernel_used;
  return 0;
}

// sum of E (or E0) over a photon list into d_results[slot], stream-ordered (no host synchronisation)
int list_energy_async(bgpu_ctx *c, const PhotonSoA &list, uint64_t off, uint64_t n, int slot, int use_E0 = 0) {
  if (!n) {
    CU(c, cudaMemsetAsync(c->d_results + slot, 0, 8, c->stream));
    return 0;
  }
  const uint32_t tiles = (uint32_t)((n + CT_TILE - 1) / CT_TILE);
  if (ensure(c, c->scr_tile_sum, 8ull * tiles + 8)) return 1;
  double *tile_E = (double *)c->scr_tile_sum.p;
  ++c->launches;
  k_list_E_tiles<<<tiles, CT_THREADS, 0, c->stream>>>(list.ee + off, n, tile_E, use_E0);
  ++c->launches;
  k_sum_ordered<<<1, 1024, 0, c->stream>>>(tile_E, tiles, c->d_results + slot);
  CU(c, cudaGetLastError());
  return 0;
}

int list_energy(bgpu_ctx *c, const PhotonSoA &list, uint64_t off, uint64_t n, double *out, int use_E0 = 0) {
  *out = 0.0;
  if (!n) return 0;
  if (list_energy_async(c, list, off, n, 2, use_E0)) return 1;
  CU(c, cudaMemcpyAsync(out, c->d_results + 2, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return 0;
}

}  // namespace

// =================================================================================================================
// C ABI
// =================================================================================================================
extern "C" {

int bgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char *bgpu_last_error(const bgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int bgpu_create(bgpu_ctx **out, const bgpu_mesh_desc *d) {
  if (!out || !d) return fail(nullptr, "bgpu_create: null argument");
  *out = nullptr;
  if (d->abi_version != BGPU_ABI_VERSION) return fail(nullptr, "bgpu_create: ABI version %u != %u", d->abi_version, BGPU_ABI_VERSION);
  if (!d->nx || !d->ny || !d->nz || !d->n_groups || !d->x_faces || !d->y_faces || !d->z_faces)
    return fail(nullptr, "bgpu_create: empty mesh description");
  if ((uint64_t)d->nx * d->ny * d->nz >= (1ull << 32)) return fail(nullptr, "bgpu_create: more than 2^32 cells");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(nullptr, "bgpu_create: no CUDA device (%s); this library has no CPU fallback",
                e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
  bgpu_ctx *c = new bgpu_ctx();
  // rank -> device map of the reference (src/gpu_setup.h:68-78)
  int dev = d->device;
  if (dev < 0) dev = (d->n_ranks <= n_dev) ? d->rank : d->rank % n_dev;
  if (dev >= n_dev) dev = dev % n_dev;
  c->device = dev;
#define CUC(call)                                                                                              \
  do {                                                                                                         \
    cudaError_t e__ = (call);                                                                                  \
    if (e__ != cudaSuccess) {                                                                                  \
      fail(nullptr, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, #call);           \
      bgpu_destroy(c); /* frees whatever has been created so far (every member is null-checked) */            \
      return 1;                                                                                                \
    }                                                                                                          \
  } while (0)
  CUC(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CUC(cudaGetDeviceProperties(&prop, dev));
  c->n_sm = prop.multiProcessorCount;
  // tuning sweeps (tools/): BGPU_SCATTER_BATCH, BGPU_AGGREGATE, BGPU_CHUNK override the defaults of a new context
  if (const char *e = getenv("BGPU_SCATTER_BATCH")) { const int v = atoi(e); if (v >= 1 && v <= 32) { c->scatter_batch = (uint32_t)v; c->scatter_batch_auto = false; } }
  if (const char *e = getenv("BGPU_AGGREGATE")) c->aggregate = atoi(e) ? 1 : 0;
  if (const char *e = getenv("BGPU_TALLY_COPIES")) { const int v = atoi(e); if (v >= 0 && v <= 1024) c->tally_copies = v; }
  if (const char *e = getenv("BGPU_EVENT_HBM")) c->event_hbm = atoi(e) ? 1 : 0;
  if (const char *e = getenv("BGPU_KERNEL")) {
    const std::string k(e);
    c->kernel_choice = k == "history" ? 1 : (k == "queues" ? 2 : 0);
  }
  if (const char *e = getenv("BGPU_POOL_TS")) { const int v = atoi(e); if (v >= 1 && v <= 32) c->pool_batch_scatter = (uint32_t)v; }
  if (const char *e = getenv("BGPU_POOL_TR")) { const int v = atoi(e); if (v >= 1 && v <= 32) c->pool_batch_refill = (uint32_t)v; }
  if (const char *e = getenv("BGPU_CHUNK")) { const int v = atoi(e); if (v > 0) { c->chunk = (uint32_t)v; c->chunk_auto = false; } }
  CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto &ev : c->ev) CUC(cudaEventCreate(&ev));
  c->mesh.nx = d->nx; c->mesh.ny = d->ny; c->mesh.nz = d->nz; c->mesh.G = d->n_groups;
  c->mesh.n_cells = d->nx * d->ny * d->nz;
  c->mesh.n_faces = d->nx + d->ny + d->nz + 3;
  for (int i = 0; i < 6; ++i) c->mesh.bc[i] = d->bc[i];
  c->seed = d->seed;
  c->ctr_hi = ((uint64_t)d->seed) << 32;
  c->n_user = d->n_user_photons;
  c->rank = d->rank;
  c->n_ranks = d->n_ranks < 1 ? 1 : d->n_ranks;
  std::vector<double> faces;
  faces.insert(faces.end(), d->x_faces, d->x_faces + d->nx + 1);
  faces.insert(faces.end(), d->y_faces, d->y_faces + d->ny + 1);
  faces.insert(faces.end(), d->z_faces, d->z_faces + d->nz + 1);
  CUC(cudaMalloc((void **)&c->d_faces, 8 * faces.size()));
  CUC(cudaMemcpy(c->d_faces, faces.data(), 8 * faces.size(), cudaMemcpyHostToDevice));
  c->mesh.faces = c->d_faces;
  const uint64_t nc = c->mesh.n_cells, G = c->mesh.G;
  CUC(cudaMalloc((void **)&c->d_f, 8 * nc));
  CUC(cudaMalloc((void **)&c->d_opa, 8 * nc * G));
  CUC(cudaMalloc((void **)&c->d_ops, 8 * nc * G));
  CUC(cudaMalloc((void **)&c->d_cellrec, 32 * nc));
  CUC(cudaMalloc((void **)&c->d_cell_stage, 8 * nc * 3));
  CUC(cudaMalloc((void **)&c->d_tally, 16 * nc));
  CUC(cudaMemset(c->d_tally, 0, 16 * nc));
  CUC(cudaMalloc((void **)&c->d_stats, 8 * ST_COUNT));
  CUC(cudaMemset(c->d_stats, 0, 8 * ST_COUNT));
  CUC(cudaMalloc((void **)&c->d_work_counter, 8 * WORK_COUNTERS));
  CUC(cudaMalloc((void **)&c->d_results, 8 * 8));
  uint64_t cap = d->photon_capacity;
  if (!cap) cap = (uint64_t)(1.25 * (double)d->n_user_photons / (double)c->n_ranks) + 1024;
  if (ensure_work(c, cap, 0)) {
    g_create_error = c->err;
    bgpu_destroy(c);
    return 1;
  }
#undef CUC
  if (prepare_tally_copies(c)) {  // allocated here so that no cycle's timing carries a cudaMalloc
    g_create_error = c->err;
    bgpu_destroy(c);
    return 1;
  }
  *out = c;
  return 0;
}

void bgpu_destroy(bgpu_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  DevBuf *bufs[] = {&c->scr_counts, &c->scr_offsets, &c->scr_tile_sum, &c->scr_tile_off, &c->scr_tiles, &c->scr_ndep,
                    &c->scr_dep_off, &c->scr_dep_cell, &c->scr_dep_val, &c->scr_sort, &c->scr_keys_out,
                    &c->scr_vals_in, &c->scr_vals_out, &c->scr_seg, &c->scr_aos, &c->scr_event, &c->scr_tally_rep,
                    &c->scr_comb, &c->scr_src_win};
  for (DevBuf *b : bufs)
    if (b->p) cudaFree(b->p);
  void *ptrs[] = {c->d_faces, c->d_f, c->d_opa, c->d_ops, c->d_cellrec, c->d_cell_stage, c->d_tally, c->d_stats, c->d_work_counter,
                  c->d_results, c->work.base, c->census.base, c->comb_scratch.base, c->d_desc, c->d_counters, c->d_regions,
                  c->d_region_of_cell, c->d_mesh, c->d_tile_sums, c->d_mesh_sums};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  c->comm.reset();
  if (c->h_comm) cudaFreeHost(c->h_comm);
  if (c->d_comm_scratch) cudaFree(c->d_comm_scratch);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  for (auto &ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_k2) cudaStreamDestroy(c->s_k2);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int bgpu_set_cell_data(bgpu_ctx *c, const double *f, const double *op_a, const double *op_s) {
  if (!c || !f || !op_a || !op_s) return fail(c, "bgpu_set_cell_data: null argument");
  CU(c, cudaSetDevice(c->device));
  const uint64_t nc = c->mesh.n_cells;
  CU(c, cudaMemcpyAsync(c->d_f, f, 8 * nc, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->d_cell_stage, op_a, 8 * nc, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->d_cell_stage + nc, op_s, 8 * nc, cudaMemcpyHostToDevice, c->stream));
  ++c->launches;
  k_expand_groups<<<grid_for(nc * c->mesh.G, 256), 256, 0, c->stream>>>(c->mesh.n_cells, c->mesh.G, c->d_cell_stage,
                                                                        c->d_cell_stage + nc, c->d_opa, c->d_ops);
  ++c->launches;
  k_fill_cellrec<<<grid_for(nc, 256), 256, 0, c->stream>>>(c->mesh.n_cells, c->mesh.G, c->d_f, c->d_opa, c->d_ops,
                                                          c->d_cellrec);
  CU(c, cudaGetLastError());
  CU(c, cudaStreamSynchronize(c->stream));  // the host arrays may be rewritten as soon as we return
  c->have_cell_data = true;
  c->uniform_groups = true;
  return 0;
}

int bgpu_set_cell_groups(bgpu_ctx *c, const double *f, const double *abs_groups, const double *sct_groups) {
  if (!c || !f || !abs_groups || !sct_groups) return fail(c, "bgpu_set_cell_groups: null argument");
  CU(c, cudaSetDevice(c->device));
  const uint64_t nc = c->mesh.n_cells, G = c->mesh.G;
  CU(c, cudaMemcpyAsync(c->d_f, f, 8 * nc, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->d_opa, abs_groups, 8 * nc * G, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->d_ops, sct_groups, 8 * nc * G, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemsetAsync(c->d_stats, 0, 8 * ST_COUNT, c->stream));
  ++c->launches;
  k_count_nonuniform_cells<<<grid_for(nc, 256), 256, 0, c->stream>>>(c->mesh.n_cells, c->mesh.G, c->d_opa, c->d_ops,
                                                                    c->d_stats);
  ++c->launches;
  k_fill_cellrec<<<grid_for(nc, 256), 256, 0, c->stream>>>(c->mesh.n_cells, c->mesh.G, c->d_f, c->d_opa, c->d_ops,
                                                          c->d_cellrec);  // (only read if the groups turn out uniform)
  unsigned long long nonuniform = 0;
  CU(c, cudaMemcpyAsync(&nonuniform, c->d_stats, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->have_cell_data = true;
  c->uniform_groups = nonuniform == 0;
  return 0;
}

}  // extern "C"

namespace {
// k_source_windows + k_source_sample for the S.n photons of one kind set
int launch_source_sample(bgpu_ctx *c, SourceParams S) {
  const uint32_t n_blocks = grid_for(S.n, SOURCE_BLOCK);
  if (ensure(c, c->scr_src_win, 4ull * (n_blocks + 1))) return 1;
  S.first_entry = (const uint32_t *)c->scr_src_win.p;
  c->launches += 2;
  k_source_windows<<<grid_for(n_blocks + 1, 256), 256, 0, c->stream>>>(S.offsets, S.n_entries, S.n, SOURCE_BLOCK, n_blocks,
                                                                     (uint32_t *)c->scr_src_win.p);
  k_source_sample<<<n_blocks, SOURCE_BLOCK, 0, c->stream>>>(S);
  CU(c, cudaGetLastError());
  return 0;
}

// make_photons / make_initial_census_photons / join_photon_arrays from per-cell energies that are already on the
// device (ev[0] has been recorded by the caller)
int source_from_device(bgpu_ctx *c, uint32_t cycle, double dt, const double *dE_emission, const double *dE_source,
                       const double *dE_census /* nullptr unless cycle 1 */, double total_E, uint64_t *n_new_out,
                       uint64_t *n_total_out) {
  const uint32_t nc = c->mesh.n_cells;
  const double *E_census = dE_census;
  if (ensure(c, c->scr_counts, 4ull * 3 * nc)) return 1;
  if (ensure(c, c->scr_offsets, 8ull * (3ull * nc + 2))) return 1;
  uint32_t *cnt2 = (uint32_t *)c->scr_counts.p, *cnt1 = cnt2 + 2ull * nc;
  uint64_t *off2 = (uint64_t *)c->scr_offsets.p, *off1 = off2 + 2ull * nc + 1;
  ++c->launches;
  k_source_count<<<grid_for(nc, 256), 256, 0, c->stream>>>(nc, 2, dE_emission, dE_source, c->n_user, total_E, cnt2);
  if (device_scan(c, cnt2, 2ull * nc, off2)) return 1;
  uint64_t n_new = 0, n_init = 0;
  CU(c, cudaMemcpyAsync(&n_new, off2 + 2ull * nc, 8, cudaMemcpyDeviceToHost, c->stream));
  if (E_census) {
    ++c->launches;
    k_source_count<<<grid_for(nc, 256), 256, 0, c->stream>>>(nc, 1, dE_census, nullptr, c->n_user, total_E, cnt1);
    if (device_scan(c, cnt1, nc, off1)) return 1;
    CU(c, cudaMemcpyAsync(&n_init, off1 + nc, 8, cudaMemcpyDeviceToHost, c->stream));
  }
  CU(c, cudaStreamSynchronize(c->stream));
  // cycle 1: the initial census replaces whatever census was there (census_photons = make_initial_census_photons)
  const uint64_t n_cen = E_census ? n_init : c->n_census;
  const uint64_t n_total = n_new + n_cen;
  if (n_total >= (1ull << 32)) return fail(c, "bgpu_source: %llu photons exceed the reference's 2^32 limit", (unsigned long long)n_total);
  if (ensure_work(c, n_total, 0)) return 1;
  SourceParams S{};
  S.ph = c->work;
  S.mesh = c->mesh;
  S.ctr_hi = c->ctr_hi;
  S.dt = dt;
  if (n_new) {
    S.dst_offset = 0;
    S.n = n_new;
    S.offsets = off2;
    S.n_entries = 2 * nc;
    S.kinds = 2;
    S.E0 = dE_emission;
    S.E1 = dE_source;
    // src/source.h:221-222
    S.stream_base = 10000000000000ULL * (uint64_t)cycle + c->n_user * (uint64_t)c->rank;
    if (launch_source_sample(c, S)) return 1;
  }
  if (E_census) {
    if (n_init) {
      S.dst_offset = n_new;
      S.n = n_init;
      S.offsets = off1;
      S.n_entries = nc;
      S.kinds = 1;
      S.E0 = dE_census;
      S.E1 = nullptr;
      S.stream_base = c->n_user * (uint64_t)c->rank;  // src/source.h:144
      if (launch_source_sample(c, S)) return 1;
    }
  } else if (n_cen) {
    // join_photon_arrays: all = [new ..., census ...] (src/census_functions.h:21-29)
    ++c->launches;
    k_copy_soa<<<grid_for(n_cen, 256), 256, 0, c->stream>>>(c->census, 0, c->work, n_new, n_cen);
  }
  CU(c, cudaGetLastError());
  // every photon of the work list starts as PASS (the reference's SoA setters do the same, src/source.h:27-62): a dump of
  // the list before transport reads defined descriptors
  if (n_total) CU(c, cudaMemsetAsync(c->d_desc, EV_PASS, n_total, c->stream));
  c->n_new = n_new;
  c->n_work = n_total;
  // pre-transport census energy, get_photon_list_E (src/replicated_driver.h:61,71)
  // (the second reduction reuses the tile scratch of the first: both are ordered on the ctx stream)
  if (list_energy_async(c, c->work, n_new, n_cen, 2)) return 1;
  if (list_energy_async(c, c->work, 0, n_new, 3, 1)) return 1;
  double h_E[2] = {0.0, 0.0};
  CU(c, cudaMemcpyAsync(h_E, c->d_results + 2, 16, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaEventRecord(c->ev[1], c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->pre_census_E = h_E[0];
  c->new_photon_E = h_E[1];
  c->stats = bgpu_cycle_stats{};
  c->stats.pre_census_E = c->pre_census_E;
  c->stats.new_photon_E = c->new_photon_E;
  c->stats.n_new = n_new;
  c->stats.n_transported = n_total;
  CU(c, cudaEventElapsedTime(&c->stats.ms_source, c->ev[0], c->ev[1]));
  if (n_new_out) *n_new_out = n_new;
  if (n_total_out) *n_total_out = n_total;
  return 0;
}
}  // namespace

extern "C" {

int bgpu_source(bgpu_ctx *c, uint32_t cycle, double dt, const double *E_emission, const double *E_source,
                const double *E_census, double total_E, uint64_t *n_new_out, uint64_t *n_total_out) {
  if (!c || !E_emission || !E_source) return fail(c, "bgpu_source: null argument");
  CU(c, cudaSetDevice(c->device));
  const uint32_t nc = c->mesh.n_cells;
  double *dE = c->d_cell_stage;
  CU(c, cudaMemcpyAsync(dE, E_emission, 8ull * nc, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(dE + nc, E_source, 8ull * nc, cudaMemcpyHostToDevice, c->stream));
  if (E_census) CU(c, cudaMemcpyAsync(dE + 2ull * nc, E_census, 8ull * nc, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaEventRecord(c->ev[0], c->stream));  // inputs are resident from here on
  return source_from_device(c, cycle, dt, dE, dE + nc, E_census ? dE + 2ull * nc : nullptr, total_E, n_new_out,
                            n_total_out);
}

// ---------------------------------------------------------------------------------------------------------------
// device-resident mesh physics (mesh_dev.cuh)
// ---------------------------------------------------------------------------------------------------------------
namespace {
enum : int { MA_T_E = 0, MA_T_R0, MA_T_S, MA_T_R, MA_OP_A, MA_OP_S, MA_E_EMISSION, MA_E_SOURCE, MA_E_CENSUS, MA_E_EMISSION_GLOBAL, MA_N };
inline double *mesh_arr(bgpu_ctx *c, int which) { return c->d_mesh + (uint64_t)which * c->mesh.n_cells; }

MeshPhysParams mesh_params(bgpu_ctx *c) {
  MeshPhysParams P{};
  P.mesh = c->mesh;
  P.regions = c->d_regions;
  P.region_of_cell = c->d_region_of_cell;
  P.T_e = mesh_arr(c, MA_T_E);
  P.T_r0 = mesh_arr(c, MA_T_R0);
  P.T_s = mesh_arr(c, MA_T_S);
  P.T_r = mesh_arr(c, MA_T_R);
  P.f = c->d_f;
  P.op_a = mesh_arr(c, MA_OP_A);
  P.op_s = mesh_arr(c, MA_OP_S);
  P.E_emission = mesh_arr(c, MA_E_EMISSION);
  P.E_source = mesh_arr(c, MA_E_SOURCE);
  P.E_census = mesh_arr(c, MA_E_CENSUS);
  P.E_emission_global = mesh_arr(c, MA_E_EMISSION_GLOBAL);
  P.tally = (double2 *)c->d_tally;
  P.tile_sums = c->d_tile_sums;
  P.n_tiles = c->mesh_tiles;
  P.dt = c->mesh_dt;
  P.replicated_factor = 1.0 / static_cast<double>(c->n_ranks);  // src/mesh.h:87
  P.n_user = (uint64_t)(uint32_t)c->n_user;  // the reference passes n_user_photons as uint32_t here (src/mesh.h:237)
  P.step = c->mesh_step;
  P.rank = c->rank;
  P.n_ranks = c->n_ranks;
  return P;
}

void unpack_sums(const double *h, uint32_t q_mask, bgpu_mesh_sums *out) {
  if (q_mask & (1u << MS_PRE_MAT)) out->pre_mat_E = h[MS_PRE_MAT];
  if (q_mask & (1u << MS_EMISSION)) out->emission_E = h[MS_EMISSION];
  if (q_mask & (1u << MS_CENSUS)) out->census_E = h[MS_CENSUS];
  if (q_mask & (1u << MS_SOURCE)) out->source_E = h[MS_SOURCE];
  if (q_mask & (1u << MS_TOTAL)) out->total_photon_E = h[MS_TOTAL];
  if (q_mask & (1u << MS_ABS)) out->absorbed_E = h[MS_ABS];
  if (q_mask & (1u << MS_POST_MAT)) out->post_mat_E = h[MS_POST_MAT];
}

// the tile sums -> d_mesh_sums[block][q] for `n_blocks` rank blocks (no host synchronisation)
int mesh_final_sums_async(bgpu_ctx *c, uint32_t q_mask, uint32_t n_blocks) {
  ++c->launches;
  k_mesh_final_sums<<<n_blocks, 256, 0, c->stream>>>(c->d_tile_sums, c->mesh_tiles, q_mask, c->d_mesh_sums);
  CU(c, cudaGetLastError());
  return 0;
}

int mesh_fetch_sums(bgpu_ctx *c, uint32_t q_mask, bgpu_mesh_sums *out, uint32_t n_blocks = 1) {
  if (mesh_final_sums_async(c, q_mask, n_blocks)) return 1;
  std::vector<double> h((size_t)MS_N * n_blocks);
  CU(c, cudaMemcpyAsync(h.data(), c->d_mesh_sums, 8 * h.size(), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (uint32_t b = 0; b < n_blocks; ++b) unpack_sums(h.data() + (size_t)MS_N * b, q_mask, out + b);
  return 0;
}
}  // namespace

int bgpu_mesh_init(bgpu_ctx *c, uint32_t n_regions, const bgpu_region *regions, const uint32_t *region_of_cell,
                   const double *T_e, const double *T_r, const double *T_s) {
  if (!c || !n_regions || !regions || !region_of_cell || !T_e || !T_r || !T_s)
    return fail(c, "bgpu_mesh_init: null argument");
  static_assert(sizeof(bgpu_region) == sizeof(RegionDev), "bgpu_region layout");
  CU(c, cudaSetDevice(c->device));
  const uint64_t nc = c->mesh.n_cells;
  for (uint64_t i = 0; i < nc; ++i)
    if (region_of_cell[i] >= n_regions) return fail(c, "bgpu_mesh_init: cell %llu names region %u of %u",
                                                    (unsigned long long)i, region_of_cell[i], n_regions);
  if (!c->d_mesh) {
    c->mesh_tiles = (uint32_t)((nc + MESH_TILE - 1) / MESH_TILE);
    CU(c, cudaMalloc((void **)&c->d_region_of_cell, 4 * nc));
    CU(c, cudaMalloc((void **)&c->d_mesh, 8 * nc * MA_N));
    // (one block of sums per rank: k_mesh_redistribute forms every rank's totals, mesh_dev.cuh)
    CU(c, cudaMalloc((void **)&c->d_tile_sums, 8ull * MS_N * c->mesh_tiles * c->n_ranks));
    CU(c, cudaMalloc((void **)&c->d_mesh_sums, 8ull * MS_N * c->n_ranks));
  }
  if (c->d_regions) CU(c, cudaFree(c->d_regions));
  CU(c, cudaMalloc((void **)&c->d_regions, sizeof(RegionDev) * n_regions));
  CU(c, cudaMemcpyAsync(c->d_regions, regions, sizeof(RegionDev) * n_regions, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->d_region_of_cell, region_of_cell, 4 * nc, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemsetAsync(c->d_mesh, 0, 8 * nc * MA_N, c->stream));
  CU(c, cudaMemcpyAsync(mesh_arr(c, MA_T_E), T_e, 8 * nc, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(mesh_arr(c, MA_T_R0), T_r, 8 * nc, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(mesh_arr(c, MA_T_S), T_s, 8 * nc, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->mesh_ready = true;
  return 0;
}

int bgpu_mesh_calculate_photon_energy(bgpu_ctx *c, double dt, uint32_t step, bgpu_mesh_sums *sums) {
  if (!c || !sums) return fail(c, "bgpu_mesh_calculate_photon_energy: null argument");
  if (!c->mesh_ready) return fail(c, "bgpu_mesh_calculate_photon_energy: call bgpu_mesh_init first");
  CU(c, cudaSetDevice(c->device));
  c->mesh_dt = dt;
  c->mesh_step = step;
  c->mesh_redistributed = false;
  const MeshPhysParams P = mesh_params(c);
  ++c->launches;
  k_mesh_energy<<<c->mesh_tiles, MESH_TILE_THREADS, 0, c->stream>>>(P);
  // every group gets the cell's gray opacity (Cell::set_op_a / set_op_s, src/cell.h:260-275)
  ++c->launches;
  k_expand_groups<<<grid_for((uint64_t)c->mesh.n_cells * c->mesh.G, 256), 256, 0, c->stream>>>(
      c->mesh.n_cells, c->mesh.G, P.op_a, P.op_s, c->d_opa, c->d_ops);
  ++c->launches;
  k_fill_cellrec<<<grid_for(c->mesh.n_cells, 256), 256, 0, c->stream>>>(c->mesh.n_cells, c->mesh.G, c->d_f, c->d_opa,
                                                                        c->d_ops, c->d_cellrec);
  CU(c, cudaGetLastError());
  c->have_cell_data = true;
  c->uniform_groups = true;
  *sums = bgpu_mesh_sums{};
  return mesh_fetch_sums(c, (1u << MS_PRE_MAT) | (1u << MS_EMISSION) | (1u << MS_CENSUS) | (1u << MS_SOURCE) |
                                (1u << MS_TOTAL), sums);
}

int bgpu_mesh_redistribute(bgpu_ctx *c, double global_source_E, bgpu_mesh_sums *sums) {
  if (!c || !sums) return fail(c, "bgpu_mesh_redistribute: null argument");
  if (!c->mesh_ready) return fail(c, "bgpu_mesh_redistribute: call bgpu_mesh_init first");
  CU(c, cudaSetDevice(c->device));
  MeshPhysParams P = mesh_params(c);
  P.global_source_E = global_source_E;
  ++c->launches;
  k_mesh_redistribute<<<c->mesh_tiles, MESH_TILE_THREADS, 0, c->stream>>>(P, nullptr);
  CU(c, cudaGetLastError());
  c->mesh_redistributed = true;
  // (this rank's block of the per-rank sums)
  std::vector<bgpu_mesh_sums> all((size_t)c->n_ranks);
  const uint32_t mask = (1u << MS_EMISSION) | (1u << MS_CENSUS) | (1u << MS_SOURCE) | (1u << MS_TOTAL);
  if (mesh_fetch_sums(c, mask, all.data(), (uint32_t)c->n_ranks)) return 1;
  const bgpu_mesh_sums &mine = all[(size_t)c->rank];
  sums->emission_E = mine.emission_E;
  sums->census_E = mine.census_E;
  sums->source_E = mine.source_E;
  sums->total_photon_E = mine.total_photon_E;
  return 0;
}

// calculate_photon_energy for a replicated run without its collectives (see k_mesh_redistribute): one launch chain, one
// host synchronisation; rank_sums[r] = rank r's totals after the redistribution (pre_mat_E is the same for all)
int bgpu_mesh_calculate_photon_energy_replicated(bgpu_ctx *c, double dt, uint32_t step, bgpu_mesh_sums *rank_sums) {
  if (!c || !rank_sums) return fail(c, "bgpu_mesh_calculate_photon_energy_replicated: null argument");
  if (!c->mesh_ready) return fail(c, "bgpu_mesh_calculate_photon_energy_replicated: call bgpu_mesh_init first");
  if (c->n_ranks == 1) return bgpu_mesh_calculate_photon_energy(c, dt, step, rank_sums);
  CU(c, cudaSetDevice(c->device));
  c->mesh_dt = dt;
  c->mesh_step = step;
  const MeshPhysParams P = mesh_params(c);
  ++c->launches;
  k_mesh_energy<<<c->mesh_tiles, MESH_TILE_THREADS, 0, c->stream>>>(P);
  const uint32_t m_energy = (1u << MS_PRE_MAT) | (1u << MS_EMISSION) | (1u << MS_CENSUS) | (1u << MS_SOURCE) |
                            (1u << MS_TOTAL);
  if (mesh_final_sums_async(c, m_energy, 1)) return 1;
  double pre_mat = 0.0;
  CU(c, cudaMemcpyAsync(&pre_mat, c->d_mesh_sums + MS_PRE_MAT, 8, cudaMemcpyDeviceToHost, c->stream));
  ++c->launches;
  k_mesh_redistribute<<<c->mesh_tiles, MESH_TILE_THREADS, 0, c->stream>>>(P, c->d_mesh_sums);
  // every group gets the cell's gray opacity (Cell::set_op_a / set_op_s, src/cell.h:260-275)
  ++c->launches;
  k_expand_groups<<<grid_for((uint64_t)c->mesh.n_cells * c->mesh.G, 256), 256, 0, c->stream>>>(
      c->mesh.n_cells, c->mesh.G, P.op_a, P.op_s, c->d_opa, c->d_ops);
  ++c->launches;
  k_fill_cellrec<<<grid_for(c->mesh.n_cells, 256), 256, 0, c->stream>>>(c->mesh.n_cells, c->mesh.G, c->d_f, c->d_opa,
                                                                        c->d_ops, c->d_cellrec);
  CU(c, cudaGetLastError());
  c->have_cell_data = true;
  c->uniform_groups = true;
  c->mesh_redistributed = true;
  for (int r = 0; r < c->n_ranks; ++r) rank_sums[r] = bgpu_mesh_sums{};
  const uint32_t m_red = (1u << MS_EMISSION) | (1u << MS_CENSUS) | (1u << MS_SOURCE) | (1u << MS_TOTAL);
  if (mesh_fetch_sums(c, m_red, rank_sums, (uint32_t)c->n_ranks)) return 1;  // (synchronises: pre_mat has arrived)
  for (int r = 0; r < c->n_ranks; ++r) rank_sums[r].pre_mat_E = pre_mat;
  return 0;
}

int bgpu_mesh_source(bgpu_ctx *c, uint32_t cycle, double total_E, uint64_t *n_new_out, uint64_t *n_total_out) {
  if (!c) return fail(c, "bgpu_mesh_source: null ctx");
  if (!c->mesh_ready || c->mesh_step != cycle)
    return fail(c, "bgpu_mesh_source: bgpu_mesh_calculate_photon_energy has not run for cycle %u", cycle);
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaEventRecord(c->ev[0], c->stream));
  return source_from_device(c, cycle, c->mesh_dt, mesh_arr(c, MA_E_EMISSION), mesh_arr(c, MA_E_SOURCE),
                            cycle == 1 ? mesh_arr(c, MA_E_CENSUS) : nullptr, total_E, n_new_out, n_total_out);
}

int bgpu_mesh_update_temperature(bgpu_ctx *c, bgpu_mesh_sums *sums) {
  if (!c || !sums) return fail(c, "bgpu_mesh_update_temperature: null argument");
  if (!c->mesh_ready) return fail(c, "bgpu_mesh_update_temperature: call bgpu_mesh_init first");
  CU(c, cudaSetDevice(c->device));
  const MeshPhysParams P = mesh_params(c);
  ++c->launches;
  // multi-rank: the all-reduced emission (src/mesh.h:343-345); one rank: m_emission_E itself
  k_mesh_update_temperature<<<c->mesh_tiles, MESH_TILE_THREADS, 0, c->stream>>>(
      P, c->mesh_redistributed ? P.E_emission_global : P.E_emission);
  CU(c, cudaGetLastError());
  return mesh_fetch_sums(c, (1u << MS_ABS) | (1u << MS_POST_MAT), sums);
}

int bgpu_mesh_get(bgpu_ctx *c, const char *name, double *out) {
  if (!c || !name || !out) return fail(c, "bgpu_mesh_get: null argument");
  if (!c->mesh_ready) return fail(c, "bgpu_mesh_get: call bgpu_mesh_init first");
  CU(c, cudaSetDevice(c->device));
  const uint64_t nc = c->mesh.n_cells;
  const std::string k(name);
  const double *src = nullptr;
  if (k == "T_e") src = mesh_arr(c, MA_T_E);
  else if (k == "T_r") src = mesh_arr(c, MA_T_R);
  else if (k == "T_s") src = mesh_arr(c, MA_T_S);
  else if (k == "f") src = c->d_f;
  else if (k == "op_a") src = mesh_arr(c, MA_OP_A);
  else if (k == "op_s") src = mesh_arr(c, MA_OP_S);
  else if (k == "E_emission") src = mesh_arr(c, MA_E_EMISSION);
  else if (k == "E_source") src = mesh_arr(c, MA_E_SOURCE);
  else if (k == "E_census") src = mesh_arr(c, MA_E_CENSUS);
  else if (k == "abs_E" || k == "track_E") {  // the (rank-summed) tallies of the last transport
    if (ensure_pinned(c, 16 * nc)) return 1;
    double *h = (double *)c->h_pinned;
    CU(c, cudaMemcpyAsync(h, c->d_tally, 16 * nc, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    const int o = k == "abs_E" ? 0 : 1;
    for (uint64_t i = 0; i < nc; ++i) out[i] = h[2 * i + o];
    return 0;
  } else {
    return fail(c, "bgpu_mesh_get: unknown array '%s'", name);
  }
  // through the context's pinned buffer: a copy straight into the caller's pageable array is staged by the driver in
  // small pieces and runs at a third of the link rate
  if (ensure_pinned(c, 8 * nc)) return 1;
  CU(c, cudaMemcpyAsync(c->h_pinned, src, 8 * nc, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  memcpy(out, c->h_pinned, 8 * nc);
  return 0;
}

int bgpu_transport(bgpu_ctx *c, double next_dt, int algorithm, int tally_mode) {
  if (!c) return fail(c, "bgpu_transport: null ctx");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaEventRecord(c->ev[2], c->stream));
  CU(c, cudaMemsetAsync(c->d_tally, 0, 16ull * c->mesh.n_cells, c->stream));
  if (run_transport(c, algorithm, tally_mode, c->counters_on /* validation: full state of every photon */)) return 1;
  CU(c, cudaEventRecord(c->ev[3], c->stream));
  if (run_census(c, next_dt, tally_mode == BGPU_TALLY_DETERMINISTIC)) return 1;
  CU(c, cudaEventRecord(c->ev[4], c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->stats.pre_census_E = c->pre_census_E;
  c->stats.new_photon_E = c->new_photon_E;
  c->stats.n_new = c->n_new;
  c->stats.n_transported = c->n_work;
  CU(c, cudaEventElapsedTime(&c->stats.ms_transport, c->ev[2], c->ev[3]));
  CU(c, cudaEventElapsedTime(&c->stats.ms_census, c->ev[3], c->ev[4]));
  c->stats.ms_total = c->stats.ms_source + c->stats.ms_transport + c->stats.ms_census;
  c->stats_valid = true;
  return 0;
}

int bgpu_get_tallies(bgpu_ctx *c, double *abs_E, double *track_E, bgpu_cycle_stats *stats) {
  if (!c) return fail(c, "bgpu_get_tallies: null ctx");
  CU(c, cudaSetDevice(c->device));
  const uint64_t nc = c->mesh.n_cells;
  if (abs_E || track_E) {
    if (ensure_pinned(c, 16 * nc)) return 1;
    double *h = (double *)c->h_pinned;
    CU(c, cudaMemcpyAsync(h, c->d_tally, 16 * nc, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (abs_E)
      for (uint64_t i = 0; i < nc; ++i) abs_E[i] = h[2 * i];
    if (track_E)
      for (uint64_t i = 0; i < nc; ++i) track_E[i] = h[2 * i + 1];
  }
  c->stats.n_launches = c->launches;
  if (stats) *stats = c->stats;
  return 0;
}

}  // extern "C"
namespace {
// the tally buffer with room for `extra` doubles behind the 2 * n_cells tallies (contents kept)
int ensure_tally_extra(bgpu_ctx *c, uint64_t extra) {
  const uint64_t nc = c->mesh.n_cells;
  if (extra <= c->tally_extra) return 0;
  double *p = nullptr;
  CU(c, cudaMalloc((void **)&p, 8 * (2 * nc + extra)));
  CU(c, cudaMemcpyAsync(p, c->d_tally, 8 * (2 * nc + c->tally_extra), cudaMemcpyDeviceToDevice, c->stream));
  CU(c, cudaMemsetAsync(p + 2 * nc + c->tally_extra, 0, 8 * (extra - c->tally_extra), c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaFree(c->d_tally));
  c->d_tally = p;
  c->tally_extra = extra;
  return 0;
}

int ensure_comm_buffers(bgpu_ctx *c) {
  const uint64_t tail = (uint64_t)c->n_ranks * BGPU_RANK_SCALARS;
  if (ensure_tally_extra(c, tail)) return 1;
  if (!c->h_comm) CU(c, cudaHostAlloc((void **)&c->h_comm, 8 * (BGPU_RANK_SCALARS + tail + MS_N + 64), cudaHostAllocDefault));
  if (!c->d_comm_scratch) CU(c, cudaMalloc((void **)&c->d_comm_scratch, 8 * bgpu_ctx::COMM_SCRATCH_DOUBLES));
  return 0;
}

// in-place reduction of n device doubles over the ranks of this ctx's communicator, ordered on the ctx stream
int comm_allreduce_device(bgpu_ctx *c, double *ptr, uint64_t n, int op /* ncclRedOp_t value: 0 sum, 2 max, 3 min */) {
  if (c->n_ranks == 1) return 0;
  if (c->comm.kind == COMM_NCCL) {
    NcclApi &api = NcclApi::get();
    const ncclResult_t r = api.AllReduce(ptr, ptr, n, ncclDouble, (ncclRedOp_t)op, c->comm.nccl, c->stream);
    if (r != ncclSuccess) return fail(c, "ncclAllReduce: %s", api.GetErrorString(r));
  } else if (c->comm.kind == COMM_LOCAL) {
    const std::string e = local_allreduce(*c->comm.local, c->rank, ptr, n, op, c->stream);
    if (!e.empty()) return fail(c, "in-process all-reduce: %s", e.c_str());
  } else {
    return fail(c, "rank %d of %d has no communicator (bgpu_comm_init_rank / bgpu_comm_init_local)", c->rank, c->n_ranks);
  }
  c->comm.bytes += 8 * n;
  ++c->comm.calls;
  return 0;
}

// The cycle's one collective, enqueued on the ctx stream without a host synchronisation: this rank's row of the tail is
// filled from `rank_scalars`, the others are zeroed, and {tallies, tail} are summed in place over the ranks.
int reduce_tallies_async(bgpu_ctx *c, const double *rank_scalars) {
  if (ensure_comm_buffers(c)) return 1;
  const uint64_t nc = c->mesh.n_cells, tail = (uint64_t)c->n_ranks * BGPU_RANK_SCALARS;
  double *d_tail = c->d_tally + 2 * nc;
  for (int i = 0; i < BGPU_RANK_SCALARS; ++i) c->h_comm[i] = rank_scalars ? rank_scalars[i] : 0.0;
  CU(c, cudaMemsetAsync(d_tail, 0, 8 * tail, c->stream));
  CU(c, cudaMemcpyAsync(d_tail + (uint64_t)c->rank * BGPU_RANK_SCALARS, c->h_comm, 8 * BGPU_RANK_SCALARS,
                        cudaMemcpyHostToDevice, c->stream));
  return comm_allreduce_device(c, c->d_tally, 2 * nc + tail, 0);
}
}  // namespace
extern "C" {

int bgpu_tally_buffer(bgpu_ctx *c, uint64_t extra, void **device_ptr, uint64_t *n_doubles) {
  if (!c || !device_ptr || !n_doubles) return fail(c, "bgpu_tally_buffer: null argument");
  CU(c, cudaSetDevice(c->device));
  if (ensure_tally_extra(c, extra)) return 1;
  *device_ptr = c->d_tally;
  *n_doubles = 2 * (uint64_t)c->mesh.n_cells + extra;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// replicated-mode collectives (comm_native.cuh)
// ---------------------------------------------------------------------------------------------------------------
int bgpu_comm_unique_id(char id[BGPU_COMM_ID_BYTES]) {
  static_assert(BGPU_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");
  if (!id) return fail(nullptr, "bgpu_comm_unique_id: null argument");
  NcclApi &api = NcclApi::get();
  if (!api.ok()) return fail(nullptr, "bgpu_comm_unique_id: %s", api.error.c_str());
  ncclUniqueId u;
  const ncclResult_t r = api.GetUniqueId(&u);
  if (r != ncclSuccess) return fail(nullptr, "ncclGetUniqueId: %s", api.GetErrorString(r));
  memcpy(id, u.internal, BGPU_COMM_ID_BYTES);
  return 0;
}

int bgpu_comm_init_rank(bgpu_ctx *c, const char id[BGPU_COMM_ID_BYTES]) {
  if (!c || !id) return fail(c, "bgpu_comm_init_rank: null argument");
  if (c->n_ranks == 1) return 0;
  NcclApi &api = NcclApi::get();
  if (!api.ok()) return fail(c, "bgpu_comm_init_rank: %s", api.error.c_str());
  CU(c, cudaSetDevice(c->device));
  c->comm.reset();
  ncclUniqueId u;
  memcpy(u.internal, id, BGPU_COMM_ID_BYTES);
  const ncclResult_t r = api.CommInitRank(&c->comm.nccl, c->n_ranks, u, c->rank);
  if (r != ncclSuccess) return fail(c, "ncclCommInitRank(rank %d of %d): %s", c->rank, c->n_ranks, api.GetErrorString(r));
  c->comm.kind = COMM_NCCL;
  return ensure_comm_buffers(c);
}

int bgpu_comm_init_local(bgpu_ctx **ctxs, int n) {
  if (!ctxs || n < 1) return fail(nullptr, "bgpu_comm_init_local: null argument");
  bool same_device = true, distinct = true;
  for (int r = 0; r < n; ++r) {
    if (!ctxs[r]) return fail(nullptr, "bgpu_comm_init_local: null ctx");
    if (ctxs[r]->rank != r || ctxs[r]->n_ranks != n)
      return fail(ctxs[r], "bgpu_comm_init_local: ctxs[%d] was created as rank %d of %d", r, ctxs[r]->rank, ctxs[r]->n_ranks);
    same_device = same_device && ctxs[r]->device == ctxs[0]->device;
    for (int q = 0; q < r; ++q) distinct = distinct && ctxs[q]->device != ctxs[r]->device;
  }
  if (n == 1) return 0;
  for (int r = 0; r < n; ++r) ctxs[r]->comm.reset();
  if (distinct) {
    // one GPU per rank: NCCL over NVLink (one process, one host thread per rank)
    NcclApi &api = NcclApi::get();
    if (!api.ok()) return fail(ctxs[0], "bgpu_comm_init_local: %s", api.error.c_str());
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n);
    for (int r = 0; r < n; ++r) devs[r] = ctxs[r]->device;
    const ncclResult_t e = api.CommInitAll(comms.data(), n, devs.data());
    if (e != ncclSuccess) return fail(ctxs[0], "ncclCommInitAll: %s", api.GetErrorString(e));
    for (int r = 0; r < n; ++r) {
      ctxs[r]->comm.nccl = comms[r];
      ctxs[r]->comm.kind = COMM_NCCL;
    }
  } else if (same_device) {
    if (n > LOCAL_MAX_RANKS) return fail(ctxs[0], "bgpu_comm_init_local: at most %d ranks on one device", LOCAL_MAX_RANKS);
    auto g = std::make_shared<LocalGroup>();
    g->n_ranks = n;
    g->device = ctxs[0]->device;
    CU(ctxs[0], cudaSetDevice(g->device));
    for (int r = 0; r < n; ++r) CU(ctxs[0], cudaEventCreateWithFlags(&g->ev_ready[r], cudaEventDisableTiming));
    CU(ctxs[0], cudaEventCreateWithFlags(&g->ev_done, cudaEventDisableTiming));
    for (int r = 0; r < n; ++r) {
      ctxs[r]->comm.local = g;
      ctxs[r]->comm.kind = COMM_LOCAL;
    }
  } else {
    return fail(ctxs[0], "bgpu_comm_init_local: ranks must sit on distinct devices (NCCL) or all on one (in-process sum)");
  }
  for (int r = 0; r < n; ++r) {
    CU(ctxs[r], cudaSetDevice(ctxs[r]->device));
    if (ensure_comm_buffers(ctxs[r])) return 1;
  }
  return 0;
}

int bgpu_comm_info(const bgpu_ctx *c, int *kind, uint64_t *calls, uint64_t *bytes) {
  if (!c) return 1;
  if (kind) *kind = c->comm.kind;
  if (calls) *calls = c->comm.calls;
  if (bytes) *bytes = c->comm.bytes;
  return 0;
}

int bgpu_comm_allreduce_host(bgpu_ctx *c, double *buf, uint64_t n, int op) {
  if (!c || (!buf && n)) return fail(c, "bgpu_comm_allreduce_host: null argument");
  if (op != BGPU_OP_SUM && op != BGPU_OP_MAX && op != BGPU_OP_MIN) return fail(c, "bgpu_comm_allreduce_host: unknown op %d", op);
  if (c->n_ranks == 1 || n == 0) return 0;
  CU(c, cudaSetDevice(c->device));
  if (ensure_comm_buffers(c)) return 1;
  const int nccl_op = op == BGPU_OP_SUM ? 0 : (op == BGPU_OP_MAX ? 2 : 3);
  for (uint64_t off = 0; off < n; off += bgpu_ctx::COMM_SCRATCH_DOUBLES) {
    const uint64_t m = std::min<uint64_t>(bgpu_ctx::COMM_SCRATCH_DOUBLES, n - off);
    CU(c, cudaMemcpyAsync(c->d_comm_scratch, buf + off, 8 * m, cudaMemcpyHostToDevice, c->stream));
    if (comm_allreduce_device(c, c->d_comm_scratch, m, nccl_op)) return 1;
    CU(c, cudaMemcpyAsync(buf + off, c->d_comm_scratch, 8 * m, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int bgpu_comm_allreduce_tallies(bgpu_ctx *c, const double *rank_scalars, double *all_scalars) {
  if (!c) return fail(c, "bgpu_comm_allreduce_tallies: null ctx");
  CU(c, cudaSetDevice(c->device));
  if (reduce_tallies_async(c, rank_scalars)) return 1;
  const uint64_t tail = (uint64_t)c->n_ranks * BGPU_RANK_SCALARS;
  double *h_tail = c->h_comm + BGPU_RANK_SCALARS;
  CU(c, cudaMemcpyAsync(h_tail, c->d_tally + 2ull * c->mesh.n_cells, 8 * tail, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (all_scalars) memcpy(all_scalars, h_tail, 8 * tail);
  return 0;
}

int bgpu_mesh_finish_cycle(bgpu_ctx *c, const double *rank_scalars, double *all_scalars, bgpu_mesh_sums *sums) {
  if (!c || !sums) return fail(c, "bgpu_mesh_finish_cycle: null argument");
  if (!c->mesh_ready) return fail(c, "bgpu_mesh_finish_cycle: call bgpu_mesh_init first");
  CU(c, cudaSetDevice(c->device));
  // tallies + tail summed over the ranks; src/replicated_driver.h:91-94 and the reductions of src/imc_state.h:207-252
  if (reduce_tallies_async(c, rank_scalars)) return 1;
  // mesh.update_temperature (src/replicated_driver.h:96) consumes the reduced tallies in place, stream-ordered
  const MeshPhysParams P = mesh_params(c);
  ++c->launches;
  k_mesh_update_temperature<<<c->mesh_tiles, MESH_TILE_THREADS, 0, c->stream>>>(
      P, c->mesh_redistributed ? P.E_emission_global : P.E_emission);
  CU(c, cudaGetLastError());
  const uint32_t mask = (1u << MS_ABS) | (1u << MS_POST_MAT);
  if (mesh_final_sums_async(c, mask, 1)) return 1;
  const uint64_t tail = (uint64_t)c->n_ranks * BGPU_RANK_SCALARS;
  double *h_tail = c->h_comm + BGPU_RANK_SCALARS, *h_sums = h_tail + tail;
  CU(c, cudaMemcpyAsync(h_tail, c->d_tally + 2ull * c->mesh.n_cells, 8 * tail, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(h_sums, c->d_mesh_sums, 8 * MS_N, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));  // the cycle's one host synchronisation after transport
  if (all_scalars) memcpy(all_scalars, h_tail, 8 * tail);
  unpack_sums(h_sums, mask, sums);
  return 0;
}

int bgpu_sync(bgpu_ctx *c) {
  if (!c) return 1;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  return 0;
}
void *bgpu_stream(bgpu_ctx *c) { return c ? (void *)c->stream : nullptr; }
int bgpu_device(const bgpu_ctx *c) { return c ? c->device : -1; }

namespace {

PhotonSoA soa_view(const PhotonSoA &s, uint64_t off) {
  PhotonSoA v = s;
  v.xy += off; v.za += off; v.bc += off; v.ee += off; v.lc += off; v.sg += off;
  v.cap = s.cap - off;
  return v;
}

// Pipelined form of the drop-in (history algorithm, atomic tallies): the photon list goes through the device in
// slices, so that the upload of slice j+1, the transport of slice j and the download of slice j-1 overlap.  Histories
// are independent, so slicing changes nothing per photon; all slices accumulate into the same
// tallies.
//
// The caller's photons live in pageable memory (a std::vector): a plain cudaMemcpyAsync from / to it is staged by the
// driver through one bounce buffer by the issuing thread and runs at a third of the link rate.  Here the staging is
// ours: aos_copiers() threads per direction, each with two pinned buffers of AOS_CHUNK photons, copy between the vector
// and their buffers (several cores' worth of memcpy bandwidth) while the DMA engines move the other buffers, so both
// PCIe directions stay busy behind the transport kernel.
constexpr uint64_t AOS_CHUNK = 1ull << 16;  // photons per staged copy (7.5 MB)
constexpr int AOS_MAX_COPIERS = 8;
constexpr int AOS_SLOTS = 2;                // pinned buffers per thread

// threads per direction: BRANSON_AOS_COPIERS (1..8); default 6, or 3/8 of the host's hardware threads if that is fewer
// (both directions together then use 3/4 of them).  Measured on a 16-thread host: 2 -> 118 ms, 4 -> 87 ms, 6 -> 79 ms,
// 8 -> no better (tools/aos_dropin_sweep.py; the copies are then bound by the host's memory bandwidth).
int aos_copiers() {
  int n = 6;
  const unsigned hw = std::thread::hardware_concurrency();
  if (hw && (int)(hw * 3 / 8) < n) n = std::max(1, (int)(hw * 3 / 8));
  if (const char *e = getenv("BRANSON_AOS_COPIERS")) n = atoi(e);
  return std::min(AOS_MAX_COPIERS, std::max(1, n));
}

int transport_aos_pipelined(bgpu_ctx *c, uint8_t *photons, uint64_t n, uint64_t *bad_out) {
  if (!c->s_in) {
    CU(c, cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    CU(c, cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    CU(c, cudaStreamCreateWithFlags(&c->s_k2, cudaStreamNonBlocking));
  }
  const int AOS_COPIERS = aos_copiers();
  const size_t slot_bytes = 120 * AOS_CHUNK;
  const size_t stage_bytes = slot_bytes * AOS_SLOTS * AOS_COPIERS * 2;
  if (c->h_stage_bytes < stage_bytes) {
    if (c->h_stage) cudaFreeHost(c->h_stage);
    c->h_stage = nullptr;
    c->h_stage_bytes = 0;
    CU(c, cudaHostAlloc(&c->h_stage, stage_bytes, cudaHostAllocDefault));
    c->h_stage_bytes = stage_bytes;
  }
  // slices of at least 2^20 photons (enough to fill the persistent grid several times over), at most WORK_COUNTERS of
  // them, whole chunks each
  // (slice size swept on the bench workload, 1.05e7 photons, one box: 2^16 111 ms, 2^17 99, 2^18 90, 2^19 81, 2^20 78 --
  // every slice is a launch of the persistent grid with its own tail of long histories, which costs more than the
  // longer fill and drain of big slices; BRANSON_AOS_SLICE overrides)
  uint64_t min_slice = 1ull << 20;
  if (const char *e = getenv("BRANSON_AOS_SLICE")) { const long long v = atoll(e); if (v >= (1ll << 16)) min_slice = (uint64_t)v; }
  uint64_t m = std::max<uint64_t>(min_slice, (n + WORK_COUNTERS - 1) / WORK_COUNTERS);
  m = (m + AOS_CHUNK - 1) / AOS_CHUNK * AOS_CHUNK;
  const uint32_t n_slices = (uint32_t)((n + m - 1) / m);
  const uint64_t chunks_per_slice = m / AOS_CHUNK;
  const uint64_t n_chunks = (n + AOS_CHUNK - 1) / AOS_CHUNK;
  auto chunks_of_slice = [&](uint32_t j) {
    const uint64_t first = (uint64_t)j * chunks_per_slice;
    return std::min<uint64_t>(chunks_per_slice, n_chunks - first);
  };
  // every event of the call lives in one holder, destroyed on every return path
  struct EventBag {
    std::vector<cudaEvent_t> all;
    cudaError_t make(cudaEvent_t &e) {
      const cudaError_t r = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      if (r == cudaSuccess) all.push_back(e);
      return r;
    }
    ~EventBag() {
      for (cudaEvent_t e : all) cudaEventDestroy(e);
    }
  } bag;
  std::vector<cudaEvent_t> ev_in(n_slices), ev_done(n_slices);
  std::vector<cudaEvent_t> ev_slot((size_t)2 * AOS_COPIERS * AOS_SLOTS);
  for (auto &e : ev_in) CU(c, bag.make(e));
  for (auto &e : ev_done) CU(c, bag.make(e));
  for (auto &e : ev_slot) CU(c, bag.make(e));
  cudaEvent_t ev_ready;
  CU(c, bag.make(ev_ready));
  uint8_t *d_aos = (uint8_t *)c->scr_aos.p;
  TransportParams P0 = make_params(c, true);
  if (prepare_tally_copies(c)) return 1;
  if (c->tally_copies_live > 1) {
    P0.tally_rep = (double2 *)c->scr_tally_rep.p;
    P0.tally_copies = c->tally_copies_live;
  }
  std::atomic<uint64_t> next_up{0}, next_down{0};
  std::vector<std::atomic<uint64_t>> uploaded(n_slices);  // chunks of the slice whose H2D copy has been enqueued
  for (auto &u : uploaded) u.store(0, std::memory_order_relaxed);
  std::atomic<uint32_t> issued{0};  // slices whose kernels (and ev_done) have been enqueued
  std::atomic<int> abort_flag{0};
  std::atomic<uint64_t> bad_rng{0};  // photons whose RNG seed / key-high words differ from the ctx's
  std::atomic<int> worker_err{(int)cudaSuccess};
  auto fail_worker = [&](cudaError_t e) {
    int expect = (int)cudaSuccess;
    worker_err.compare_exchange_strong(expect, (int)e);
    abort_flag.store(1, std::memory_order_release);
  };
  auto chunk_range = [&](uint64_t i, uint64_t &off, uint64_t &cnt) {
    off = i * AOS_CHUNK;
    cnt = std::min<uint64_t>(AOS_CHUNK, n - off);
  };
  // uploaders: vector -> pinned buffer -> device, chunks handed out in order
  auto uploader = [&](int t) {
    cudaError_t e = cudaSetDevice(c->device);
    uint8_t *buf[AOS_SLOTS];
    bool used[AOS_SLOTS] = {};
    for (int k = 0; k < AOS_SLOTS; ++k) buf[k] = (uint8_t *)c->h_stage + slot_bytes * (size_t)(t * AOS_SLOTS + k);
    for (int turn = 0; e == cudaSuccess; ++turn) {
      if (abort_flag.load(std::memory_order_acquire)) return;
      const uint64_t i = next_up.fetch_add(1, std::memory_order_relaxed);
      if (i >= n_chunks) break;
      const int k = turn % AOS_SLOTS;
      cudaEvent_t ev = ev_slot[t * AOS_SLOTS + k];
      if (used[k]) e = cudaEventSynchronize(ev);  // the buffer's previous copy has left it
      if (e != cudaSuccess) break;
      uint64_t off, cnt;
      chunk_range(i, off, cnt);
      memcpy(buf[k], photons + 120 * off, 120 * cnt);
      {
        // The RNG words the device does not carry (counter high word = seed << 32, key high word = 0; src/RNG.h:318-330)
        // are checked HERE, on the staged copy, before the chunk can reach the device: a slice holding a photon of
        // another seed is never transported, so it is never written back over the caller's vector (slices in front
        // of it, whose photons were valid, may already be final when the call fails).
        const uint64_t *w = (const uint64_t *)buf[k];
        uint64_t bad = 0;
        for (uint64_t q = 0; q < cnt; ++q) bad += (w[15 * q + 12] != c->ctr_hi) | (w[15 * q + 14] != 0ull);
        if (bad) {
          bad_rng.fetch_add(bad, std::memory_order_relaxed);
          abort_flag.store(1, std::memory_order_release);
          return;
        }
      }
      e = cudaMemcpyAsync(d_aos + 120 * off, buf[k], 120 * cnt, cudaMemcpyHostToDevice, c->s_in);
      if (e == cudaSuccess) e = cudaEventRecord(ev, c->s_in);
      used[k] = true;
      if (e == cudaSuccess) uploaded[i / chunks_per_slice].fetch_add(1, std::memory_order_release);
    }
    if (e != cudaSuccess) fail_worker(e);
  };
  // downloaders: device -> pinned buffer -> vector, once the chunk's slice has been transported
  auto downloader = [&](int t) {
    cudaError_t e = cudaSetDevice(c->device);
    uint8_t *buf[AOS_SLOTS];
    for (int k = 0; k < AOS_SLOTS; ++k)
      buf[k] = (uint8_t *)c->h_stage + slot_bytes * (size_t)((AOS_COPIERS + t) * AOS_SLOTS + k);
    uint64_t pend_off[AOS_SLOTS] = {}, pend_cnt[AOS_SLOTS] = {};
    bool pending[AOS_SLOTS] = {};
    auto drain = [&](int k) {
      if (!pending[k] || e != cudaSuccess) return;
      e = cudaEventSynchronize(ev_slot[(AOS_COPIERS + t) * AOS_SLOTS + k]);
      if (e == cudaSuccess) memcpy(photons + 120 * pend_off[k], buf[k], 120 * pend_cnt[k]);
      pending[k] = false;
    };
    for (int turn = 0; e == cudaSuccess; ++turn) {
      const uint64_t i = next_down.fetch_add(1, std::memory_order_relaxed);
      if (i >= n_chunks) break;
      const uint32_t j = (uint32_t)(i / chunks_per_slice);
      bool stop = false;
      while (issued.load(std::memory_order_acquire) <= j) {
        if (abort_flag.load(std::memory_order_acquire)) { stop = true; break; }
        std::this_thread::yield();
      }
      if (stop) break;  // (the copies already in flight are still drained into the caller's vector below)
      const int k = turn % AOS_SLOTS;
      drain(k);
      if (e != cudaSuccess) break;
      chunk_range(i, pend_off[k], pend_cnt[k]);
      cudaEvent_t ev = ev_slot[(AOS_COPIERS + t) * AOS_SLOTS + k];
      // (s_out is FIFO: one wait on the slice's event orders every copy queued behind it)
      e = cudaStreamWaitEvent(c->s_out, ev_done[j], 0);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(buf[k], d_aos + 120 * pend_off[k], 120 * pend_cnt[k], cudaMemcpyDeviceToHost, c->s_out);
      if (e == cudaSuccess) e = cudaEventRecord(ev, c->s_out);
      pending[k] = e == cudaSuccess;
    }
    for (int k = 0; k < AOS_SLOTS; ++k) drain(k);
    if (e != cudaSuccess) fail_worker(e);
  };
  const bool trace = getenv("BRANSON_AOS_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto ms_since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  double t_issued_all = 0.0;
  // The slices run on two compute streams in turn, each slice with its own work counter: a slice's photons come from one
  // part of the list (new photons are in cell order), so their histories are alike and its persistent grid drains
  // unevenly -- the next slice's CTAs take over the SMs as they fall idle.  (One stream: 94 ms of kernel time for the
  // 69 ms the unsliced list takes under the same profiler.)
  cudaError_t e = cudaSuccess;
  e = cudaEventRecord(ev_ready, c->stream);  // tallies uploaded, statistics and tally copies zeroed
  if (e == cudaSuccess) e = cudaStreamWaitEvent(c->s_k2, ev_ready, 0);
  std::vector<std::thread> threads;
  for (int t = 0; t < AOS_COPIERS; ++t) threads.emplace_back(uploader, t);
  for (int t = 0; t < AOS_COPIERS; ++t) threads.emplace_back(downloader, t);
  for (uint32_t j = 0; j < n_slices && e == cudaSuccess; ++j) {
    const uint64_t off = (uint64_t)j * m, cnt = std::min<uint64_t>(m, n - off);
    cudaStream_t sk = (j & 1u) ? c->s_k2 : c->stream;
    while (uploaded[j].load(std::memory_order_acquire) < chunks_of_slice(j)) {
      if (abort_flag.load(std::memory_order_acquire)) break;
      std::this_thread::yield();
    }
    if (abort_flag.load(std::memory_order_acquire)) break;
    // every H2D copy of the slice is in s_in's queue by now: an event recorded behind them covers them all
    e = cudaEventRecord(ev_in[j], c->s_in);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(sk, ev_in[j], 0);
    if (e != cudaSuccess) break;
    const PhotonSoA view = soa_view(c->work, off);
    c->launches += 2;  // + the history kernel, counted by launch_history
    k_aos_to_soa<<<grid_for(cnt, 128), 128, 0, sk>>>((const uint64_t *)(d_aos + 120 * off), cnt, view, c->ctr_hi,
                                                     c->d_stats);
    TransportParams P = P0;
    P.ph = view;
    P.n = cnt;
    P.desc = c->d_desc + off;
    if (P.counters) P.counters += 4 * off;
    P.work_counter = c->d_work_counter + j;
    P.chunk = chunk_for(c, cnt);
    e = cudaMemsetAsync(P.work_counter, 0, 8, sk);
    if (e == cudaSuccess && launch_history<TM_ATOMIC>(c, P, sk)) e = cudaErrorUnknown;
    k_soa_to_aos<<<grid_for(cnt, 128), 128, 0, sk>>>((uint64_t *)(d_aos + 120 * off), cnt, view, c->d_desc + off);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaEventRecord(ev_done[j], sk);
    if (e == cudaSuccess) issued.store(j + 1, std::memory_order_release);
  }
  // the caller's stream goes on (tally fold, tally download) once the other compute stream has finished its slices
  if (e == cudaSuccess && n_slices > 1) {
    const uint32_t last_odd = (n_slices - 1) | 1u;
    e = cudaStreamWaitEvent(c->stream, ev_done[last_odd < n_slices ? last_odd : last_odd - 2], 0);
  }
  if (e != cudaSuccess) abort_flag.store(1, std::memory_order_release);
  t_issued_all = ms_since();
  if (trace) {
    cudaStreamSynchronize(c->s_in);
    const double t_up = ms_since();
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->s_k2);
    const double t_k = ms_since();
    for (auto &th : threads) th.join();
    fprintf(stderr, "[aos] %u slices, %d copiers: last slice issued %.1f ms, uploads done %.1f, kernels done %.1f, downloads done %.1f\n",
            n_slices, AOS_COPIERS, t_issued_all, t_up, t_k, ms_since());
    threads.clear();
  }
  for (auto &th : threads) th.join();
  if (e == cudaSuccess && !abort_flag.load() && c->tally_copies_live > 1) {
    ++c->launches;
    k_fold_tally<<<grid_for(c->mesh.n_cells, 256), 256, 0, c->stream>>>((double2 *)c->d_tally, P0.tally_rep,
                                                                          c->mesh.n_cells, c->tally_copies_live - 1);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_in);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_out);
  *bad_out = bad_rng.load();
  if (*bad_out) {  // reported by the caller; the streams are drained first
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->s_k2);
    return 0;
  }
  if (e != cudaSuccess) return fail(c, "bgpu_transport_photons_aos: %s", cudaGetErrorString(e));
  if (worker_err.load() != (int)cudaSuccess)
    return fail(c, "bgpu_transport_photons_aos (copy thread): %s", cudaGetErrorString((cudaError_t)worker_err.load()));
  return 0;
}

}  // namespace

int bgpu_transport_photons_aos(bgpu_ctx *c, void *photons, uint64_t n, void *cell_tallies, int algorithm,
                               int tally_mode) {
  if (!c || (!photons && n) || !cell_tallies) return fail(c, "bgpu_transport_photons_aos: null argument");
  CU(c, cudaSetDevice(c->device));
  const uint64_t nc = c->mesh.n_cells;
  if (n >= (1ull << 32)) return fail(c, "bgpu_transport_photons_aos: too many photons");
  if (!c->have_cell_data) return fail(c, "bgpu_transport: cell data not set (call bgpu_set_cell_data first)");
  if (ensure_work(c, n, 0)) return 1;
  if (ensure(c, c->scr_aos, 120 * n)) return 1;
  CU(c, cudaMemcpyAsync(c->d_tally, cell_tallies, 16 * nc, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemsetAsync(c->d_stats, 0, 8 * ST_COUNT, c->stream));
  c->n_work = n;
  c->n_new = n;
  const bool pipelined = algorithm == BGPU_HISTORY && tally_mode == BGPU_TALLY_ATOMIC && n >= (1ull << 21);
  if (n && pipelined) {
    uint64_t bad = 0;
    if (transport_aos_pipelined(c, (uint8_t *)photons, n, &bad)) return 1;
    if (bad)
      return fail(c, "bgpu_transport_photons_aos: %llu photons carry an RNG seed/spawn word different from the ctx seed "
                     "(their slices were not transported)", (unsigned long long)bad);
  } else if (n) {
    CU(c, cudaMemcpyAsync(c->scr_aos.p, photons, 120 * n, cudaMemcpyHostToDevice, c->stream));
    ++c->launches;
    k_aos_to_soa<<<grid_for(n, 128), 128, 0, c->stream>>>((const uint64_t *)c->scr_aos.p, n, c->work, c->ctr_hi,
                                                          c->d_stats);
    unsigned long long bad = 0;
    CU(c, cudaMemcpyAsync(&bad, c->d_stats + ST_BAD_RNG, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (bad)
      return fail(c, "bgpu_transport_photons_aos: %llu photons carry an RNG seed/spawn word different from the ctx seed",
                  bad);
    if (run_transport(c, algorithm, tally_mode, true)) return 1;
    ++c->launches;
    k_soa_to_aos<<<grid_for(n, 128), 128, 0, c->stream>>>((uint64_t *)c->scr_aos.p, n, c->work, c->d_desc);
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(photons, c->scr_aos.p, 120 * n, cudaMemcpyDeviceToHost, c->stream));
  }
  CU(c, cudaMemcpyAsync(cell_tallies, c->d_tally, 16 * nc, cudaMemcpyDeviceToHost, c->stream));
  unsigned long long st[ST_COUNT] = {};
  CU(c, cudaMemcpyAsync(st, c->d_stats, 8 * ST_COUNT, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  // this call's statistics (no post-processing here: census / exit sums are the caller's, src/replicated_transport.h)
  c->stats = bgpu_cycle_stats{};
  c->stats.n_new = c->stats.n_transported = n;
  c->stats.n_events = st[ST_EVENTS];
  c->stats.n_scatters = st[ST_SCATTERS];
  c->stats.n_crossings = st[ST_CROSSINGS];
  c->stats.n_reflections = st[ST_REFLECTIONS];
  c->stats.n_deposits = st[ST_DEPOSITS];
  c->stats.n_group_lookups = st[ST_LOOKUPS];
  c->stats.n_launches = c->launches;
  return 0;
}

int bgpu_census_energy(bgpu_ctx *c, double *census_E) {
  if (!c || !census_E) return fail(c, "bgpu_census_energy: null argument");
  CU(c, cudaSetDevice(c->device));
  return list_energy(c, c->census, 0, c->n_census, census_E);
}

int bgpu_comb_census(bgpu_ctx *c, uint64_t max_census_photons, double global_census_E, uint64_t rng_stream,
                     bgpu_comb_stats *out) {
  if (!c) return 1;
  if (max_census_photons == 0) return fail(c, "bgpu_comb_census: max_census_photons must be positive");
  CU(c, cudaSetDevice(c->device));
  bgpu_comb_stats st{};
  const uint64_t n = c->n_census;
  st.n_before = st.n_after = n;
  st.rng_draws = n;
  if (list_energy(c, c->census, 0, n, &st.E_before)) return 1;
  st.E_after = st.E_before;
  if (!(global_census_E > 0.0)) global_census_E = st.E_before;
  st.comb_photon_E = global_census_E / (double)(int64_t)max_census_photons;  // (:65)
  if (n == 0) {
    if (out) *out = st;
    return 0;
  }
  if (n >= (1ull << 32)) return fail(c, "bgpu_comb_census: %llu census photons (limit 2^32 - 1)", (unsigned long long)n);
  const uint32_t nc = c->mesh.n_cells;
  if (ensure(c, c->scr_dep_cell, 4 * n) || ensure(c, c->scr_vals_in, 4 * n) || ensure(c, c->scr_keys_out, 4 * n) ||
      ensure(c, c->scr_vals_out, 4 * n) || ensure(c, c->scr_ndep, 4 * n) || ensure(c, c->scr_dep_off, 8 * (n + 1)) ||
      ensure(c, c->scr_seg, 16ull * nc) || ensure(c, c->scr_comb, 8ull * nc))
    return 1;
  uint32_t *keys = (uint32_t *)c->scr_dep_cell.p, *idx = (uint32_t *)c->scr_vals_in.p;
  uint32_t *keys_out = (uint32_t *)c->scr_keys_out.p, *order = (uint32_t *)c->scr_vals_out.p;
  uint32_t *keep = (uint32_t *)c->scr_ndep.p;
  uint64_t *offset = (uint64_t *)c->scr_dep_off.p;
  uint64_t *seg_start = (uint64_t *)c->scr_seg.p, *seg_end = seg_start + nc;
  double *new_E = (double *)c->scr_comb.p;
  ++c->launches;
  k_comb_draw<<<grid_for(n, 256), 256, 0, c->stream>>>(c->census.ee, c->census.sg, n, st.comb_photon_E, c->ctr_hi,
                                                       rng_stream, keys, idx, keep);
  int end_bit = 1;
  while ((1ull << end_bit) < nc) ++end_bit;
  size_t tmp_bytes = 0;
  CU(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, idx, order, (uint64_t)n, 0, end_bit, c->stream));
  if (ensure(c, c->scr_sort, tmp_bytes)) return 1;
  CU(c, cub::DeviceRadixSort::SortPairs(c->scr_sort.p, tmp_bytes, keys, keys_out, idx, order, (uint64_t)n, 0, end_bit,
                                        c->stream));
  CU(c, cudaMemsetAsync(c->scr_seg.p, 0, 16ull * nc, c->stream));
  c->launches += 3;
  k_seg_bounds<<<grid_for(n, 256), 256, 0, c->stream>>>(keys_out, n, seg_start, seg_end);
  k_comb_cells<<<grid_for(nc, 128), 128, 0, c->stream>>>(nc, seg_start, seg_end, order, c->census.ee, st.comb_photon_E,
                                                         keep, new_E);
  CU(c, cudaGetLastError());
  if (device_scan(c, keep, n, offset)) return 1;
  uint64_t kept = 0;
  CU(c, cudaMemcpyAsync(&kept, offset + n, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (ensure_soa(c, c->comb_scratch, kept, 0)) return 1;
  ++c->launches;
  k_comb_gather<<<grid_for(n, 256), 256, 0, c->stream>>>(c->census, n, keep, offset, new_E, c->comb_scratch);
  CU(c, cudaGetLastError());
  std::swap(c->census, c->comb_scratch);
  c->n_census = kept;
  st.n_after = kept;
  if (list_energy(c, c->census, 0, kept, &st.E_after)) return 1;
  if (out) *out = st;
  return 0;
}

int bgpu_sort_census_by_cell(bgpu_ctx *c) {
  if (!c) return 1;
  CU(c, cudaSetDevice(c->device));
  const uint64_t n = c->n_census;
  if (n < 2) return 0;
  if (n >= (1ull << 32)) return fail(c, "bgpu_sort_census_by_cell: %llu census photons (limit 2^32 - 1)", (unsigned long long)n);
  if (ensure(c, c->scr_dep_cell, 4 * n) || ensure(c, c->scr_vals_in, 4 * n) || ensure(c, c->scr_keys_out, 4 * n) ||
      ensure(c, c->scr_vals_out, 4 * n))
    return 1;
  uint32_t *keys = (uint32_t *)c->scr_dep_cell.p, *idx = (uint32_t *)c->scr_vals_in.p;
  uint32_t *keys_out = (uint32_t *)c->scr_keys_out.p, *order = (uint32_t *)c->scr_vals_out.p;
  ++c->launches;
  k_cell_keys<<<grid_for(n, 256), 256, 0, c->stream>>>(c->census.sg, n, keys, idx);
  int end_bit = 1;
  while ((1ull << end_bit) < c->mesh.n_cells) ++end_bit;
  size_t tmp_bytes = 0;
  CU(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, idx, order, (uint64_t)n, 0, end_bit, c->stream));
  if (ensure(c, c->scr_sort, tmp_bytes)) return 1;
  CU(c, cub::DeviceRadixSort::SortPairs(c->scr_sort.p, tmp_bytes, keys, keys_out, idx, order, (uint64_t)n, 0, end_bit,
                                        c->stream));
  if (ensure_soa(c, c->comb_scratch, n, 0)) return 1;
  c->launches += 2;
  k_permute_soa<<<grid_for(n, 256), 256, 0, c->stream>>>(c->census, n, order, c->comb_scratch);
  CU(c, cudaGetLastError());
  std::swap(c->census, c->comb_scratch);
  return 0;
}

uint64_t bgpu_list_size(const bgpu_ctx *c, int which) {
  if (!c) return 0;
  return which == BGPU_LIST_CENSUS ? c->n_census : c->n_work;
}

int bgpu_enable_counters(bgpu_ctx *c, int on) {
  if (!c) return 1;
  CU(c, cudaSetDevice(c->device));
  c->counters_on = on != 0;
  return ensure_work(c, c->work.cap, c->n_work);
}

int bgpu_set_launch(bgpu_ctx *c, int block_threads, int blocks_per_sm, int chunk) {
  if (!c) return 1;
  if (block_threads) {
    if (block_threads != 128) return fail(c, "bgpu_set_launch: the transport kernel is built for 128-thread CTAs");
    c->block_threads = block_threads;
  }
  c->blocks_per_sm = blocks_per_sm;
  if (chunk > 0) { c->chunk = (uint32_t)chunk; c->chunk_auto = false; }
  return 0;
}

int bgpu_set_divergence(bgpu_ctx *c, int scatter_batch, int aggregate_deposits) {
  if (!c) return 1;
  if (scatter_batch > 32) return fail(c, "bgpu_set_divergence: scatter_batch is a lane count (1..32, 0 = keep)");
  if (scatter_batch > 0) { c->scatter_batch = (uint32_t)scatter_batch; c->scatter_batch_auto = false; }
  if (aggregate_deposits >= 0) c->aggregate = aggregate_deposits ? 1 : 0;
  return 0;
}

int bgpu_set_tally_copies(bgpu_ctx *c, int copies) {
  if (!c) return 1;
  if (copies < 0 || copies > 1024) return fail(c, "bgpu_set_tally_copies: 0 (auto), 1 (off) .. 1024");
  c->tally_copies = copies;
  return 0;
}

int bgpu_set_group_walk(bgpu_ctx *c, int closed_form) {
  if (!c) return 1;
  c->closed_form_walk = closed_form != 0;
  return 0;
}

int bgpu_set_event_tail(bgpu_ctx *c, uint64_t n_active) {
  if (!c) return 1;
  c->event_tail = n_active;
  return 0;
}

int bgpu_set_kernel(bgpu_ctx *c, int choice) {
  if (!c) return 1;
  if (choice < 0 || choice > 2) return fail(c, "bgpu_set_kernel: 0 (auto), 1 (history kernel), 2 (event queues)");
  c->kernel_choice = choice;
  return 0;
}

int bgpu_set_event_mode(bgpu_ctx *c, int hbm_passes, int batch_scatter, int batch_refill) {
  if (!c) return 1;
  if (batch_scatter > 32 || batch_refill > 32) return fail(c, "bgpu_set_event_mode: batches are lane counts (1..32, 0 = keep)");
  if (hbm_passes >= 0) c->event_hbm = hbm_passes ? 1 : 0;
  if (batch_scatter > 0) c->pool_batch_scatter = (uint32_t)batch_scatter;
  if (batch_refill > 0) c->pool_batch_refill = (uint32_t)batch_refill;
  return 0;
}

int bgpu_upload_photons(bgpu_ctx *c, int which, const bgpu_photon_soa *h) {
  if (!c || !h) return fail(c, "bgpu_upload_photons: null argument");
  CU(c, cudaSetDevice(c->device));
  const uint64_t n = h->n;
  if (n && (!h->cell || !h->group || !h->pos || !h->angle || !h->E || !h->E0 || !h->life_dx || !h->ctr || !h->stream))
    return fail(c, "bgpu_upload_photons: every state field is required");
  PhotonSoA *dst;
  if (which == BGPU_LIST_CENSUS) {
    if (ensure_soa(c, c->census, n, 0)) return 1;
    dst = &c->census;
    c->n_census = n;
  } else {
    if (ensure_work(c, n, 0)) return 1;
    dst = &c->work;
    c->n_work = n;
    c->n_new = n;
  }
  if (!n) return 0;
  std::vector<double2> a(n);
  std::vector<ulonglong2> u(n);
  auto put = [&](void *d, const void *s) { return cudaMemcpy(d, s, 16 * n, cudaMemcpyHostToDevice); };
  for (uint64_t i = 0; i < n; ++i) a[i] = make_double2(h->pos[3 * i], h->pos[3 * i + 1]);
  CU(c, put(dst->xy, a.data()));
  for (uint64_t i = 0; i < n; ++i) a[i] = make_double2(h->pos[3 * i + 2], h->angle[3 * i]);
  CU(c, put(dst->za, a.data()));
  for (uint64_t i = 0; i < n; ++i) a[i] = make_double2(h->angle[3 * i + 1], h->angle[3 * i + 2]);
  CU(c, put(dst->bc, a.data()));
  for (uint64_t i = 0; i < n; ++i) a[i] = make_double2(h->E[i], h->E0[i]);
  CU(c, put(dst->ee, a.data()));
  for (uint64_t i = 0; i < n; ++i) {
    unsigned long long bits;
    memcpy(&bits, &h->life_dx[i], 8);
    u[i] = make_ulonglong2(bits, h->ctr[i]);
  }
  CU(c, put(dst->lc, u.data()));
  for (uint64_t i = 0; i < n; ++i)
    u[i] = make_ulonglong2(h->stream[i], (unsigned long long)h->cell[i] | ((unsigned long long)h->group[i] << 32));
  CU(c, put(dst->sg, u.data()));
  return 0;
}

int bgpu_download_photons(bgpu_ctx *c, int which, bgpu_photon_soa *h) {
  if (!c || !h) return fail(c, "bgpu_download_photons: null argument");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  const PhotonSoA &src = (which == BGPU_LIST_CENSUS) ? c->census : c->work;
  const uint64_t n = (which == BGPU_LIST_CENSUS) ? c->n_census : c->n_work;
  if (h->n < n) return fail(c, "bgpu_download_photons: host arrays hold %llu photons, list has %llu",
                            (unsigned long long)h->n, (unsigned long long)n);
  h->n = n;
  if (!n) return 0;
  std::vector<double2> a(n);
  std::vector<ulonglong2> u(n);
  auto get = [&](void *d, const void *s) { return cudaMemcpy(d, s, 16 * n, cudaMemcpyDeviceToHost); };
  if (h->pos) {
    CU(c, get(a.data(), src.xy));
    for (uint64_t i = 0; i < n; ++i) { h->pos[3 * i] = a[i].x; h->pos[3 * i + 1] = a[i].y; }
  }
  if (h->pos || h->angle) {
    CU(c, get(a.data(), src.za));
    for (uint64_t i = 0; i < n; ++i) {
      if (h->pos) h->pos[3 * i + 2] = a[i].x;
      if (h->angle) h->angle[3 * i] = a[i].y;
    }
  }
  if (h->angle) {
    CU(c, get(a.data(), src.bc));
    for (uint64_t i = 0; i < n; ++i) { h->angle[3 * i + 1] = a[i].x; h->angle[3 * i + 2] = a[i].y; }
  }
  if (h->E || h->E0) {
    CU(c, get(a.data(), src.ee));
    for (uint64_t i = 0; i < n; ++i) {
      if (h->E) h->E[i] = a[i].x;
      if (h->E0) h->E0[i] = a[i].y;
    }
  }
  if (h->life_dx || h->ctr) {
    CU(c, get(u.data(), src.lc));
    for (uint64_t i = 0; i < n; ++i) {
      if (h->life_dx) memcpy(&h->life_dx[i], &u[i].x, 8);
      if (h->ctr) h->ctr[i] = u[i].y;
    }
  }
  if (h->stream || h->cell || h->group) {
    CU(c, get(u.data(), src.sg));
    for (uint64_t i = 0; i < n; ++i) {
      if (h->stream) h->stream[i] = u[i].x;
      if (h->cell) h->cell[i] = (uint32_t)u[i].y;
      if (h->group) h->group[i] = (uint32_t)(u[i].y >> 32);
    }
  }
  if (which == BGPU_LIST_WORK) {
    if (h->descriptor) CU(c, cudaMemcpy(h->descriptor, c->d_desc, n, cudaMemcpyDeviceToHost));
    if (h->counters) {
      if (!c->counters_on || !c->d_counters) return fail(c, "bgpu_download_photons: counters were not enabled");
      CU(c, cudaMemcpy(h->counters, c->d_counters, 16 * n, cudaMemcpyDeviceToHost));
    }
  }
  return 0;
}

// known-answer hook for the RNG unit tests: out[i] = i-th draw of RNG(seed, stream) (src/RNG.h:318-330)
__global__ void k_rng_draws(uint32_t seed, uint64_t stream, uint32_t n, double *out) {
  if (threadIdx.x || blockIdx.x) return;
  uint64_t ctr = 0;
  for (uint32_t i = 0; i < n; ++i) out[i] = rng_next(ctr, ((uint64_t)seed) << 32, stream);
}
__global__ void k_threefry_kat(const uint64_t *in, uint64_t *out) {
  if (threadIdx.x || blockIdx.x) return;
  threefry2x64_20(in, in + 2, out);
}
// accuracy hook for fastmath.cuh: which = 0 exp, 1 log, 2 sincos (out = sin, out2 = cos), 3 libdevice sincos,
// 4 division: in[i] / in[i ^ 1] (neighbouring entries paired), 5 square root
__global__ void k_fastmath(int which, uint64_t n, const double *in, double *out, double *out2) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double x = in[i];
    if (which == 4) out[i] = ((i ^ 1) < n) ? fm_div(x, in[i ^ 1]) : 0.0;
    else if (which == 5) out[i] = fm_sqrt(x);
    else if (which == 0) out[i] = fm_exp_flush(x);
    else if (which == 1) out[i] = fm_log_pos(x);
    else if (which == 2) fm_sincos(x, &out[i], &out2[i]);
    else sincos(x, &out[i], &out2[i]);
  }
}
int bgpu_test_fastmath(int which, uint64_t n, const double *in, double *out, double *out2) {
  double *d = nullptr;
  if (cudaMalloc((void **)&d, 24ull * n) != cudaSuccess) return 1;
  cudaMemcpy(d, in, 8ull * n, cudaMemcpyHostToDevice);
  k_fastmath<<<296, 256>>>(which, n, d, d + n, d + 2 * n);
  cudaError_t e = cudaMemcpy(out, d + n, 8ull * n, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && out2) e = cudaMemcpy(out2, d + 2 * n, 8ull * n, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e != cudaSuccess;
}
int bgpu_test_rng_draws(uint32_t seed, uint64_t stream, uint32_t n, double *out) {
  double *d = nullptr;
  if (cudaMalloc((void **)&d, 8ull * n) != cudaSuccess) return 1;
  k_rng_draws<<<1, 32>>>(seed, stream, n, d);
  const cudaError_t e = cudaMemcpy(out, d, 8ull * n, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e != cudaSuccess;
}
int bgpu_test_threefry(const uint64_t ctr_key[4], uint64_t out[2]) {
  uint64_t *d = nullptr;
  if (cudaMalloc((void **)&d, 48) != cudaSuccess) return 1;
  cudaMemcpy(d, ctr_key, 32, cudaMemcpyHostToDevice);
  k_threefry_kat<<<1, 32>>>(d, d + 4);
  const cudaError_t e = cudaMemcpy(out, d + 4, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e != cudaSuccess;
}

}  // extern "C"
