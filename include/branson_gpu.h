/* branson_gpu.h -- C ABI of the B200 (sm_100a) IMC photon-transport hot path.
 *
 * This is the drop-in boundary for the part of lanl/branson that sits between
 * `imc_replicated_driver` (reference src/replicated_driver.h:64-88) and the
 * per-photon history loop (reference src/history_based_transport.h).  The
 * reference has no FFI layer: its "operator API" for this path is the set of
 * header calls
 *
 *   GPU_Setup(rank, n_ranks, use_gpu, cells)              src/gpu_setup.h:19-40
 *   make_initial_census_photons<Census_T>(...)             src/source.h:138-204
 *   make_photons<Census_T>(...)                            src/source.h:212-366
 *   join_photon_arrays(all, census)                        src/census_functions.h:21-29
 *   replicated_transport<Census_T>(...)                    src/replicated_transport.h:33-158
 *     -> gpu_transport_photons(off, photons, cells, tallies) src/history_based_transport.h:348-413
 *     -> post_process_photons(next_dt, ...)                src/post_process_functions.h:33-59
 *   get_photon_list_E(census)                              src/census_functions.h:31-46
 *
 * Each entry point below names the reference call it replaces.  Only plain C
 * types cross the boundary.  Conventions:
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from bgpu_last_error() (the reference prints and MPI_Aborts,
 *     src/config.h.in:86-91 -- the host driver does the same with our code);
 *   - a bgpu_ctx owns all device memory of one device for the whole run; the
 *     census stays resident in HBM between cycles;
 *   - a ctx is not thread safe; one ctx per device, one host thread per ctx;
 *   - there is NO CPU fallback: without a CUDA device bgpu_create fails.
 */
#ifndef BRANSON_GPU_H
#define BRANSON_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGPU_ABI_VERSION 2

/* Constants::bc_type, reference src/constants.h:28 */
enum { BGPU_REFLECT = 0, BGPU_VACUUM = 1, BGPU_ELEMENT = 2, BGPU_SOURCE = 3, BGPU_PROCESSOR = 4 };
/* Constants::event_type, reference src/constants.h:27 */
enum { BGPU_EXIT = 0, BGPU_PASS = 1, BGPU_CENSUS = 2, BGPU_SCATTER = 3, BGPU_KILLED = 4, BGPU_BOUND = 5 };
/* Constants::transport algorithm, reference src/constants.h (HISTORY / EVENT) */
enum { BGPU_HISTORY = 0, BGPU_EVENT = 1 };
/* tally reduction mode */
enum {
  BGPU_TALLY_ATOMIC = 0,       /* FP64 atomics (production) */
  BGPU_TALLY_DETERMINISTIC = 1 /* deposit log -> stable sort by cell -> in-order sum == the reference's serial
                                  photon-order summation (history_cpu_transport_photons with one thread) */
};

typedef struct bgpu_ctx bgpu_ctx;

/* Replaces the per-cycle GPU_Setup cell upload (src/gpu_setup.h:23-40).  The reference mesh is a tensor-product
 * grid (src/proto_mesh.h:106-218): cell = i + nx*(j + ny*k), cell bounds are the per-axis face arrays below,
 * neighbours are +-1 / +-nx / +-nx*ny and only domain faces carry a non-ELEMENT boundary condition. */
typedef struct {
  uint32_t abi_version;  /* = BGPU_ABI_VERSION */
  uint32_t n_groups;     /* BRANSON_N_GROUPS */
  uint32_t nx, ny, nz;
  const double *x_faces; /* [nx+1] */
  const double *y_faces; /* [ny+1] */
  const double *z_faces; /* [nz+1] */
  int32_t bc[6];         /* X_NEG X_POS Y_NEG Y_POS Z_NEG Z_POS */
  uint32_t seed;         /* IMC_Parameters::get_rng_seed */
  uint64_t n_user_photons;
  int32_t rank, n_ranks; /* replicated-mode rank of this device (stream offsets, src/source.h:144,221-222) */
  int32_t device;        /* CUDA device ordinal; <0: rank % n_devices (src/gpu_setup.h:68-78) */
  uint64_t photon_capacity; /* initial capacity hint (0 = 1.25 * n_user_photons / n_ranks); grows on demand */
} bgpu_mesh_desc;

/* per-cycle results: feeds IMC_State::set_* (src/replicated_transport.h:151-155, src/replicated_driver.h:61,71,80) */
typedef struct {
  double census_E;      /* post-transport census energy */
  double exit_E;
  double pre_census_E;  /* get_photon_list_E of the census entering this cycle */
  double new_photon_E;  /* sum of E0 over the photons make_photons created this cycle (exact-balance diagnostic) */
  uint64_t n_new;       /* photons created by make_photons this cycle */
  uint64_t n_transported; /* all_photons.size() */
  uint64_t n_census;    /* census_list.size() after transport */
  uint64_t n_killed, n_exit;
  /* exact event counts of the cycle (for the algorithmic-bytes roofline) */
  uint64_t n_events, n_scatters, n_crossings, n_reflections, n_deposits, n_group_lookups;
  uint64_t n_launches; /* kernels launched through this ctx since bgpu_create (cumulative) */
  /* device times (CUDA events on the ctx stream), milliseconds */
  float ms_source, ms_transport, ms_census, ms_total;
  uint32_t transport_kernel; /* which kernel ran the histories: 0 history, 1 event queues (shared memory), 2 event passes */
  uint32_t reserved;
} bgpu_cycle_stats;

/* host-side SoA view used by bgpu_upload_photons / bgpu_download_photons (validation and tests).  Any pointer may be
 * NULL on download to skip that field.  pos/angle are [n][3]. */
typedef struct {
  uint64_t n;
  uint32_t *cell, *group;
  double *pos, *angle, *E, *E0, *life_dx;
  uint64_t *ctr;    /* RNG counter low word == number of draws consumed (src/RNG.h:262-285) */
  uint64_t *stream; /* RNG key low word (src/RNG.h:318-330) */
  uint8_t *descriptor;
  uint32_t *counters; /* [n][4] events, scatters, cell crossings, reflections (only if counters were enabled) */
} bgpu_photon_soa;

int bgpu_device_count(void);
const char *bgpu_last_error(const bgpu_ctx *ctx); /* ctx may be NULL: error of the last failed bgpu_create */

/* GPU_Setup ctor (src/gpu_setup.h:19-40): picks the device, uploads the (static) geometry once. */
int bgpu_create(bgpu_ctx **out, const bgpu_mesh_desc *desc);
void bgpu_destroy(bgpu_ctx *ctx);

/* The per-cycle part of GPU_Setup: Mesh::calculate_photon_energy rewrites op_a/op_s/f every cycle
 * (src/mesh.h:253-270).  f, op_a, op_s are per-cell gray values [n_cells]; like Cell::set_op_a/set_op_s
 * (src/cell.h:260-275) every group of a cell receives the same value. */
int bgpu_set_cell_data(bgpu_ctx *ctx, const double *f, const double *op_a, const double *op_s);
/* General multigroup form: abs_groups / sct_groups [n_cells][n_groups] (src/cell.h:318-337). */
int bgpu_set_cell_groups(bgpu_ctx *ctx, const double *f, const double *abs_groups, const double *sct_groups);

/* make_initial_census_photons (cycle 1 only, pass E_census != NULL; src/source.h:138-204) + make_photons
 * (src/source.h:212-366) + join_photon_arrays (src/census_functions.h:21-29): after the call the device work list
 * is [new photons ..., census ...].  E_* are this rank's per-cell source energies [n_cells] from
 * Mesh::get_emission_E / get_source_E / get_census_E, total_E is the global source energy
 * (src/replicated_driver.h:56-59), cycle is IMC_State::get_step(). */
int bgpu_source(bgpu_ctx *ctx, uint32_t cycle, double dt, const double *E_emission, const double *E_source,
                const double *E_census_or_null, double total_E, uint64_t *n_new, uint64_t *n_total);

/* replicated_transport (src/replicated_transport.h:33-158) on the device work list: history loop, then
 * post_process_photons (census compaction with life_dx = c*next_dt, exit/census energy sums).  Tallies are zeroed at
 * entry (a fresh vector<Cell_Tally>, src/replicated_transport.h:71).
 * algorithm: BGPU_HISTORY | BGPU_EVENT; tally_mode: BGPU_TALLY_ATOMIC | BGPU_TALLY_DETERMINISTIC. */
int bgpu_transport(bgpu_ctx *ctx, double next_dt, int algorithm, int tally_mode);

/* Population control of the device census: comb_photons(census_photons, max_census_photons,
 * rng) of src/census_functions.h:48-93 -- a function the reference defines but never calls in this snapshot, so nothing
 * here runs unless the host asks for it.  In list order each census photon takes one draw of RNG(ctx seed, rng_stream)
 * and survives with probability E / (global_census_E / max_census_photons); a cell's last photon always survives;
 * survivors share their cell's energy so that every cell's census energy is conserved (:86-92).  global_census_E is the
 * census energy summed over all ranks (:61-64; bgpu_census_energy gives this rank's part; <= 0: use this rank's own).
 * Survivors keep their list order, positions, directions and RNG state.  Same survivors and energies as the reference
 * function on the same list (tests/test_gpu_comb.py against the oracle and the golden fixtures). */
typedef struct {
  uint64_t n_before, n_after; /* census size */
  double E_before, E_after;   /* census energy (fixed-tree sums) */
  double comb_photon_E;       /* global_census_E / max_census_photons */
  uint64_t rng_draws;         /* = n_before: one per photon */
} bgpu_comb_stats;
int bgpu_census_energy(bgpu_ctx *ctx, double *census_E);
int bgpu_comb_census(bgpu_ctx *ctx, uint64_t max_census_photons, double global_census_E, uint64_t rng_stream,
                     bgpu_comb_stats *stats_or_null);

/* Locality-aware photon order between cycles: stable sort of the device census by cell, so
 * that neighbouring lanes of the next cycle's transport work in neighbouring cells (cell data and tally lines shared,
 * more same-cell deposits for the warp aggregation).  The reference keeps the census in the order post_process_photons
 * appended it (src/post_process_functions.h:33-59); a photon's history depends on nothing but its own state (SURVEY
 * section 8a, N5), so the order changes no per-photon result -- only the order in which tallies are summed (and with
 * it the meaning of "the reference's serial order" in BGPU_TALLY_DETERMINISTIC).  Off unless the host asks for it. */
int bgpu_sort_census_by_cell(bgpu_ctx *ctx);

/* rank_abs_E / rank_track_E hand-off (src/replicated_transport.h:135-140); either pointer may be NULL. */
int bgpu_get_tallies(bgpu_ctx *ctx, double *abs_E, double *track_E, bgpu_cycle_stats *stats);

/* Packed device buffer for the end-of-cycle all-reduce (replaces the MPI_Allreduce calls of
 * src/replicated_driver.h:91-94): n_doubles doubles = interleaved Cell_Tally {abs_E, track_E}[n_cells] followed by
 * `extra` caller-owned doubles.  The caller all-reduces it in place (NCCL) and then calls bgpu_get_tallies. */
int bgpu_tally_buffer(bgpu_ctx *ctx, uint64_t extra, void **device_ptr, uint64_t *n_doubles);
/* ---- replicated-mode collectives, native (csrc/comm_native.cuh) -----------------------------------------------------
 * Replace the MPI calls of the reference's replicated cycle: MPI_Allreduce of abs_E / track_E
 * (src/replicated_driver.h:91-94), of the source energies (src/replicated_driver.h:56-59, src/mesh.h:291-294), of
 * m_emission_E (src/mesh.h:343-345) and the scalar reductions of IMC_State::print_conservation
 * (src/imc_state.h:207-252).  A rank is a bgpu_ctx (bgpu_mesh_desc.rank / n_ranks).  Two back ends:
 *   NCCL (one GPU per rank, NVLink / NVSwitch) -- several processes: rank 0 calls bgpu_comm_unique_id, the launcher
 *     carries the 128 bytes to every process (torchrun store, a file, MPI_Bcast ...), every rank calls
 *     bgpu_comm_init_rank; one process with a host thread per rank: bgpu_comm_init_local on the array of contexts;
 *   in-process (bgpu_comm_init_local with all ranks on ONE device -- the reference maps rank % n_devices,
 *     src/gpu_setup.h:68-78): the ranks' threads meet and one kernel forms the rank-ordered sum.
 * libnccl.so.2 is loaded on first use; a single-rank run never touches it.  A collective call blocks the calling host
 * thread only until every rank has made the same call (in-process back end) or not at all (NCCL): the reduction itself
 * is ordered on the ctx stream. */
#define BGPU_COMM_ID_BYTES 128
enum { BGPU_COMM_NONE = 0, BGPU_COMM_NCCL = 1, BGPU_COMM_LOCAL = 2 };
enum { BGPU_OP_SUM = 0, BGPU_OP_MAX = 1, BGPU_OP_MIN = 2 };
int bgpu_comm_unique_id(char id[BGPU_COMM_ID_BYTES]);                 /* ncclGetUniqueId */
int bgpu_comm_init_rank(bgpu_ctx *ctx, const char id[BGPU_COMM_ID_BYTES]); /* ncclCommInitRank(n_ranks, id, rank) */
int bgpu_comm_init_local(bgpu_ctx **ctxs, int n_ranks);               /* ctxs[r] = rank r, all in this process */
int bgpu_comm_info(const bgpu_ctx *ctx, int *kind, uint64_t *calls, uint64_t *bytes);
/* in-place reduction of n HOST doubles over the ranks (staged through the device; the host-mesh path's scalars) */
int bgpu_comm_allreduce_host(bgpu_ctx *ctx, double *buf, uint64_t n, int op);
/* The cycle's ONE collective: the packed buffer {abs_E, track_E}[n_cells] + tail[n_ranks][BGPU_RANK_SCALARS] is summed
 * in place over the ranks on the ctx stream.  Every rank contributes `rank_scalars` (BGPU_RANK_SCALARS doubles of the
 * caller's choosing: census / exit energy, photon counts below 2^53, transport time ...) in its own row of the tail and
 * zeros elsewhere, so that after the sum `all_scalars` [n_ranks][BGPU_RANK_SCALARS] holds every rank's values and sums,
 * maxima and minima over ranks can be formed locally in rank order.  With one rank: a copy. */
#define BGPU_RANK_SCALARS 12
int bgpu_comm_allreduce_tallies(bgpu_ctx *ctx, const double *rank_scalars, double *all_scalars);

/* cudaStreamSynchronize of the ctx stream / raw stream handle for collective plumbing */
int bgpu_sync(bgpu_ctx *ctx);
void *bgpu_stream(bgpu_ctx *ctx);
int bgpu_device(const bgpu_ctx *ctx);

/* Drop-in for gpu_transport_photons (src/history_based_transport.h:348-413): `photons` is the reference's host
 * std::vector<Photon>::data() (120-byte records, src/photon.h:171-182), `cell_tallies` its vector<Cell_Tally>::data()
 * (16-byte records, src/cell_tally.h:53-54).  Photons are updated in place exactly as transport_photon does and the
 * tallies are ACCUMULATED into cell_tallies.  Cell data must have been set with bgpu_set_cell_data. */
int bgpu_transport_photons_aos(bgpu_ctx *ctx, void *photons, uint64_t n_photons, void *cell_tallies,
                               int algorithm, int tally_mode);

/* validation / tests: replace or read the device work list (pre- or post-transport) and the census */
enum { BGPU_LIST_WORK = 0, BGPU_LIST_CENSUS = 1 };
int bgpu_upload_photons(bgpu_ctx *ctx, int which, const bgpu_photon_soa *host);
int bgpu_download_photons(bgpu_ctx *ctx, int which, bgpu_photon_soa *host);
uint64_t bgpu_list_size(const bgpu_ctx *ctx, int which);
/* Validation mode: per-photon event counters (bit-exact observable) are recorded and EVERY photon's final state is
 * written back (production writes the full state of CENSUS photons only, plus E and the descriptor of the rest). */
int bgpu_enable_counters(bgpu_ctx *ctx, int on);

/* tuning knobs (defaults chosen for B200: 148 SMs) */
int bgpu_set_launch(bgpu_ctx *ctx, int block_threads, int blocks_per_sm, int chunk_photons);
/* History kernel, divergence and tally-contention control (per-photon results do not depend on either; only the
 * summation order of the atomic tallies does): scatter_batch = lanes of a warp that must be parked at a scatter before
 * the warp samples them together (1 = sample immediately, 0 = keep the current value; default 12, or 6 when the previous
 * launch's histories averaged fewer than 16 events); aggregate_deposits =
 * 1 combines the same-cell deposits of a warp trip into one pair of atomics, 0 (default: with replicated tallies in place
 * the matching costs more than it saves) issues them per lane, < 0 keeps the current setting. */
int bgpu_set_divergence(bgpu_ctx *ctx, int scatter_batch, int aggregate_deposits);
/* Replicated tallies of the history kernel (atomic mode): warps deposit into `copies` separate tally arrays that are
 * summed into the tally buffer after the launch, which divides the same-address reduction traffic of hot cells.
 * 0 = auto (meshes below 2^17 cells: as many as fit 64 MB, at most 64; larger meshes: off), 1 = off. */
int bgpu_set_tally_copies(bgpu_ctx *ctx, int copies);
/* BGPU_EVENT: active-list size at or below which the lockstep passes hand the remaining histories to the persistent
 * history kernel (0 = auto: twice the number of resident lanes) */
int bgpu_set_event_tail(bgpu_ctx *ctx, uint64_t n_active);
/* BGPU_HISTORY: which kernel runs the histories.  0 (default) = auto: the event-queue kernel (see below) on decks that mix
 * event types -- work lists of >= 5e5 photons whose previous launch saw >= 16 events per history with 8..45 % of them scatters (big_cube, hot_zone), where
 * the history kernel loses a third of its lanes to divergence -- and the history kernel elsewhere; 1 = always the history
 * kernel; 2 = always the event queues.  Per-photon results do not depend on the choice (tests/test_gpu_parity.py). */
int bgpu_set_kernel(bgpu_ctx *ctx, int choice);
/* BGPU_EVENT has two forms.  Default (hbm_passes = 0): event queues in shared memory (csrc/pool.cuh) -- every lane owns
 * two photon slots, each trip the warp elects the event type most lanes can serve (advance / scatter / retire+refill)
 * and runs that block alone: the regrouping of the reference's event_based_transport.h at warp scope, without HBM
 * traffic.  batch_scatter / batch_refill = lanes that must hold a parked scatter / a finished or empty slot before that
 * block is elected over a fuller advance block (0 = keep; defaults 20 / 8).  hbm_passes = 1: the first form, lockstep
 * passes over active lists in HBM (csrc/event.cuh; bgpu_set_event_tail applies to it); < 0 keeps the current form. */
int bgpu_set_event_mode(bgpu_ctx *ctx, int hbm_passes, int batch_scatter, int batch_refill);
/* sample_emission_group (src/sampling_functions.h:126-138): 1 (default) = when every cell's groups are equal, use the
 * provably-equivalent closed form of the sequential walk; 0 = always walk the group array like the reference */
int bgpu_set_group_walk(bgpu_ctx *ctx, int closed_form);

/* ---- device-resident mesh physics (optional) ---------------------------------------------
 * Mesh::calculate_photon_energy (src/mesh.h:237-323) and Mesh::update_temperature (src/mesh.h:327-362) on the device:
 * T_e, T_r, opacities, Fleck factor, the emission / source / census energies and the tallies stay in HBM, only the
 * running sums come back.  Replaces the per-cycle bgpu_set_cell_data + bgpu_source(host arrays) + bgpu_get_tallies
 * sequence by
 *   bgpu_mesh_calculate_photon_energy -> [n_ranks > 1: all-reduce the scalar, bgpu_mesh_redistribute] ->
 *   bgpu_mesh_source -> bgpu_transport -> [all-reduce bgpu_tally_buffer] -> bgpu_mesh_update_temperature.
 * Same expressions in the same operand order as the reference; pow() and the order of the running sums differ from the
 * host path in the last bits (csrc/mesh_dev.cuh). */
typedef struct {
  double opac_A, opac_B, opac_C, opac_S, cV, rho; /* src/region.h */
} bgpu_region;
typedef struct {
  double pre_mat_E, emission_E, census_E, source_E, total_photon_E; /* calculate_photon_energy / redistribute */
  double absorbed_E, post_mat_E;                                    /* update_temperature */
} bgpu_mesh_sums;
/* static cell data: region table, region of every cell, initial T_e / T_r, source temperature T_s (0 without a
 * SOURCE face) -- initialize_physical_properties, src/mesh.h:426-440 */
int bgpu_mesh_init(bgpu_ctx *ctx, uint32_t n_regions, const bgpu_region *regions, const uint32_t *region_of_cell,
                   const double *T_e, const double *T_r, const double *T_s);
/* src/mesh.h:253-287: fills f / op_a / op_s (all groups) and this rank's E_emission / E_census / E_source */
int bgpu_mesh_calculate_photon_energy(bgpu_ctx *ctx, double dt, uint32_t step, bgpu_mesh_sums *sums);
/* src/mesh.h:291-315 (replicated runs with more than one rank): global_source_E is the all-reduced
 * emission + census + source total of bgpu_mesh_calculate_photon_energy; the four energy sums are recomputed */
int bgpu_mesh_redistribute(bgpu_ctx *ctx, double global_source_E, bgpu_mesh_sums *sums);
/* calculate_photon_energy + the replicated redistribution WITHOUT their collectives (src/mesh.h:291-294,
 * src/replicated_driver.h:56-59): every rank holds the same cell state and takes the same decisions, so this rank forms
 * the post-redistribution totals of every rank itself.  rank_sums [n_ranks]: rank r's emission / census / source /
 * total_photon_E (pre_mat_E is the same for all); the caller's global source energy is their rank-ordered sum.  With
 * one rank: bgpu_mesh_calculate_photon_energy. */
int bgpu_mesh_calculate_photon_energy_replicated(bgpu_ctx *ctx, double dt, uint32_t step, bgpu_mesh_sums *rank_sums);
/* End of a device-mesh cycle in one stream-ordered chain: bgpu_comm_allreduce_tallies -> update_temperature
 * (src/mesh.h:343-362) on the reduced tallies -> one host synchronisation that brings back `all_scalars` and the
 * absorbed / post-material energy sums. */
int bgpu_mesh_finish_cycle(bgpu_ctx *ctx, const double *rank_scalars, double *all_scalars, bgpu_mesh_sums *sums);
/* bgpu_source from the device-resident energies of this cycle */
int bgpu_mesh_source(bgpu_ctx *ctx, uint32_t cycle, double total_E, uint64_t *n_new, uint64_t *n_total);
/* src/mesh.h:343-362 from the tally buffer (all-reduced by the caller in multi-rank runs) */
int bgpu_mesh_update_temperature(bgpu_ctx *ctx, bgpu_mesh_sums *sums);
/* copy one per-cell array to the host: T_e T_r T_s f op_a op_s E_emission E_source E_census abs_E track_E */
int bgpu_mesh_get(bgpu_ctx *ctx, const char *name, double *out);

/* known-answer hooks for the RNG unit tests (RNG(seed, stream) draws, src/RNG.h:262-285,318-330; raw Threefry2x64-20
 * of {ctr0, ctr1, key0, key1}, src/random123/threefry.h:196-282) */
int bgpu_test_rng_draws(uint32_t seed, uint64_t stream, uint32_t n, double *out);
int bgpu_test_threefry(const uint64_t ctr_key[4], uint64_t out[2]);
/* accuracy hook for the loop's own log / exp / sincos (csrc/fastmath.cuh; they stand in for the libm calls of
 * src/history_based_transport.h:63,74 and src/sampling_functions.h:66-68): which = 0 exp, 1 log (positive normal
 * arguments), 2 sincos (out = sin, out2 = cos), 3 CUDA's own sincos for comparison */
int bgpu_test_fastmath(int which, uint64_t n, const double *in, double *out, double *out2);

#ifdef __cplusplus
}
#endif
#endif /* BRANSON_GPU_H */
