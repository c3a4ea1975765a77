/* branson_host.h -- C API of the C++ host layer (Input / IMC_Parameters / IMC_State / Mesh / replicated driver).
 *
 * The host layer is C++ (headers under branson_b200/csrc/host) and mirrors the reference's classes; this flat API exists so
 * that tests and bench.py can step the replicated driver (reference src/replicated_driver.h:33-122) one cycle at a
 * time and read its state, and so that a torchrun harness can supply the collectives (csrc/host/comm.h).
 * Returns 0 on success; bhost_last_error() gives the message (the reference prints and exits / MPI_Aborts).
 */
#ifndef BRANSON_HOST_H
#define BRANSON_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "branson_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bhost_driver bhost_driver;

#ifndef BHOST_COMM_DEFINED
#define BHOST_COMM_DEFINED
typedef struct {
  void *user;
  int (*allreduce_sum_f64)(void *user, double *buf, uint64_t n);
  int (*allreduce_sum_f64_device)(void *user, void *device_ptr, uint64_t n, void *stream);
  int (*allreduce_sum_u64)(void *user, uint64_t *buf, uint64_t n);
  int (*allreduce_max_f64)(void *user, double *buf, uint64_t n);
  int (*allreduce_min_f64)(void *user, double *buf, uint64_t n);
  int (*barrier)(void *user);
} bhost_comm;
#endif

typedef struct {
  uint32_t n_groups;        /* BRANSON_N_GROUPS of the reference build this run stands in for (compile-time there) */
  int32_t device;           /* CUDA device; <0: rank % n_devices */
  int32_t tally_mode;       /* BGPU_TALLY_ATOMIC | BGPU_TALLY_DETERMINISTIC */
  int32_t algorithm;        /* -1: from the deck (particle_algorithm); else BGPU_HISTORY | BGPU_EVENT */
  int32_t print;            /* 1: the reference's per-cycle stdout report */
  int32_t validate;         /* 1: per-photon event counters + full write-back (bgpu_enable_counters) */
  int32_t no_gpu;           /* 1: host logic only (Input/Mesh/IMC_State; CPU tests of partitioning) -- cycles unavailable */
  uint64_t photons_override;  /* 0: keep the deck's <photons> */
  double t_stop_override;     /* <=0: keep the deck's <t_stop> */
  int32_t force_replicated;   /* 1: treat dd_transport_type as REPLICATED (the multi-node deck says PARTICLE_PASS) */
  int32_t mesh_on_device;     /* 1: calculate_photon_energy / update_temperature on the device (bgpu_mesh_*); 0: host Mesh */
  uint64_t comb_max_census;   /* > 0: comb the census down to about this many photons whenever it exceeds that after a
                                 cycle (bgpu_comb_census); 0: never (the reference's driver never calls its comb) */
  int32_t sort_census;        /* 1: sort the census by cell after every cycle (bgpu_sort_census_by_cell); 0: keep the
                                 reference's order */
} bhost_options;

typedef struct {
  uint32_t step;
  double dt, time, next_dt, global_source_energy;
  bgpu_cycle_stats gpu;
  double t_calc_energy, t_cell_upload, t_source, t_transport, t_allreduce, t_tally_download, t_update_T, t_cycle;
  /* IMC_State after print_conservation (global sums) */
  double absorbed_E, emission_E, source_E, pre_census_E, post_census_E, pre_mat_E, post_mat_E, exit_E;
  double rad_conservation, mat_conservation;
  /* the same radiation balance from compensated sums of what the device actually made and tallied (see
   * Replicated_Driver::cycle): independent of the serial cell-order sum the reference formula uses */
  double rad_balance_exact;
  uint64_t trans_particles, census_size;
  uint64_t comb_n_before, comb_n_after; /* global census size around this cycle's comb; both 0: no comb ran */
} bhost_cycle_report;

bhost_driver *bhost_create(const char *xml_path, int rank, int n_ranks, const bhost_options *opt,
                           const bhost_comm *comm /* NULL: single rank */, char *err, size_t err_len);
void bhost_destroy(bhost_driver *d);
/* Multi-rank runs: the native communicator of the ranks' device contexts (bgpu_comm_*, include/branson_gpu.h) -- NCCL
 * between GPUs, an in-process sum between ranks that share one.  Several processes (one per GPU): rank 0 calls
 * bhost_comm_unique_id, the launcher carries the id to every process, every rank calls bhost_comm_init_rank.  One
 * process holding all ranks (one host thread per rank steps its driver): bhost_comm_init_local(drivers, n_ranks).
 * After either, bhost_cycle needs no `comm` callbacks. */
int bhost_comm_unique_id(char id[BGPU_COMM_ID_BYTES]);
int bhost_comm_init_rank(bhost_driver *d, const char id[BGPU_COMM_ID_BYTES]);
int bhost_comm_init_local(bhost_driver **drivers, int n_ranks);
const char *bhost_last_error(const bhost_driver *d);
int bhost_finished(const bhost_driver *d);
/* Mesh::calculate_photon_energy only (host; used by the CPU tests of the rank partitioning) */
int bhost_calculate_photon_energy(bhost_driver *d, double *global_source_energy);
int bhost_cycle(bhost_driver *d, bhost_cycle_report *out);
/* IMC_State::next_time_step alone (src/imc_state.h:292-296); host-only tests of the time stepping */
int bhost_next_time_step(bhost_driver *d);
/* named host arrays: T_e T_r T_s f op_a op_s E_emission E_source E_census abs_E track_E x_faces y_faces z_faces */
int bhost_get_array(const bhost_driver *d, const char *name, const double **data, uint64_t *n);
/* scalars: n_cells nx ny nz n_user_photons seed dd_mode particle_algorithm batch_size n_omp_threads ... */
int bhost_get_param(const bhost_driver *d, const char *name, double *value);
bgpu_ctx *bhost_gpu_ctx(bhost_driver *d);
double bhost_total_transport_time(const bhost_driver *d);

#ifdef __cplusplus
}
#endif
#endif
