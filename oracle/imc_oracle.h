/* imc_oracle.h -- PARITY ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C CPU restatement of the lanl/branson replicated-mode IMC cycle
 * (source -> transport -> census/tally -> temperature update).  Every function
 * in imc_oracle.c cites the reference file:line it follows.  Pinned bit-for-bit
 * against the UNMODIFIED reference (oracle/_ref/ref_harness_g*) by
 * tests/test_oracle_vs_reference.py and by the committed fixtures under
 * tests/golden/ (generated with oracle/gen_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (branson_b200)
 * never links or imports it.
 */
#ifndef IMC_ORACLE_H
#define IMC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/constants.h:27-29 */
enum { ORC_REFLECT = 0, ORC_VACUUM = 1, ORC_ELEMENT = 2, ORC_SOURCE = 3, ORC_PROCESSOR = 4 };
enum { ORC_EXIT = 0, ORC_PASS = 1, ORC_CENSUS = 2, ORC_SCATTER = 3, ORC_KILLED = 4, ORC_BOUND = 5 };

typedef struct {
  double t_start, t_stop, dt_start, t_mult, dt_max;
  uint64_t n_photons;
  uint32_t seed;
  uint32_t n_groups;
  int32_t n_xdiv, n_ydiv, n_zdiv;
  const double *x_start, *x_end;
  const uint32_t *x_cells;
  const double *y_start, *y_end;
  const uint32_t *y_cells;
  const double *z_start, *z_end;
  const uint32_t *z_cells;
  const uint32_t *div_region; /* [n_zdiv][n_ydiv][n_xdiv] user region IDs */
  int32_t bc[6];              /* X_NEG X_POS Y_NEG Y_POS Z_NEG Z_POS */
  double T_source;
  int32_t n_regions;
  const uint32_t *region_id;  /* [n_regions] */
  const double *region_props; /* [n_regions][8]: density cV opacA opacB opacC opacS T_e T_r */
  int32_t n_ranks;            /* emulated MPI ranks (replicated mode) */
} orc_problem;

typedef struct orc_sim orc_sim;

orc_sim *orc_create(const orc_problem *p);
void orc_destroy(orc_sim *s);
/* 1 if the simulation reached t_stop (reference src/imc_state.h:127-134) */
int orc_finished(const orc_sim *s);
/* run one cycle for all emulated ranks; keep_photons!=0 keeps the per-photon
 * pre/post arrays of that cycle queryable through orc_get */
int orc_cycle(orc_sim *s, int keep_photons);
/* Look up a named result array of emulated rank `rank` for the last cycle.
 * dtype: 0=f64 1=u32 2=u64 3=u8.  Returns 0 if found. */
int orc_get(const orc_sim *s, int rank, const char *name, const void **data, uint64_t *count, int *dtype);
double orc_last_transport_seconds(const orc_sim *s);

/* ---- stand-alone pieces for unit parity tests ---- */
/* reference random123/threefry.h:196-282 (Threefry2x64, 20 rounds) */
void orc_threefry2x64_20(const uint64_t ctr[2], const uint64_t key[2], uint64_t out[2]);
/* reference RNG.h:262-285: state = {ctr_lo, ctr_hi, key0, key1}; advances state */
double orc_rng_next(uint64_t state[4]);
/* reference RNG.h:318-330 */
void orc_rng_init(uint64_t state[4], uint32_t seed, uint64_t stream);
/* reference sampling_functions.h:57-70 */
void orc_uniform_angle(uint64_t state[4], double angle[3]);
/* reference cell.h:116-132 */
double orc_distance_to_boundary(const double nodes[6], const double pos[3], const double angle[3], uint32_t *surface);

/* Transport a caller-provided photon list on the current mesh state of `rank`
 * (reference history_based_transport.h:32-141), photons in SoA arrays updated
 * in place; tallies (abs_E/track_E, n_cells each) accumulated serially in
 * photon order.  counters (optional) = [n][4] u32: events, scatters, cell
 * crossings, reflections. */
int orc_transport_list(orc_sim *s, int rank, uint64_t n, uint32_t *cell, uint32_t *group, double *pos /*[n][3]*/,
                       double *angle /*[n][3]*/, double *E, const double *E0, double *life_dx, uint64_t *ctr,
                       const uint64_t *stream, uint8_t *descriptor, double *abs_E, double *track_E,
                       uint32_t *counters);

/* Population control of the census: reference census_functions.h:48-93 (comb_photons; defined but never called in
 * this snapshot).  cell[n], E[n] = the census in list order; global_census_E = the all-reduced census energy (:61-64;
 * for one rank: the in-order sum of E); rng_state advances by one draw per photon (:76).  Outputs: keep[n] = 1 for the
 * photons that survive (their list order is preserved, :79), new_E[n] = their corrected energy (:86-90).  Returns the
 * number kept. */
uint64_t orc_comb_photons(uint64_t n, const uint32_t *cell, const double *E, double global_census_E,
                          int64_t max_census_photons, uint64_t rng_state[4], uint8_t *keep, double *new_E);

#ifdef __cplusplus
}
#endif
#endif
