/* imc_oracle.c -- PARITY ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the lanl/branson replicated-mode IMC cycle.  Each
 * function cites the reference file:line (under /root/reference/src) whose
 * arithmetic -- including expression order -- it follows, so that on the same
 * machine (same glibc libm, no FMA contraction: build with -ffp-contract=off)
 * it reproduces the unmodified reference BIT FOR BIT.  That property is what
 * tests/test_oracle_vs_reference.py and the tests/golden fixtures pin.
 *
 * n_ranks > 1 emulates the reference's replicated MPI run inside one process:
 * ranks are stepped one after another and every MPI_Allreduce becomes a
 * rank-ordered sum (rank 0 first), the same definition oracle/refshim/mpi.h
 * gives the real reference build.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may
 * use this file.
 */
#include "imc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* reference src/constants.h:16-24 */
static const double K_PI = 3.1415926535897932384626433832795;
static const double K_C = 299.792458;
static const double K_A = 0.01372;
static const double K_CUTOFF = 0.01;

/* ------------------------------------------------------------------------ */
/* RNG: reference src/random123/threefry.h:86-93 (rotations), :170-171      */
/* (parity), :196-282 (rounds + key injection every 4 rounds)               */
/* ------------------------------------------------------------------------ */
static inline uint64_t rotl64(uint64_t x, unsigned n) { return (x << (n & 63)) | (x >> ((64 - n) & 63)); }

void orc_threefry2x64_20(const uint64_t ctr[2], const uint64_t key[2], uint64_t out[2]) {
  static const unsigned R[8] = {16, 42, 12, 31, 16, 32, 24, 21};
  uint64_t ks[3];
  ks[0] = key[0];
  ks[1] = key[1];
  ks[2] = 0x1BD11BDAA9FC1A22ULL ^ key[0] ^ key[1];
  uint64_t x0 = ctr[0] + ks[0];
  uint64_t x1 = ctr[1] + ks[1];
  for (unsigned r = 0; r < 20; ++r) {
    x0 += x1;
    x1 = rotl64(x1, R[r & 7]);
    x1 ^= x0;
    if ((r & 3) == 3) {
      unsigned j = (r >> 2) + 1; /* injection number 1..5 */
      x0 += ks[j % 3];
      x1 += ks[(j + 1) % 3];
      x1 += j;
    }
  }
  out[0] = x0;
  out[1] = x1;
}

/* reference src/RNG.h:318-330 */
void orc_rng_init(uint64_t st[4], uint32_t seed, uint64_t stream) {
  st[0] = 0;
  st[1] = ((uint64_t)seed) << 32;
  st[2] = stream;
  st[3] = 0;
}

/* reference src/RNG.h:262-285 (_ran) and :202-238 (u01fixedpt<double,uint64_t>),
 * counter increment src/random123/array.h:133-165 */
double orc_rng_next(uint64_t st[4]) {
  uint64_t out[2];
  orc_threefry2x64_20(&st[0], &st[2], out);
  st[0] += 1;
  if (st[0] == 0) st[1] += 1;
  return (double)(1ULL | (out[0] >> 11)) * (1.0 / 9007199254740992.0);
}

/* reference src/sampling_functions.h:57-70 */
void orc_uniform_angle(uint64_t st[4], double angle[3]) {
  double mu = orc_rng_next(st) * 2.0 - 1.0;
  double phi = orc_rng_next(st) * 2.0 * K_PI;
  double sin_theta = sqrt(1.0 - mu * mu);
  angle[0] = sin_theta * cos(phi);
  angle[1] = sin_theta * sin(phi);
  angle[2] = mu;
}

/* reference src/sampling_functions.h:94-121 (signs exactly as the reference has them) */
static void source_angle_on_face(uint64_t st[4], int face, double angle[3]) {
  double theta = acos(sqrt(orc_rng_next(st)));
  double phi = orc_rng_next(st) * 2.0 * K_PI;
  double sign = (face % 2) ? -1.0 : 1.0;
  if (face == 0 || face == 1) {
    angle[0] = cos(theta) * sign;
    angle[1] = sin(theta) * sin(phi);
    angle[2] = sin(theta) * cos(phi);
  } else if (face == 2 || face == 3) {
    angle[0] = sin(theta) * sin(phi);
    angle[1] = cos(theta);
    angle[2] = sin(theta) * cos(phi);
  } else {
    angle[0] = sin(theta) * cos(phi);
    angle[1] = sin(theta) * sin(phi);
    angle[2] = cos(theta);
  }
}

/* reference src/cell.h:23-25 (sgn) and :116-132 */
double orc_distance_to_boundary(const double nodes[6], const double pos[3], const double angle[3], uint32_t *surface) {
  double min_dist = 1.0e16;
  for (uint32_t i = 0; i < 3; ++i) {
    uint32_t index = 2 * i + (0.0 < angle[i] ? 1u : 0u);
    double dist = (nodes[index] - pos[i]) / angle[i];
    if (dist < min_dist) {
      min_dist = dist;
      *surface = index;
    }
  }
  return min_dist;
}

/* ------------------------------------------------------------------------ */
/* data                                                                     */
/* ------------------------------------------------------------------------ */
typedef struct { /* reference src/photon.h:171-182 (+ event counters, oracle only) */
  uint32_t cell, group, source_type;
  uint8_t descriptor;
  double pos[3], angle[3], E, E0, life_dx;
  uint64_t rng[4];
  uint32_t cnt[4]; /* events, scatters, element crossings, reflections */
} Photon;

typedef struct { /* reference src/cell.h:318-337 */
  uint32_t region_index;
  uint32_t e_next[6];
  int32_t bc[6];
  double nodes[6];
  double cV, op_a, op_s, f, rho, T_e, T_r, T_s;
} Cell;

typedef struct {
  uint32_t id;
  double rho, cV, opacA, opacB, opacC, opacS, T_e, T_r;
} Region;

typedef struct {
  double *E_emission, *E_census, *E_source; /* reference src/mesh.h m_*_E */
  double total_photon_E;
  Photon *census;
  uint64_t n_census, cap_census;
  Photon *photons; /* all_photons of the current cycle */
  uint64_t n_photons, cap_photons, n_new;
  double *rank_abs_E, *rank_track_E;
  /* IMC_State diagnostics, reference src/imc_state.h:448-480 */
  double pre_census_E, post_census_E, pre_mat_E, post_mat_E, emission_E, exit_E, absorbed_E, source_E;
  double rank_total_photon_E;
  /* kept SoA copies for orc_get */
  uint32_t *k_cell[2], *k_group[2], *k_stype, *k_cnt;
  uint64_t *k_ctr[2], *k_stream;
  double *k_pos[2], *k_ang[2], *k_E[2], *k_E0, *k_life[2];
  uint8_t *k_desc;
  uint64_t k_n;
} Rank;

typedef struct {
  char name[48];
  int rank;
  const void *data;
  uint64_t count;
  int dtype;
} Entry;

struct orc_sim {
  uint32_t ngx, ngy, ngz, n_cells, n_groups;
  int n_ranks;
  uint64_t n_user;
  uint32_t seed;
  int n_regions;
  Region *regions;
  Cell *cells;
  double *abs_groups, *sct_groups; /* [n_cells][G], reference src/cell.h:260-275 */
  double *T_r_diag;
  double *abs_E, *track_E, *T_e_pre;
  double *f_arr, *opa_arr, *ops_arr, *Te_arr;
  double *mesh_nodes;
  uint32_t *mesh_region, *mesh_enext, *mesh_bc;
  Rank *ranks;
  /* IMC_State time stepping, reference src/imc_state.h:42-45 */
  double dt, time, time_stop, dt_mult, dt_max;
  uint32_t step;
  double replicated_factor;
  double global_source_energy, next_dt_used, cycle_dt, cycle_time;
  double transport_seconds;
  Entry *entries;
  int n_entries, cap_entries;
  /* scalars exposed through orc_get (per cycle) */
  double sc_dt, sc_time, sc_next_dt, sc_gse;
  uint64_t *sc_u64; /* per rank: n_new, n_photons, n_census */
};

static void *xcalloc(size_t n, size_t sz) {
  void *p = calloc(n ? n : 1, sz);
  if (!p) {
    fprintf(stderr, "imc_oracle: out of memory\n");
    abort();
  }
  return p;
}
static void *xrealloc(void *p, size_t bytes) {
  void *q = realloc(p, bytes ? bytes : 1);
  if (!q) {
    fprintf(stderr, "imc_oracle: out of memory\n");
    abort();
  }
  return q;
}

static void reg(orc_sim *s, int rank, const char *name, const void *data, uint64_t count, int dtype) {
  if (s->n_entries == s->cap_entries) {
    s->cap_entries = s->cap_entries ? 2 * s->cap_entries : 256;
    s->entries = (Entry *)xrealloc(s->entries, sizeof(Entry) * (size_t)s->cap_entries);
  }
  Entry *e = &s->entries[s->n_entries++];
  snprintf(e->name, sizeof(e->name), "%s", name);
  e->rank = rank;
  e->data = data;
  e->count = count;
  e->dtype = dtype;
}

int orc_get(const orc_sim *s, int rank, const char *name, const void **data, uint64_t *count, int *dtype) {
  for (int i = s->n_entries - 1; i >= 0; --i) {
    const Entry *e = &s->entries[i];
    if ((e->rank == rank || e->rank < 0) && strcmp(e->name, name) == 0) {
      *data = e->data;
      *count = e->count;
      *dtype = e->dtype;
      return 0;
    }
  }
  return 1;
}

/* ------------------------------------------------------------------------ */
/* mesh generation: reference src/proto_mesh.h:106-218, dx from             */
/* src/input.h:788-798, region lookup src/input.h:871-875                   */
/* ------------------------------------------------------------------------ */
static int region_index_of(const orc_sim *s, uint32_t id) {
  for (int i = 0; i < s->n_regions; ++i)
    if (s->regions[i].id == id) return i;
  fprintf(stderr, "imc_oracle: region id %u not defined\n", id);
  abort();
}

static void build_mesh(orc_sim *s, const orc_problem *p) {
  uint32_t count = 0, g_i, g_j = 0, g_k = 0;
  const uint32_t ngx = s->ngx, ngy = s->ngy, ngz = s->ngz;
  for (int izd = 0; izd < p->n_zdiv; ++izd) {
    double dz = (p->z_end[izd] - p->z_start[izd]) / p->z_cells[izd];
    uint32_t nz = p->z_cells[izd];
    double z_start = p->z_start[izd];
    for (uint32_t k = 0; k < nz; ++k) {
      g_j = 0;
      for (int iyd = 0; iyd < p->n_ydiv; ++iyd) {
        double dy = (p->y_end[iyd] - p->y_start[iyd]) / p->y_cells[iyd];
        uint32_t ny = p->y_cells[iyd];
        double y_start = p->y_start[iyd];
        for (uint32_t j = 0; j < ny; ++j) {
          g_i = 0;
          for (int ixd = 0; ixd < p->n_xdiv; ++ixd) {
            double dx = (p->x_end[ixd] - p->x_start[ixd]) / p->x_cells[ixd];
            uint32_t nx = p->x_cells[ixd];
            double x_start = p->x_start[ixd];
            for (uint32_t i = 0; i < nx; ++i) {
              Cell *e = &s->cells[count];
              uint32_t rid = p->div_region[((size_t)izd * p->n_ydiv + iyd) * p->n_xdiv + ixd];
              e->region_index = (uint32_t)region_index_of(s, rid);
              double x_end, y_end, z_end;
              if (i == nx - 1 && ixd != p->n_xdiv - 1) x_end = p->x_start[ixd + 1];
              else x_end = x_start + (i + 1) * dx;
              if (j == ny - 1 && iyd != p->n_ydiv - 1) y_end = p->y_start[iyd + 1];
              else y_end = y_start + (j + 1) * dy;
              if (k == nz - 1 && izd != p->n_zdiv - 1) z_end = p->z_start[izd + 1];
              else z_end = z_start + (k + 1) * dz;
              e->nodes[0] = x_start + i * dx;
              e->nodes[1] = x_end;
              e->nodes[2] = y_start + j * dy;
              e->nodes[3] = y_end;
              e->nodes[4] = z_start + k * dz;
              e->nodes[5] = z_end;
              /* neighbours / boundary conditions */
              if (g_i < ngx - 1) { e->e_next[1] = count + 1; e->bc[1] = ORC_ELEMENT; }
              else { e->e_next[1] = count; e->bc[1] = p->bc[1]; }
              if (g_i > 0) { e->e_next[0] = count - 1; e->bc[0] = ORC_ELEMENT; }
              else { e->e_next[0] = count; e->bc[0] = p->bc[0]; }
              if (g_j < ngy - 1) { e->e_next[3] = count + ngx; e->bc[3] = ORC_ELEMENT; }
              else { e->e_next[3] = count; e->bc[3] = p->bc[3]; }
              if (g_j > 0) { e->e_next[2] = count - ngx; e->bc[2] = ORC_ELEMENT; }
              else { e->e_next[2] = count; e->bc[2] = p->bc[2]; }
              if (g_k < ngz - 1) { e->e_next[5] = count + ngx * ngy; e->bc[5] = ORC_ELEMENT; }
              else { e->e_next[5] = count; e->bc[5] = p->bc[5]; }
              if (g_k > 0) { e->e_next[4] = count - ngx * ngy; e->bc[4] = ORC_ELEMENT; }
              else { e->e_next[4] = count; e->bc[4] = p->bc[4]; }
              ++count;
              ++g_i;
            }
          }
          ++g_j;
        }
      }
      ++g_k;
    }
  }
}

/* reference src/cell.h:69-76 */
static int source_face(const Cell *c) {
  for (int i = 0; i < 6; ++i)
    if (c->bc[i] == ORC_SOURCE) return i;
  return -1;
}
/* reference src/cell.h:83-100 */
static double face_area(const Cell *c, int face) {
  const double *n = c->nodes;
  if (face == 0 || face == 1) return (n[3] - n[2]) * (n[5] - n[4]);
  if (face == 2 || face == 3) return (n[1] - n[0]) * (n[5] - n[4]);
  if (face == 4 || face == 5) return (n[1] - n[0]) * (n[3] - n[2]);
  return -1.0;
}
/* reference src/cell.h:188-191 */
static double cell_volume(const Cell *c) {
  const double *n = c->nodes;
  return (n[1] - n[0]) * (n[3] - n[2]) * (n[5] - n[4]);
}

orc_sim *orc_create(const orc_problem *p) {
  orc_sim *s = (orc_sim *)xcalloc(1, sizeof(orc_sim));
  s->ngx = s->ngy = s->ngz = 0;
  for (int i = 0; i < p->n_xdiv; ++i) s->ngx += p->x_cells[i];
  for (int i = 0; i < p->n_ydiv; ++i) s->ngy += p->y_cells[i];
  for (int i = 0; i < p->n_zdiv; ++i) s->ngz += p->z_cells[i];
  s->n_cells = s->ngx * s->ngy * s->ngz;
  s->n_groups = p->n_groups;
  s->n_ranks = p->n_ranks < 1 ? 1 : p->n_ranks;
  s->n_user = p->n_photons;
  s->seed = p->seed;
  s->n_regions = p->n_regions;
  s->regions = (Region *)xcalloc((size_t)p->n_regions, sizeof(Region));
  for (int i = 0; i < p->n_regions; ++i) {
    const double *q = &p->region_props[8 * i];
    Region *r = &s->regions[i];
    r->id = p->region_id[i];
    r->rho = q[0]; r->cV = q[1]; r->opacA = q[2]; r->opacB = q[3];
    r->opacC = q[4]; r->opacS = q[5]; r->T_e = q[6]; r->T_r = q[7];
  }
  size_t nc = s->n_cells, G = s->n_groups;
  s->cells = (Cell *)xcalloc(nc, sizeof(Cell));
  s->abs_groups = (double *)xcalloc(nc * G, 8);
  s->sct_groups = (double *)xcalloc(nc * G, 8);
  s->T_r_diag = (double *)xcalloc(nc, 8);
  s->abs_E = (double *)xcalloc(nc, 8);
  s->track_E = (double *)xcalloc(nc, 8);
  s->T_e_pre = (double *)xcalloc(nc, 8);
  s->f_arr = (double *)xcalloc(nc, 8);
  s->opa_arr = (double *)xcalloc(nc, 8);
  s->ops_arr = (double *)xcalloc(nc, 8);
  s->Te_arr = (double *)xcalloc(nc, 8);
  build_mesh(s, p);
  /* reference src/mesh.h:426-440 initialize_physical_properties */
  for (size_t i = 0; i < nc; ++i) {
    Cell *c = &s->cells[i];
    const Region *r = &s->regions[c->region_index];
    c->cV = r->cV;
    c->T_e = r->T_e;
    c->T_r = r->T_r;
    c->rho = r->rho;
    c->T_s = 0.0;
    if (source_face(c) != -1) c->T_s = p->T_source;
  }
  s->mesh_nodes = (double *)xcalloc(nc * 6, 8);
  s->mesh_region = (uint32_t *)xcalloc(nc, 4);
  s->mesh_enext = (uint32_t *)xcalloc(nc * 6, 4);
  s->mesh_bc = (uint32_t *)xcalloc(nc * 6, 4);
  for (size_t i = 0; i < nc; ++i) {
    for (int k = 0; k < 6; ++k) {
      s->mesh_nodes[6 * i + k] = s->cells[i].nodes[k];
      s->mesh_enext[6 * i + k] = s->cells[i].e_next[k];
      s->mesh_bc[6 * i + k] = (uint32_t)s->cells[i].bc[k];
    }
    s->mesh_region[i] = s->regions[s->cells[i].region_index].id;
  }
  s->ranks = (Rank *)xcalloc((size_t)s->n_ranks, sizeof(Rank));
  for (int r = 0; r < s->n_ranks; ++r) {
    Rank *k = &s->ranks[r];
    k->E_emission = (double *)xcalloc(nc, 8);
    k->E_census = (double *)xcalloc(nc, 8);
    k->E_source = (double *)xcalloc(nc, 8);
    k->rank_abs_E = (double *)xcalloc(nc, 8);
    k->rank_track_E = (double *)xcalloc(nc, 8);
  }
  s->sc_u64 = (uint64_t *)xcalloc((size_t)s->n_ranks * 3, 8);
  /* reference src/imc_state.h:42-45 and src/mesh.h:87 */
  s->dt = p->dt_start;
  s->time = p->t_start;
  s->time_stop = p->t_stop;
  s->dt_mult = p->t_mult;
  s->dt_max = p->dt_max;
  s->step = 1;
  s->replicated_factor = 1.0 / (double)s->n_ranks;
  return s;
}

static void free_kept(Rank *k) {
  for (int i = 0; i < 2; ++i) {
    free(k->k_cell[i]); free(k->k_group[i]); free(k->k_ctr[i]); free(k->k_pos[i]);
    free(k->k_ang[i]); free(k->k_E[i]); free(k->k_life[i]);
    k->k_cell[i] = k->k_group[i] = NULL; k->k_ctr[i] = NULL;
    k->k_pos[i] = k->k_ang[i] = k->k_E[i] = k->k_life[i] = NULL;
  }
  free(k->k_stype); free(k->k_cnt); free(k->k_stream); free(k->k_E0); free(k->k_desc);
  k->k_stype = k->k_cnt = NULL; k->k_stream = NULL; k->k_E0 = NULL; k->k_desc = NULL;
}

void orc_destroy(orc_sim *s) {
  if (!s) return;
  for (int r = 0; r < s->n_ranks; ++r) {
    Rank *k = &s->ranks[r];
    free(k->E_emission); free(k->E_census); free(k->E_source);
    free(k->rank_abs_E); free(k->rank_track_E); free(k->census); free(k->photons);
    free_kept(k);
  }
  free(s->ranks); free(s->regions); free(s->cells); free(s->abs_groups); free(s->sct_groups);
  free(s->T_r_diag); free(s->abs_E); free(s->track_E); free(s->T_e_pre); free(s->f_arr);
  free(s->opa_arr); free(s->ops_arr); free(s->Te_arr); free(s->mesh_nodes); free(s->mesh_region);
  free(s->mesh_enext); free(s->mesh_bc); free(s->entries); free(s->sc_u64);
  free(s);
}

/* reference src/imc_state.h:112-125 */
static double get_next_dt(const orc_sim *s) {
  double next_dt;
  if (s->dt * s->dt_mult < s->dt_max) next_dt = s->dt * s->dt_mult;
  else next_dt = s->dt_max;
  if (s->time + next_dt > s->time_stop) next_dt = s->time_stop - s->time;
  return next_dt;
}
/* reference src/imc_state.h:127-134 */
int orc_finished(const orc_sim *s) { return fabs(s->time - s->time_stop) < 1.0e-8; }

/* ------------------------------------------------------------------------ */
/* reference src/mesh.h:237-323 calculate_photon_energy                     */
/* ------------------------------------------------------------------------ */
static void calculate_photon_energy(orc_sim *s) {
  const uint32_t n_user32 = (uint32_t)s->n_user; /* the reference narrows to uint32_t (mesh.h:237) */
  const double dt = s->dt;
  const size_t nc = s->n_cells, G = s->n_groups;
  /* opacities and Fleck factor are rank independent */
  for (size_t i = 0; i < nc; ++i) {
    Cell *e = &s->cells[i];
    const Region *region = &s->regions[e->region_index];
    double T = e->T_e;
    double op_a = region->opacA + region->opacB * pow(T, region->opacC); /* region.h:44-46 */
    double op_s = region->opacS;
    double f = 1.0 / (1.0 + dt * op_a * K_C * (4.0 * K_A * pow(T, 3) / (e->cV * e->rho)));
    e->op_a = op_a;
    e->op_s = op_s;
    e->f = f;
    for (size_t g = 0; g < G; ++g) {
      s->abs_groups[i * G + g] = op_a;
      s->sct_groups[i * G + g] = op_s;
    }
  }
  double rank_sum[64];
  for (int r = 0; r < s->n_ranks; ++r) {
    Rank *k = &s->ranks[r];
    double tot_census_E = 0.0, tot_emission_E = 0.0, tot_source_E = 0.0, pre_mat_E = 0.0;
    k->total_photon_E = 0.0;
    for (size_t i = 0; i < nc; ++i) {
      Cell *e = &s->cells[i];
      double vol = cell_volume(e);
      double T = e->T_e, Tr = e->T_r, Ts = e->T_s;
      k->E_emission[i] = s->replicated_factor * dt * vol * e->f * e->op_a * K_A * K_C * pow(T, 4);
      if (s->step > 1) k->E_census[i] = 0.0;
      else k->E_census[i] = s->replicated_factor * vol * K_A * pow(Tr, 4);
      k->E_source[i] = s->replicated_factor * 0.25 * K_A * K_C * face_area(e, source_face(e)) * pow(Ts, 4) * dt;
      pre_mat_E += T * e->cV * vol * e->rho;
      tot_emission_E += k->E_emission[i];
      tot_census_E += k->E_census[i];
      tot_source_E += k->E_source[i];
      k->total_photon_E += k->E_source[i] + k->E_census[i] + k->E_emission[i];
    }
    k->pre_mat_E = pre_mat_E;
    k->emission_E = tot_emission_E;
    k->source_E = tot_source_E;
    if (s->step == 1) k->pre_census_E = tot_census_E;
    rank_sum[r] = tot_emission_E + tot_census_E + tot_source_E;
  }
  /* replicated redistribution, mesh.h:291-315 (always taken: replicated mode) */
  double global_source_E = rank_sum[0];
  for (int r = 1; r < s->n_ranks; ++r) global_source_E = global_source_E + rank_sum[r];
  for (int r = 0; r < s->n_ranks; ++r) {
    Rank *k = &s->ranks[r];
    double tot_census_E = 0.0, tot_emission_E = 0.0, tot_source_E = 0.0;
    k->total_photon_E = 0.0;
    for (uint32_t i = 0; i < nc; ++i) {
      if (s->step == 1 && k->E_census[i] > 0.0 && (int)(n_user32 * (k->E_census[i] / global_source_E)) == 0)
        k->E_census[i] = ((int)(i % (uint32_t)s->n_ranks) == r) ? k->E_census[i] / s->replicated_factor : 0.0;
      if (k->E_emission[i] > 0.0 && (int)(n_user32 * (k->E_emission[i] / global_source_E)) == 0)
        k->E_emission[i] = ((int)(i % (uint32_t)s->n_ranks) == r) ? k->E_emission[i] / s->replicated_factor : 0.0;
      if (k->E_source[i] > 0.0 && (int)(n_user32 * (k->E_source[i] / global_source_E)) == 0)
        k->E_source[i] = ((int)(i % (uint32_t)s->n_ranks) == r) ? k->E_source[i] / s->replicated_factor : 0.0;
      tot_emission_E += k->E_emission[i];
      tot_census_E += k->E_census[i];
      tot_source_E += k->E_source[i];
      k->total_photon_E += k->E_source[i] + k->E_census[i] + k->E_emission[i];
    }
    k->emission_E = tot_emission_E;
    k->source_E = tot_source_E;
    if (s->step == 1) k->pre_census_E = tot_census_E;
    k->rank_total_photon_E = k->total_photon_E;
  }
}

/* ------------------------------------------------------------------------ */
/* sourcing: reference src/source.h                                         */
/* ------------------------------------------------------------------------ */
static void uniform_position_in_cell(const Cell *c, uint64_t st[4], double pos[3]) { /* sampling_functions.h:22-29 */
  const double *n = c->nodes;
  pos[0] = n[0] + orc_rng_next(st) * (n[1] - n[0]);
  pos[1] = n[2] + orc_rng_next(st) * (n[3] - n[2]);
  pos[2] = n[4] + orc_rng_next(st) * (n[5] - n[4]);
}
static void uniform_position_on_face(const Cell *c, uint64_t st[4], int face, double pos[3]) { /* :32-52 */
  const double *n = c->nodes;
  if (face == 0 || face == 1) {
    pos[0] = (face == 0) ? n[0] : n[1];
    pos[1] = n[2] + orc_rng_next(st) * (n[3] - n[2]);
    pos[2] = n[4] + orc_rng_next(st) * (n[5] - n[4]);
  } else if (face == 2 || face == 3) {
    pos[0] = n[0] + orc_rng_next(st) * (n[1] - n[0]);
    pos[1] = (face == 2) ? n[2] : n[3];
    pos[2] = n[4] + orc_rng_next(st) * (n[5] - n[4]);
  } else {
    pos[0] = n[0] + orc_rng_next(st) * (n[1] - n[0]);
    pos[1] = n[2] + orc_rng_next(st) * (n[3] - n[2]);
    pos[2] = (face == 4) ? n[4] : n[5];
  }
}

static Photon *push_photon(Photon **arr, uint64_t *n, uint64_t *cap) {
  if (*n == *cap) {
    *cap = *cap ? 2 * *cap : 1024;
    *arr = (Photon *)xrealloc(*arr, sizeof(Photon) * (size_t)*cap);
  }
  Photon *p = &(*arr)[(*n)++];
  memset(p, 0, sizeof(Photon));
  return p;
}

/* source.h:84-97 */
static void emission_photon(Photon *p, const orc_sim *s, uint32_t cell_id, double E, double dt, uint64_t stream) {
  const Cell *c = &s->cells[cell_id];
  orc_rng_init(p->rng, s->seed, stream);
  p->source_type = 2;
  uniform_position_in_cell(c, p->rng, p->pos);
  orc_uniform_angle(p->rng, p->angle);
  p->E0 = E;
  p->E = E;
  p->life_dx = orc_rng_next(p->rng) * K_C * dt;
  p->cell = cell_id;
  p->group = (uint32_t)floor(orc_rng_next(p->rng) * (double)s->n_groups);
  p->descriptor = ORC_PASS;
}
/* source.h:100-114 */
static void boundary_source_photon(Photon *p, const orc_sim *s, uint32_t cell_id, double E, double dt,
                                   uint64_t stream, int face) {
  const Cell *c = &s->cells[cell_id];
  orc_rng_init(p->rng, s->seed, stream);
  p->source_type = 1;
  uniform_position_on_face(c, p->rng, face, p->pos);
  source_angle_on_face(p->rng, face, p->angle);
  p->E0 = E;
  p->E = E;
  p->life_dx = orc_rng_next(p->rng) * K_C * dt;
  p->cell = cell_id;
  p->group = (uint32_t)floor(orc_rng_next(p->rng) * (double)s->n_groups);
  p->descriptor = ORC_PASS;
}
/* source.h:117-131 */
static void initial_census_photon(Photon *p, const orc_sim *s, uint32_t cell_id, double E, double dt, uint64_t stream) {
  const Cell *c = &s->cells[cell_id];
  orc_rng_init(p->rng, s->seed, stream);
  p->source_type = 0;
  uniform_position_in_cell(c, p->rng, p->pos);
  orc_uniform_angle(p->rng, p->angle);
  p->E0 = E;
  p->E = E;
  p->life_dx = K_C * dt;
  p->cell = cell_id;
  p->group = (uint32_t)floor(orc_rng_next(p->rng) * (double)s->n_groups);
  p->descriptor = ORC_PASS;
}

/* source.h:183-204 */
static void make_initial_census_photons(orc_sim *s, int rank, double total_E) {
  Rank *k = &s->ranks[rank];
  const uint64_t rank_off = s->n_user * (uint64_t)rank;
  uint64_t ith = 0;
  k->n_census = 0;
  for (uint32_t i = 0; i < s->n_cells; ++i) {
    if (k->E_census[i] > 0.0) {
      uint32_t n = (uint32_t)(int)(s->n_user * k->E_census[i] / total_E);
      if (n == 0) n = 1;
      const double E = k->E_census[i] / n;
      for (uint32_t q = 0; q < n; ++q) {
        initial_census_photon(push_photon(&k->census, &k->n_census, &k->cap_census), s, i, E, s->dt, rank_off + ith);
        ++ith;
      }
    }
  }
}

/* source.h:294-366 */
static void make_photons(orc_sim *s, int rank, double total_E) {
  Rank *k = &s->ranks[rank];
  const uint64_t cycle_off = 10000000000000ULL * (uint64_t)s->step;
  const uint64_t rank_off = s->n_user * (uint64_t)rank;
  uint64_t ith = 0;
  k->n_photons = 0;
  for (uint32_t i = 0; i < s->n_cells; ++i) {
    if (k->E_emission[i] > 0.0) {
      uint32_t n = (uint32_t)(int)(s->n_user * k->E_emission[i] / total_E);
      if (n == 0) n = 1;
      const double E = k->E_emission[i] / n;
      for (uint32_t q = 0; q < n; ++q) {
        emission_photon(push_photon(&k->photons, &k->n_photons, &k->cap_photons), s, i, E, s->dt,
                        cycle_off + rank_off + ith);
        ++ith;
      }
    }
    if (k->E_source[i] > 0.0) {
      uint32_t n = (uint32_t)(int)(s->n_user * k->E_source[i] / total_E);
      if (n == 0) n = 1;
      const double E = k->E_source[i] / n;
      const int face = source_face(&s->cells[i]);
      for (uint32_t q = 0; q < n; ++q) {
        boundary_source_photon(push_photon(&k->photons, &k->n_photons, &k->cap_photons), s, i, E, s->dt,
                               cycle_off + rank_off + ith, face);
        ++ith;
      }
    }
  }
  k->n_new = k->n_photons;
}

/* ------------------------------------------------------------------------ */
/* transport: reference src/history_based_transport.h:32-141                */
/* ------------------------------------------------------------------------ */
/* sampling_functions.h:126-138; the read of abs_groups[G] on round-off is guarded */
static int sample_emission_group(const orc_sim *s, uint64_t st[4], uint32_t cell) {
  const double *ag = &s->abs_groups[(size_t)cell * s->n_groups];
  double cdf_value = orc_rng_next(st);
  int new_group = -1;
  double norm_factor = 1.0 / (ag[0] * s->n_groups);
  while (cdf_value > 0) {
    new_group++;
    if ((uint32_t)new_group >= s->n_groups) break;
    cdf_value -= ag[new_group] * norm_factor;
  }
  return new_group;
}

static void transport_photon(const orc_sim *s, Photon *ph, double *abs_E, double *track_E) {
  const size_t G = s->n_groups;
  uint32_t surface_cross = 0;
  uint32_t ci = ph->cell;
  const Cell *cell = &s->cells[ci];
  int active = 1;
  double thread_absorbed_E = 0.0, thread_track_E = 0.0;
  while (active) {
    const double sigma_s = s->sct_groups[ci * G + ph->group];
    const double sigma_a = s->abs_groups[ci * G + ph->group];
    const double f = cell->f;
    const double total_sigma_s = (1.0 - f) * sigma_a + sigma_s;
    const double dist_to_scatter = (total_sigma_s > 0.0) ? -log(orc_rng_next(ph->rng)) / total_sigma_s : 1.0e100;
    const double dist_to_boundary = orc_distance_to_boundary(cell->nodes, ph->pos, ph->angle, &surface_cross);
    const double dist_to_census = ph->life_dx;
    const double m1 = dist_to_census < dist_to_boundary ? dist_to_census : dist_to_boundary; /* std::min(b,c) */
    const double dist_to_event = m1 < dist_to_scatter ? m1 : dist_to_scatter;               /* std::min(a,m1) */
    const double absorbed_E = ph->E * (1.0 - exp(-sigma_a * f * dist_to_event));
    thread_absorbed_E += absorbed_E;
    thread_track_E += absorbed_E / (sigma_a * f);
    ph->E = ph->E - absorbed_E;
    ph->pos[0] += ph->angle[0] * dist_to_event;
    ph->pos[1] += ph->angle[1] * dist_to_event;
    ph->pos[2] += ph->angle[2] * dist_to_event;
    ph->life_dx -= dist_to_event;
    ph->cnt[0]++;
    if (ph->E / ph->E0 < K_CUTOFF) {
      thread_absorbed_E += ph->E;
      abs_E[ci] += thread_absorbed_E;
      track_E[ci] += thread_track_E;
      active = 0;
      ph->descriptor = ORC_KILLED;
    } else if (dist_to_event == dist_to_scatter) {
      orc_uniform_angle(ph->rng, ph->angle);
      if (orc_rng_next(ph->rng) > (sigma_s / ((1.0 - f) * sigma_a + sigma_s)))
        ph->group = (uint32_t)sample_emission_group(s, ph->rng, ci);
      ph->descriptor = ORC_SCATTER;
      ph->cnt[1]++;
    } else if (dist_to_event == dist_to_boundary) {
      int boundary_event = cell->bc[surface_cross];
      if (boundary_event == ORC_ELEMENT) {
        abs_E[ci] += thread_absorbed_E;
        track_E[ci] += thread_track_E;
        ph->cell = cell->e_next[surface_cross];
        ci = ph->cell;
        cell = &s->cells[ci];
        ph->descriptor = ORC_BOUND;
        thread_absorbed_E = 0.0;
        thread_track_E = 0.0;
        ph->cnt[2]++;
      } else if (boundary_event == ORC_PROCESSOR) {
        active = 0;
        ph->cell = cell->e_next[surface_cross];
        ph->descriptor = ORC_PASS;
        abs_E[ci] += thread_absorbed_E;
        track_E[ci] += thread_track_E;
      } else if (boundary_event == ORC_VACUUM || boundary_event == ORC_SOURCE) {
        active = 0;
        ph->descriptor = ORC_EXIT;
        abs_E[ci] += thread_absorbed_E;
        track_E[ci] += thread_track_E;
      } else {
        int ax = (int)(surface_cross / 2);
        ph->angle[ax] = -ph->angle[ax];
        ph->descriptor = ORC_BOUND;
        ph->cnt[3]++;
      }
    } else if (dist_to_event == dist_to_census) {
      active = 0;
      ph->descriptor = ORC_CENSUS;
      abs_E[ci] += thread_absorbed_E;
      track_E[ci] += thread_track_E;
    }
  }
}

int orc_transport_list(orc_sim *s, int rank, uint64_t n, uint32_t *cell, uint32_t *group, double *pos, double *angle,
                       double *E, const double *E0, double *life_dx, uint64_t *ctr, const uint64_t *stream,
                       uint8_t *descriptor, double *abs_E, double *track_E, uint32_t *counters) {
  (void)rank;
  for (uint64_t i = 0; i < n; ++i) {
    Photon p;
    memset(&p, 0, sizeof(p));
    p.cell = cell[i];
    p.group = group[i];
    for (int k = 0; k < 3; ++k) {
      p.pos[k] = pos[3 * i + k];
      p.angle[k] = angle[3 * i + k];
    }
    p.E = E[i];
    p.E0 = E0[i];
    p.life_dx = life_dx[i];
    p.rng[0] = ctr[i];
    p.rng[1] = ((uint64_t)s->seed) << 32;
    p.rng[2] = stream[i];
    p.rng[3] = 0;
    transport_photon(s, &p, abs_E, track_E);
    cell[i] = p.cell;
    group[i] = p.group;
    for (int k = 0; k < 3; ++k) {
      pos[3 * i + k] = p.pos[k];
      angle[3 * i + k] = p.angle[k];
    }
    E[i] = p.E;
    life_dx[i] = p.life_dx;
    ctr[i] = p.rng[0];
    descriptor[i] = p.descriptor;
    if (counters) memcpy(&counters[4 * i], p.cnt, 16);
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
static void keep_side(Rank *k, int side, const Photon *p, uint64_t n) {
  k->k_cell[side] = (uint32_t *)xcalloc(n, 4);
  k->k_group[side] = (uint32_t *)xcalloc(n, 4);
  k->k_ctr[side] = (uint64_t *)xcalloc(n, 8);
  k->k_pos[side] = (double *)xcalloc(3 * n, 8);
  k->k_ang[side] = (double *)xcalloc(3 * n, 8);
  k->k_E[side] = (double *)xcalloc(n, 8);
  k->k_life[side] = (double *)xcalloc(n, 8);
  if (side == 0) {
    k->k_stype = (uint32_t *)xcalloc(n, 4);
    k->k_stream = (uint64_t *)xcalloc(n, 8);
    k->k_E0 = (double *)xcalloc(n, 8);
  } else {
    k->k_desc = (uint8_t *)xcalloc(n, 1);
    k->k_cnt = (uint32_t *)xcalloc(4 * n, 4);
  }
  for (uint64_t i = 0; i < n; ++i) {
    k->k_cell[side][i] = p[i].cell;
    k->k_group[side][i] = p[i].group;
    k->k_ctr[side][i] = p[i].rng[0];
    for (int q = 0; q < 3; ++q) {
      k->k_pos[side][3 * i + q] = p[i].pos[q];
      k->k_ang[side][3 * i + q] = p[i].angle[q];
    }
    k->k_E[side][i] = p[i].E;
    k->k_life[side][i] = p[i].life_dx;
    if (side == 0) {
      k->k_stype[i] = p[i].source_type;
      k->k_stream[i] = p[i].rng[2];
      k->k_E0[i] = p[i].E0;
    } else {
      k->k_desc[i] = p[i].descriptor;
      memcpy(&k->k_cnt[4 * i], p[i].cnt, 16);
    }
  }
  k->k_n = n;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* one cycle == one iteration of reference src/replicated_driver.h:47-121 */
int orc_cycle(orc_sim *s, int keep_photons) {
  const size_t nc = s->n_cells;
  s->n_entries = 0;
  s->sc_dt = s->dt;
  s->sc_time = s->time;
  for (size_t i = 0; i < nc; ++i) s->T_e_pre[i] = s->cells[i].T_e;
  calculate_photon_energy(s);
  for (size_t i = 0; i < nc; ++i) {
    s->f_arr[i] = s->cells[i].f;
    s->opa_arr[i] = s->cells[i].op_a;
    s->ops_arr[i] = s->cells[i].op_s;
  }
  /* replicated_driver.h:56-59 */
  double gse = s->ranks[0].total_photon_E;
  for (int r = 1; r < s->n_ranks; ++r) gse = gse + s->ranks[r].total_photon_E;
  s->global_source_energy = gse;
  s->sc_gse = gse;
  s->sc_next_dt = get_next_dt(s); /* replicated_transport.h:52 */
  s->transport_seconds = 0.0;

  for (int r = 0; r < s->n_ranks; ++r) {
    Rank *k = &s->ranks[r];
    free_kept(k);
    if (s->step == 1) make_initial_census_photons(s, r, gse);
    /* census_functions.h:40-46 */
    double pre = 0.0;
    for (uint64_t i = 0; i < k->n_census; ++i) pre += k->census[i].E;
    k->pre_census_E = pre;
    make_photons(s, r, gse);
    /* join_photon_arrays, census_functions.h:21-29 */
    for (uint64_t i = 0; i < k->n_census; ++i) *push_photon(&k->photons, &k->n_photons, &k->cap_photons) = k->census[i];
    if (keep_photons) keep_side(k, 0, k->photons, k->n_photons);

    /* replicated_transport.h:71-92: serial photon-order tallies */
    memset(k->rank_abs_E, 0, nc * 8);
    memset(k->rank_track_E, 0, nc * 8);
    double t0 = now_s();
    for (uint64_t i = 0; i < k->n_photons; ++i) transport_photon(s, &k->photons[i], k->rank_abs_E, k->rank_track_E);
    /* post_process_functions.h:33-59 */
    double census_E = 0.0, exit_E = 0.0;
    k->n_census = 0;
    for (uint64_t i = 0; i < k->n_photons; ++i) {
      Photon *p = &k->photons[i];
      switch (p->descriptor) {
      case ORC_EXIT: exit_E += p->E; break;
      case ORC_CENSUS:
        p->life_dx = K_C * s->sc_next_dt;
        *push_photon(&k->census, &k->n_census, &k->cap_census) = *p;
        census_E += p->E;
        break;
      default: break;
      }
    }
    s->transport_seconds += now_s() - t0;
    k->exit_E = exit_E;
    k->post_census_E = census_E;
    if (keep_photons) keep_side(k, 1, k->photons, k->n_photons);
    s->sc_u64[3 * r + 0] = k->n_new;
    s->sc_u64[3 * r + 1] = k->n_photons;
    s->sc_u64[3 * r + 2] = k->n_census;
    /* census photons carry no event counters into the next cycle */
    for (uint64_t i = 0; i < k->n_census; ++i) memset(k->census[i].cnt, 0, 16);
  }

  /* replicated_driver.h:91-94, rank-ordered */
  for (size_t i = 0; i < nc; ++i) {
    double a = s->ranks[0].rank_abs_E[i], t = s->ranks[0].rank_track_E[i];
    for (int r = 1; r < s->n_ranks; ++r) {
      a = a + s->ranks[r].rank_abs_E[i];
      t = t + s->ranks[r].rank_track_E[i];
    }
    s->abs_E[i] = a;
    s->track_E[i] = t;
  }

  /* reference src/mesh.h:327-362 update_temperature */
  double total_abs_E = 0.0, total_post_mat_E = 0.0;
  for (size_t i = 0; i < nc; ++i) {
    Cell *e = &s->cells[i];
    const Region *region = &s->regions[e->region_index];
    double emis = s->ranks[0].E_emission[i]; /* mesh.h:343-345 allreduce of m_emission_E */
    for (int r = 1; r < s->n_ranks; ++r) emis = emis + s->ranks[r].E_emission[i];
    double cV = region->cV, rho = region->rho;
    double vol = cell_volume(e);
    double T = e->T_e;
    double T_new = T + (s->abs_E[i] - emis) / (cV * vol * rho);
    s->T_r_diag[i] = pow(s->track_E[i] / (vol * s->dt * K_A * K_C), 0.25);
    e->T_e = T_new;
    total_abs_E += s->abs_E[i];
    total_post_mat_E += T_new * cV * vol * rho;
    s->Te_arr[i] = T_new;
  }
  for (int r = 0; r < s->n_ranks; ++r) {
    Rank *k = &s->ranks[r];
    k->absorbed_E = total_abs_E;
    k->post_mat_E = total_post_mat_E;
    if (r) { /* replicated_driver.h:100-104 */
      k->absorbed_E = 0.0;
      k->pre_mat_E = 0.0;
      k->post_mat_E = 0.0;
    }
  }

  /* registry */
  reg(s, -1, "dt", &s->sc_dt, 1, 0);
  reg(s, -1, "time", &s->sc_time, 1, 0);
  reg(s, -1, "next_dt", &s->sc_next_dt, 1, 0);
  reg(s, -1, "global_source_energy", &s->sc_gse, 1, 0);
  reg(s, -1, "T_e_pre", s->T_e_pre, nc, 0);
  reg(s, -1, "f", s->f_arr, nc, 0);
  reg(s, -1, "op_a", s->opa_arr, nc, 0);
  reg(s, -1, "op_s", s->ops_arr, nc, 0);
  reg(s, -1, "abs_E", s->abs_E, nc, 0);
  reg(s, -1, "track_E", s->track_E, nc, 0);
  reg(s, -1, "T_e", s->Te_arr, nc, 0);
  reg(s, -1, "T_r", s->T_r_diag, nc, 0);
  reg(s, -1, "mesh/nodes", s->mesh_nodes, nc * 6, 0);
  reg(s, -1, "mesh/region", s->mesh_region, nc, 1);
  reg(s, -1, "mesh/e_next", s->mesh_enext, nc * 6, 1);
  reg(s, -1, "mesh/bc", s->mesh_bc, nc * 6, 1);
  for (int r = 0; r < s->n_ranks; ++r) {
    Rank *k = &s->ranks[r];
    reg(s, r, "rank_total_photon_E", &k->rank_total_photon_E, 1, 0);
    reg(s, r, "E_emission", k->E_emission, nc, 0);
    reg(s, r, "E_census", k->E_census, nc, 0);
    reg(s, r, "E_source", k->E_source, nc, 0);
    reg(s, r, "n_new", &s->sc_u64[3 * r + 0], 1, 2);
    reg(s, r, "n_photons", &s->sc_u64[3 * r + 1], 1, 2);
    reg(s, r, "n_census", &s->sc_u64[3 * r + 2], 1, 2);
    reg(s, r, "pre_census_E", &k->pre_census_E, 1, 0);
    reg(s, r, "exit_E", &k->exit_E, 1, 0);
    reg(s, r, "post_census_E", &k->post_census_E, 1, 0);
    reg(s, r, "rank_abs_E", k->rank_abs_E, nc, 0);
    reg(s, r, "rank_track_E", k->rank_track_E, nc, 0);
    reg(s, r, "emission_E", &k->emission_E, 1, 0);
    reg(s, r, "source_E", &k->source_E, 1, 0);
    reg(s, r, "absorbed_E", &k->absorbed_E, 1, 0);
    reg(s, r, "pre_mat_E", &k->pre_mat_E, 1, 0);
    reg(s, r, "post_mat_E", &k->post_mat_E, 1, 0);
    if (keep_photons) {
      uint64_t n = k->k_n;
      reg(s, r, "pre/cell", k->k_cell[0], n, 1);
      reg(s, r, "pre/group", k->k_group[0], n, 1);
      reg(s, r, "pre/ctr", k->k_ctr[0], n, 2);
      reg(s, r, "pre/pos", k->k_pos[0], 3 * n, 0);
      reg(s, r, "pre/angle", k->k_ang[0], 3 * n, 0);
      reg(s, r, "pre/E", k->k_E[0], n, 0);
      reg(s, r, "pre/life_dx", k->k_life[0], n, 0);
      reg(s, r, "pre/source_type", k->k_stype, n, 1);
      reg(s, r, "pre/stream", k->k_stream, n, 2);
      reg(s, r, "pre/E0", k->k_E0, n, 0);
      reg(s, r, "post/cell", k->k_cell[1], n, 1);
      reg(s, r, "post/group", k->k_group[1], n, 1);
      reg(s, r, "post/ctr", k->k_ctr[1], n, 2);
      reg(s, r, "post/pos", k->k_pos[1], 3 * n, 0);
      reg(s, r, "post/angle", k->k_ang[1], 3 * n, 0);
      reg(s, r, "post/E", k->k_E[1], n, 0);
      reg(s, r, "post/life_dx", k->k_life[1], n, 0);
      reg(s, r, "post/descriptor", k->k_desc, n, 3);
      reg(s, r, "post/counters", k->k_cnt, 4 * n, 1);
    }
  }

  /* reference src/imc_state.h:292-296 next_time_step */
  s->time += s->dt;
  s->dt = get_next_dt(s);
  s->step++;
  return 0;
}

double orc_last_transport_seconds(const orc_sim *s) { return s->transport_seconds; }

/* ------------------------------------------------------------------------ */
/* comb_photons, reference src/census_functions.h:48-93.  The reference keeps three unordered_maps keyed by cell; plain
 * arrays indexed by cell hold the same values (every update is keyed, nothing iterates over the maps). */
uint64_t orc_comb_photons(uint64_t n, const uint32_t *cell, const double *E, double global_census_E,
                          int64_t max_census_photons, uint64_t rng_state[4], uint8_t *keep, double *new_E) {
  uint32_t max_cell = 0;
  for (uint64_t i = 0; i < n; ++i)
    if (cell[i] > max_cell) max_cell = cell[i];
  uint32_t *count = (uint32_t *)xcalloc((size_t)max_cell + 1, 4);
  double *cell_E = (double *)xcalloc((size_t)max_cell + 1, 8);
  double *corrected = (double *)xcalloc((size_t)max_cell + 1, 8);
  const double comb_photon_E = global_census_E / (double)max_census_photons; /* :65 */
  for (uint64_t i = 0; i < n; ++i) {                                         /* :67-70 */
    count[cell[i]]++;
    cell_E[cell[i]] += E[i];
  }
  uint64_t kept = 0;
  for (uint64_t i = 0; i < n; ++i) { /* :72-85 */
    const uint32_t c = cell[i];
    const double p_kill = 1.0 - E[i] / comb_photon_E;
    const double rand_check = orc_rng_next(rng_state);
    if (rand_check > p_kill || count[c] == 1) {
      keep[i] = 1;
      corrected[c] += comb_photon_E;
      ++kept;
    } else {
      keep[i] = 0;
      count[c]--;
    }
  }
  for (uint64_t i = 0; i < n; ++i) { /* :88-93 */
    const uint32_t c = cell[i];
    new_E[i] = keep[i] ? comb_photon_E + (cell_E[c] - corrected[c]) / (double)count[c] : 0.0;
  }
  free(count);
  free(cell_E);
  free(corrected);
  return kept;
}
