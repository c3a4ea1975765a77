// host_capi.cc -- flat C API (include/branson_host.h) over the C++ host layer.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "comm.h"
#include "../../../include/branson_host.h"
#include "gpu_setup.h"
#include "imc_parameters.h"
#include "imc_state.h"
#include "input.h"
#include "mesh.h"
#include "replicated_driver.h"

namespace branson {
void Comm::check(bool ok) {
  if (!ok) throw std::runtime_error("collective failed or not provided by the harness");
}
}  // namespace branson

using namespace branson;

struct bhost_driver {
  std::string err;
  int rank = 0, n_ranks = 1;
  bhost_options opt{};
  std::unique_ptr<Comm> comm;
  std::unique_ptr<Input> input;
  std::unique_ptr<IMC_Parameters> params;
  std::unique_ptr<IMC_State> state;
  std::unique_ptr<Mesh> mesh;
  std::unique_ptr<GPU_Setup> gpu;
  std::unique_ptr<Replicated_Driver> driver;
};

extern "C" {

bhost_driver *bhost_create(const char *xml_path, int rank, int n_ranks, const bhost_options *opt,
                           const bhost_comm *comm, char *err, size_t err_len) {
  auto d = std::make_unique<bhost_driver>();
  try {
    if (!xml_path || !opt) throw std::runtime_error("bhost_create: null argument");
    d->rank = rank;
    d->n_ranks = n_ranks < 1 ? 1 : n_ranks;
    d->opt = *opt;
    d->comm = std::make_unique<Comm>(comm, rank, d->n_ranks);
    d->input = std::make_unique<Input>(xml_path, d->n_ranks, !(opt->print && rank == 0));
    if (opt->photons_override) d->input->set_number_photons(opt->photons_override);
    if (opt->t_stop_override > 0.0) d->input->set_time_finish(opt->t_stop_override);
    if (opt->force_replicated) d->input->set_dd_mode(Constants::REPLICATED);
    d->params = std::make_unique<IMC_Parameters>(*d->input);
    if (opt->algorithm >= 0) d->params->set_transport_algorithm((uint32_t)opt->algorithm);
    d->state = std::make_unique<IMC_State>(*d->input, (uint32_t)rank);
    d->mesh = std::make_unique<Mesh>(*d->input, *d->params, *d->comm);
    if (!opt->no_gpu) {
      d->gpu = std::make_unique<GPU_Setup>(rank, d->n_ranks, d->params->get_use_gpu_transporter_flag(), *d->mesh,
                                           *d->params, opt->n_groups ? opt->n_groups : 1u, opt->device);
      if (opt->validate) d->gpu->check(bgpu_enable_counters(d->gpu->get_ctx(), 1), "bgpu_enable_counters");
      Driver_Options o;
      o.tally_mode = opt->tally_mode;
      o.print = opt->print != 0;
      o.mesh_on_device = opt->mesh_on_device != 0;
      o.comb_max_census = opt->comb_max_census;
      o.sort_census = opt->sort_census != 0;
      d->driver = std::make_unique<Replicated_Driver>(*d->mesh, *d->state, *d->params, *d->comm, *d->gpu, o);
    }
  } catch (const std::exception &e) {
    if (err && err_len) {
      std::strncpy(err, e.what(), err_len - 1);
      err[err_len - 1] = 0;
    }
    return nullptr;
  }
  return d.release();
}

void bhost_destroy(bhost_driver *d) { delete d; }

int bhost_comm_unique_id(char id[BGPU_COMM_ID_BYTES]) { return bgpu_comm_unique_id(id); }

int bhost_comm_init_rank(bhost_driver *d, const char id[BGPU_COMM_ID_BYTES]) {
  if (!d || !d->gpu) return 1;
  if (bgpu_comm_init_rank(d->gpu->get_ctx(), id)) {
    d->err = bgpu_last_error(d->gpu->get_ctx());
    return 1;
  }
  d->comm->attach_native(d->gpu->get_ctx());
  return 0;
}

int bhost_comm_init_local(bhost_driver **drivers, int n) {
  if (!drivers || n < 1) return 1;
  std::vector<bgpu_ctx *> ctxs((size_t)n);
  for (int r = 0; r < n; ++r) {
    if (!drivers[r] || !drivers[r]->gpu) return 1;
    ctxs[(size_t)r] = drivers[r]->gpu->get_ctx();
  }
  if (bgpu_comm_init_local(ctxs.data(), n)) {
    for (int r = 0; r < n; ++r) {
      const char *e = bgpu_last_error(ctxs[(size_t)r]);
      drivers[r]->err = (e && *e) ? e : bgpu_last_error(nullptr);
    }
    return 1;
  }
  for (int r = 0; r < n; ++r) drivers[r]->comm->attach_native(ctxs[(size_t)r]);
  return 0;
}
const char *bhost_last_error(const bhost_driver *d) { return d ? d->err.c_str() : "null driver"; }
int bhost_finished(const bhost_driver *d) { return d && d->state->finished() ? 1 : 0; }

int bhost_calculate_photon_energy(bhost_driver *d, double *gse) {
  if (!d) return 1;
  try {
    d->mesh->calculate_photon_energy(*d->state, (uint32_t)d->params->get_n_user_photons());
    double g = d->mesh->get_total_photon_E();
    d->comm->sum(&g, 1);
    if (gse) *gse = g;
  } catch (const std::exception &e) {
    d->err = e.what();
    return 1;
  }
  return 0;
}

int bhost_cycle(bhost_driver *d, bhost_cycle_report *out) {
  if (!d) return 1;
  if (!d->driver) {
    d->err = "bhost_cycle: created with no_gpu (host logic only); there is no CPU transport";
    return 1;
  }
  try {
    const Cycle_Report r = d->driver->cycle();
    if (out) {
      bhost_cycle_report o{};
      o.step = r.step; o.dt = r.dt; o.time = r.time; o.next_dt = r.next_dt;
      o.global_source_energy = r.global_source_energy;
      o.gpu = r.gpu;
      o.t_calc_energy = r.t_calc_energy; o.t_cell_upload = r.t_cell_upload; o.t_source = r.t_source;
      o.t_transport = r.t_transport; o.t_allreduce = r.t_allreduce; o.t_tally_download = r.t_tally_download;
      o.t_update_T = r.t_update_T; o.t_cycle = r.t_cycle;
      const IMC_State &s = *d->state;
      o.absorbed_E = s.g_absorbed_E; o.emission_E = s.g_emission_E; o.source_E = s.g_source_E;
      o.pre_census_E = s.g_pre_census_E; o.post_census_E = s.g_post_census_E; o.pre_mat_E = s.g_pre_mat_E;
      o.post_mat_E = s.g_post_mat_E; o.exit_E = s.g_exit_E;
      o.rad_conservation = s.rad_conservation; o.mat_conservation = s.mat_conservation;
      o.rad_balance_exact = r.rad_balance_exact;
      o.trans_particles = s.g_trans_particles; o.census_size = s.g_census_size;
      o.comb_n_before = r.comb_n_before; o.comb_n_after = r.comb_n_after;
      *out = o;
    }
  } catch (const std::exception &e) {
    d->err = e.what();
    return 1;
  }
  return 0;
}

int bhost_next_time_step(bhost_driver *d) {
  if (!d) return 1;
  d->state->next_time_step();
  return 0;
}

int bhost_get_array(const bhost_driver *d, const char *name, const double **data, uint64_t *n) {
  if (!d || !name || !data || !n) return 1;
  const std::string k(name);
  const std::vector<double> *v = nullptr;
  const Mesh &m = *d->mesh;
  if (d->driver && d->driver->mesh_on_device() &&
      (k == "T_e" || k == "T_r" || k == "f" || k == "op_a" || k == "op_s" || k == "E_emission" || k == "E_source" ||
       k == "E_census" || k == "abs_E" || k == "track_E")) {
    try {
      v = &d->driver->device_array(k);  // the cell state lives on the device: copy it over on request
    } catch (const std::exception &) {
      return 1;
    }
  } else try {
    if (k == "T_e") v = &m.get_T_e();
  else if (k == "T_r") v = &m.get_T_r();
  else if (k == "T_s") v = &m.get_T_s();
  else if (k == "f") v = &m.get_f();
  else if (k == "op_a") v = &m.get_op_a();
  else if (k == "op_s") v = &m.get_op_s();
  else if (k == "E_emission") v = &m.get_emission_E();
  else if (k == "E_source") v = &m.get_source_E();
  else if (k == "E_census") v = &m.get_census_E();
  else if (k == "x_faces") v = &m.get_x_faces();
  else if (k == "y_faces") v = &m.get_y_faces();
  else if (k == "z_faces") v = &m.get_z_faces();
  else if (k == "abs_E" && d->driver) v = &d->driver->get_last_abs_E();
  else if (k == "track_E" && d->driver) v = &d->driver->get_last_track_E();
  } catch (const std::exception &) {
    return 1;
  }
  if (!v) return 1;
  *data = v->data();
  *n = v->size();
  return 0;
}

int bhost_get_param(const bhost_driver *d, const char *name, double *value) {
  if (!d || !name || !value) return 1;
  const std::string k(name);
  const Input &in = *d->input;
  if (k == "n_cells") *value = d->mesh->get_n_global_cells();
  else if (k == "nx") *value = in.get_global_n_x_cells();
  else if (k == "ny") *value = in.get_global_n_y_cells();
  else if (k == "nz") *value = in.get_global_n_z_cells();
  else if (k == "n_user_photons") *value = (double)in.get_number_photons();
  else if (k == "seed") *value = in.get_rng_seed();
  else if (k == "dd_mode") *value = in.get_dd_mode();
  else if (k == "particle_algorithm") *value = d->params->get_transport_algorithm();
  else if (k == "particle_storage") *value = in.get_particle_storage();
  else if (k == "batch_size") *value = in.get_batch_size();
  else if (k == "n_omp_threads") *value = in.get_n_omp_threads();
  else if (k == "use_gpu_transporter") *value = in.get_use_gpu_transporter_bool();
  else if (k == "use_comb") *value = in.get_comb_bool();
  else if (k == "write_silo") *value = in.get_write_silo_bool();
  else if (k == "output_freq") *value = in.get_output_freq();
  else if (k == "t_start") *value = in.get_time_start();
  else if (k == "t_stop") *value = in.get_time_finish();
  else if (k == "dt") *value = d->state->get_dt();
  else if (k == "time") *value = d->state->get_time();
  else if (k == "step") *value = d->state->get_step();
  else if (k == "next_dt") *value = d->state->get_next_dt();
  else if (k == "t_mult") *value = in.get_time_mult();
  else if (k == "dt_max") *value = in.get_dt_max();
  else if (k == "T_source") *value = in.get_source_T();
  else if (k == "n_regions") *value = (double)in.get_regions().size();
  else if (k.rfind("bc", 0) == 0 && k.size() == 3 && k[2] >= '0' && k[2] <= '5') *value = in.get_bc(k[2] - '0');
  else if (k.rfind("region_of_cell:", 0) == 0) *value = d->mesh->get_region_ID((uint32_t)std::stoul(k.substr(15)));
  else return 1;
  return 0;
}

bgpu_ctx *bhost_gpu_ctx(bhost_driver *d) { return d && d->gpu ? d->gpu->get_ctx() : nullptr; }
double bhost_total_transport_time(const bhost_driver *d) { return d ? d->state->get_total_transport_time() : 0.0; }

}  // extern "C"
