// replicated_driver.h -- the replicated-mode cycle loop with the reference's driver shape (src/replicated_driver.h:33-122)
// and its transport seam (src/replicated_transport.h:33-158), calling the device through the C ABI only.
//
// Per cycle, in the reference's order:
//   mesh.calculate_photon_energy            (:53)   host, O(n_cells)
//   all-reduce global source energy         (:56-59)
//   GPU_Setup per-cycle part                (:64)   bgpu_set_cell_data: f, op_a, op_s -- 3 doubles per cell
//   make_initial_census_photons / make_photons / join_photon_arrays (:68-75)   bgpu_source, on the device
//   replicated_transport                    (:87-88) bgpu_transport: history loop + census compaction, on the device
//   all-reduce abs_E / track_E              (:91-94) ONE in-place NCCL all-reduce of the packed device tally buffer
//   mesh.update_temperature                 (:96)   host
//   rank != 0 zeroes its material sums, print_conservation, next_time_step (:100-120)
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "comm.h"
#include "gpu_setup.h"
#include "imc_parameters.h"
#include "imc_state.h"
#include "mesh.h"

namespace branson {

struct Cycle_Report {
  uint32_t step;
  double dt, time, next_dt, global_source_energy;
  bgpu_cycle_stats gpu;  // this rank's device statistics
  // wall-clock seconds of the host-visible phases of this cycle
  double t_calc_energy, t_cell_upload, t_source, t_transport, t_allreduce, t_tally_download, t_update_T, t_cycle;
  double rad_balance_exact;
  uint64_t comb_n_before = 0, comb_n_after = 0;  // global census size around this cycle's comb (both 0: none ran)
};

struct Driver_Options {
  int tally_mode = BGPU_TALLY_ATOMIC;
  bool print = true;
  // true: Mesh::calculate_photon_energy / update_temperature run on the device (bgpu_mesh_*): the cell state never
  // leaves HBM and only the running sums cross PCIe.  false: the host Mesh (bit-identical to the reference's host code)
  // with f / op_a / op_s / E arrays uploaded and the tallies downloaded every cycle.
  bool mesh_on_device = false;
  // > 0: population control -- when the global census exceeds this many photons after a cycle, comb it down to about
  // this many (comb_photons, src/census_functions.h:48-93, whose only trace in the reference's driver is the comment
  // of IMC_Parameters::use_comb_flag, src/imc_parameters.h:99: "Comb the census if great than n_user_photon after
  // cycle").  0 (default): never, like the reference, whose driver does not call its comb.
  uint64_t comb_max_census = 0;
  // true: the census is sorted by cell after every cycle (bgpu_sort_census_by_cell): memory
  // locality for the next cycle's transport; no per-photon result changes.  false (default): the reference's order.
  bool sort_census = false;
};

inline double wall_now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// replicated_transport (src/replicated_transport.h:33-158): runs the device transport + post-processing, fills the
// rank tallies and the IMC_State diagnostics.  The census stays on the device (the reference returns it by value).
inline void replicated_transport(const Mesh &mesh, const GPU_Setup &gpu_setup, IMC_State &imc_state,
                                 std::vector<double> &rank_abs_E, std::vector<double> &rank_track_E, const Comm &comm,
                                 const int transport_algorithm, const int tally_mode, Cycle_Report &rep) {
  const double next_dt = imc_state.get_next_dt();  // set for census photons (:52)
  bgpu_ctx *ctx = gpu_setup.get_ctx();
  const double t0 = wall_now();
  gpu_setup.check(bgpu_transport(ctx, next_dt, transport_algorithm == Constants::EVENT ? BGPU_EVENT : BGPU_HISTORY,
                                 tally_mode),
                  "bgpu_transport");
  const double t1 = wall_now();
  rep.t_transport = t1 - t0;
  // abs_E / track_E: summed over ranks on the device (NVLink), then handed to the host once
  if (!comm.single()) {
    if (comm.has_native()) {
      gpu_setup.check(bgpu_comm_allreduce_tallies(ctx, nullptr, nullptr), "bgpu_comm_allreduce_tallies");
    } else if (comm.has_device_allreduce()) {
      void *dptr = nullptr;
      uint64_t n = 0;
      gpu_setup.check(bgpu_tally_buffer(ctx, 0, &dptr, &n), "bgpu_tally_buffer");
      comm.sum_device(dptr, n, bgpu_stream(ctx));
    }
  }
  const double t2 = wall_now();
  rep.t_allreduce = t2 - t1;
  gpu_setup.check(bgpu_get_tallies(ctx, rank_abs_E.data(), rank_track_E.data(), &rep.gpu), "bgpu_get_tallies");
  if (!comm.single() && !comm.has_device_allreduce()) {
    // host fallback for the collective only (gloo in the CPU-side tests): same sums, host buffers
    comm.sum(rank_abs_E.data(), rank_abs_E.size());
    comm.sum(rank_track_E.data(), rank_track_E.size());
  }
  rep.t_tally_download = wall_now() - t2;
  (void)mesh;
  imc_state.set_exit_E(rep.gpu.exit_E);
  imc_state.set_post_census_E(rep.gpu.census_E);
  imc_state.set_census_size(rep.gpu.n_census);
  imc_state.set_rank_transport_runtime(rep.t_transport);
}

class Replicated_Driver {
public:
  Replicated_Driver(Mesh &mesh_, IMC_State &imc_state_, const IMC_Parameters &imc_p_, const Comm &comm_,
                    GPU_Setup &gpu_setup_, const Driver_Options &opt_)
      : mesh(mesh_), imc_state(imc_state_), imc_p(imc_p_), comm(comm_), gpu_setup(gpu_setup_), opt(opt_),
        abs_E(mesh_.get_n_global_cells(), 0.0), track_E(mesh_.get_n_global_cells(), 0.0) {
    if (opt.mesh_on_device) {
      // initialize_physical_properties (src/mesh.h:426-440) has run on the host Mesh; hand the static data over
      std::vector<bgpu_region> regions;
      for (const Region &r : mesh.get_regions())
        regions.push_back(bgpu_region{r.get_opac_A(), r.get_opac_B(), r.get_opac_C(), r.get_opac_S(), r.get_cV(),
                                      r.get_rho()});
      gpu_setup.check(bgpu_mesh_init(gpu_setup.get_ctx(), (uint32_t)regions.size(), regions.data(),
                                     mesh.get_region_index().data(), mesh.get_T_e().data(), mesh.get_T_r0().data(),
                                     mesh.get_T_s().data()),
                      "bgpu_mesh_init");
      mesh.set_device_resident();
    }
  }

  bool finished() const { return imc_state.finished(); }
  bool mesh_on_device() const { return opt.mesh_on_device; }

  // device-mesh runs: a host copy of one per-cell array (T_e T_r f op_a op_s E_emission E_source E_census abs_E track_E)
  const std::vector<double> &device_array(const std::string &name) {
    std::vector<double> &v = dev_cache[name];
    v.resize(mesh.get_n_global_cells());
    gpu_setup.check(bgpu_mesh_get(gpu_setup.get_ctx(), name.c_str(), v.data()), "bgpu_mesh_get");
    return v;
  }

  // One trip of the reference's while loop (src/replicated_driver.h:47-121) with the mesh physics on the device.
  // Multi-rank runs make ONE collective per cycle (csrc/comm_native.cuh): the source-energy sums of :56-59 and
  // src/mesh.h:291-294 are formed locally for every rank (all ranks hold the same cell state), and the tallies travel
  // together with every rank's conservation scalars in one packed in-place all-reduce that update_temperature consumes
  // on the stream -- one host synchronisation after transport.
  enum : int { RS_EMISSION = 0, RS_SOURCE, RS_PRE_CENSUS, RS_POST_CENSUS, RS_EXIT, RS_TRANSPORT_TIME, RS_TRANS_PARTICLES,
               RS_CENSUS_SIZE, RS_NEW_PHOTON_E, RS_USED };
  static_assert(RS_USED <= BGPU_RANK_SCALARS, "rank scalars exceed the tail row");

  Cycle_Report cycle_device_mesh() {
    Cycle_Report rep{};
    const int rank = comm.get_rank(), n_ranks = comm.get_n_rank();
    const double t_begin = wall_now();
    rep.step = imc_state.get_step();
    rep.dt = imc_state.get_dt();
    rep.time = imc_state.get_time();
    rep.next_dt = imc_state.get_next_dt();
    if (rank == 0 && opt.print) imc_state.print_timestep_header();
    bgpu_ctx *ctx = gpu_setup.get_ctx();
    if (n_ranks > 1 && !comm.has_native())
      throw GPU_Error("mesh_on_device needs the native communicator (bgpu_comm_init_rank / bgpu_comm_init_local) in "
                      "multi-rank runs");
    mesh.invalidate_mirror();

    // mesh.calculate_photon_energy (src/replicated_driver.h:53) incl. the replicated redistribution, every rank's totals
    rank_sums.assign((size_t)n_ranks, bgpu_mesh_sums{});
    gpu_setup.check(bgpu_mesh_calculate_photon_energy_replicated(ctx, imc_state.get_dt(), imc_state.get_step(),
                                                                 rank_sums.data()),
                    "bgpu_mesh_calculate_photon_energy_replicated");
    const bgpu_mesh_sums &mine = rank_sums[(size_t)rank];
    imc_state.set_pre_mat_E(mine.pre_mat_E);
    imc_state.set_emission_E(mine.emission_E);
    imc_state.set_source_E(mine.source_E);
    if (imc_state.get_step() == 1) imc_state.set_pre_census_E(mine.census_E);
    // global source energy (src/replicated_driver.h:56-59): the ranks' totals summed in rank order
    double global_source_energy = rank_sums[0].total_photon_E;
    for (int r = 1; r < n_ranks; ++r) global_source_energy += rank_sums[(size_t)r].total_photon_E;
    rep.global_source_energy = global_source_energy;
    const double t1 = wall_now();
    rep.t_calc_energy = t1 - t_begin;

    uint64_t n_new = 0, n_total = 0;
    gpu_setup.check(bgpu_mesh_source(ctx, imc_state.get_step(), global_source_energy, &n_new, &n_total),
                    "bgpu_mesh_source");
    bgpu_cycle_stats st{};
    gpu_setup.check(bgpu_get_tallies(ctx, nullptr, nullptr, &st), "bgpu_get_tallies");
    imc_state.set_pre_census_E(st.pre_census_E);
    const double t3 = wall_now();
    rep.t_source = t3 - t1;
    if (rank == 0 && opt.print) std::cout << "source time: " << rep.t_source << std::endl;
    imc_state.set_transported_particles(n_total);

    gpu_setup.check(bgpu_transport(ctx, imc_state.get_next_dt(),
                                   imc_p.get_transport_algorithm() == Constants::EVENT ? BGPU_EVENT : BGPU_HISTORY,
                                   opt.tally_mode),
                    "bgpu_transport");
    const double t5 = wall_now();
    rep.t_transport = t5 - t3;
    gpu_setup.check(bgpu_get_tallies(ctx, nullptr, nullptr, &rep.gpu), "bgpu_get_tallies");
    imc_state.set_exit_E(rep.gpu.exit_E);
    imc_state.set_post_census_E(rep.gpu.census_E);
    imc_state.set_census_size(rep.gpu.n_census);
    imc_state.set_rank_transport_runtime(rep.t_transport);

    // the cycle's one collective + mesh.update_temperature (src/replicated_driver.h:91-96)
    double mine_s[BGPU_RANK_SCALARS] = {};
    mine_s[RS_EMISSION] = mine.emission_E;
    mine_s[RS_SOURCE] = mine.source_E;
    mine_s[RS_PRE_CENSUS] = st.pre_census_E;
    mine_s[RS_POST_CENSUS] = rep.gpu.census_E;
    mine_s[RS_EXIT] = rep.gpu.exit_E;
    mine_s[RS_TRANSPORT_TIME] = rep.t_transport;
    mine_s[RS_TRANS_PARTICLES] = (double)n_total;
    mine_s[RS_CENSUS_SIZE] = (double)rep.gpu.n_census;
    mine_s[RS_NEW_PHOTON_E] = st.new_photon_E;
    all_scalars.assign((size_t)n_ranks * BGPU_RANK_SCALARS, 0.0);
    bgpu_mesh_sums sums{};
    gpu_setup.check(bgpu_mesh_finish_cycle(ctx, mine_s, all_scalars.data(), &sums), "bgpu_mesh_finish_cycle");
    imc_state.set_absorbed_E(sums.absorbed_E);
    imc_state.set_post_mat_E(sums.post_mat_E);
    rep.t_update_T = wall_now() - t5;

    // IMC_State::print_conservation (src/imc_state.h:207-252) from every rank's scalars, summed in rank order; the
    // material sums are rank 0's alone (src/replicated_driver.h:100-104) -- every rank holds the same values
    auto col = [&](int r, int k) { return all_scalars[(size_t)r * BGPU_RANK_SCALARS + k]; };
    double d[8] = {sums.absorbed_E, 0.0, 0.0, 0.0, mine.pre_mat_E, 0.0, sums.post_mat_E, 0.0};
    double trans = 0.0, census = 0.0, tmax = col(0, RS_TRANSPORT_TIME), tmin = tmax, rank_parts = 0.0;
    for (int r = 0; r < n_ranks; ++r) {
      d[1] += col(r, RS_EMISSION);
      d[2] += col(r, RS_SOURCE);
      d[3] += col(r, RS_PRE_CENSUS);
      d[5] += col(r, RS_POST_CENSUS);
      d[7] += col(r, RS_EXIT);
      trans += col(r, RS_TRANS_PARTICLES);
      census += col(r, RS_CENSUS_SIZE);
      tmax = std::max(tmax, col(r, RS_TRANSPORT_TIME));
      tmin = std::min(tmin, col(r, RS_TRANSPORT_TIME));
      rank_parts += col(r, RS_POST_CENSUS) + col(r, RS_EXIT) - col(r, RS_PRE_CENSUS) - col(r, RS_NEW_PHOTON_E);
    }
    rep.rad_balance_exact = sums.absorbed_E + rank_parts;  // absorbed_E is a tree sum here (mesh_dev.cuh)
    if (rank) {  // for replicated, just let root do conservation (:100-104)
      imc_state.set_absorbed_E(0.0);
      imc_state.set_pre_mat_E(0.0);
      imc_state.set_post_mat_E(0.0);
    }
    imc_state.set_global_sums(d, (uint64_t)trans, (uint64_t)census, tmax, tmin);
    imc_state.finish_conservation(opt.print);
    if (mesh.get_verbose_print()) {  // the reference's per-cycle temperature dump (src/mesh.h:364-381)
      mesh.mirror_from_device(device_array("T_e"), device_array("T_r"));
      mesh.print_verbose_block(device_array("abs_E"));
    }
    comb_census(rep);
    imc_state.next_time_step();
    rep.t_cycle = wall_now() - t_begin;
    return rep;
  }

  // Optional population control after a cycle (Driver_Options::comb_max_census).  Every rank combs its own census
  // against the global census energy, as comb_photons does (src/census_functions.h:61-65); the generator is
  // RNG(seed, 10^13 * step + 9 * 10^12 + rank): the photon streams of a step are 10^13 * step + n_user * rank + k
  // (src/source.h:221-222), so this stream is theirs for no photon while n_user * n_ranks < 9 * 10^12.
  void comb_census(Cycle_Report &rep) {
    comb_if_over(rep);
    // Optional locality ordering of what is left (Driver_Options::sort_census)
    if (opt.sort_census) gpu_setup.check(bgpu_sort_census_by_cell(gpu_setup.get_ctx()), "bgpu_sort_census_by_cell");
  }
  void comb_if_over(Cycle_Report &rep) {
    if (!opt.comb_max_census) return;
    bgpu_ctx *ctx = gpu_setup.get_ctx();
    uint64_t n_glob = bgpu_list_size(ctx, BGPU_LIST_CENSUS);
    comm.sum(&n_glob, 1);
    if (n_glob <= opt.comb_max_census) return;
    double census_E = 0.0;
    gpu_setup.check(bgpu_census_energy(ctx, &census_E), "bgpu_census_energy");
    comm.sum(&census_E, 1);
    const uint64_t stream = 10000000000000ull * imc_state.get_step() + 9000000000000ull + (uint64_t)comm.get_rank();
    bgpu_comb_stats cs{};
    gpu_setup.check(bgpu_comb_census(ctx, opt.comb_max_census, census_E, stream, &cs), "bgpu_comb_census");
    uint64_t n_after = cs.n_after;
    comm.sum(&n_after, 1);
    rep.comb_n_before = n_glob;
    rep.comb_n_after = n_after;
    if (comm.get_rank() == 0 && opt.print)
      std::cout << "census combed: " << n_glob << " -> " << n_after << " photons" << std::endl;
  }

  // one trip of the reference's while loop (src/replicated_driver.h:47-121)
  Cycle_Report cycle() {
    if (opt.mesh_on_device) return cycle_device_mesh();
    Cycle_Report rep{};
    const int rank = comm.get_rank();
    const double t_begin = wall_now();
    rep.step = imc_state.get_step();
    rep.dt = imc_state.get_dt();
    rep.time = imc_state.get_time();
    rep.next_dt = imc_state.get_next_dt();
    if (rank == 0 && opt.print) imc_state.print_timestep_header();

    mesh.calculate_photon_energy(imc_state, (uint32_t)imc_p.get_n_user_photons());
    double global_source_energy = mesh.get_total_photon_E();
    comm.sum(&global_source_energy, 1);
    rep.global_source_energy = global_source_energy;
    const double t1 = wall_now();
    rep.t_calc_energy = t1 - t_begin;

    bgpu_ctx *ctx = gpu_setup.get_ctx();
    gpu_setup.check(bgpu_set_cell_data(ctx, mesh.get_f().data(), mesh.get_op_a().data(), mesh.get_op_s().data()),
                    "bgpu_set_cell_data");
    const double t2 = wall_now();
    rep.t_cell_upload = t2 - t1;

    uint64_t n_new = 0, n_total = 0;
    gpu_setup.check(bgpu_source(ctx, imc_state.get_step(), imc_state.get_dt(), mesh.get_emission_E().data(),
                                mesh.get_source_E().data(),
                                imc_state.get_step() == 1 ? mesh.get_census_E().data() : nullptr,
                                global_source_energy, &n_new, &n_total),
                    "bgpu_source");
    bgpu_cycle_stats st{};
    gpu_setup.check(bgpu_get_tallies(ctx, nullptr, nullptr, &st), "bgpu_get_tallies");
    imc_state.set_pre_census_E(st.pre_census_E);
    const double t3 = wall_now();
    rep.t_source = t3 - t2;
    if (rank == 0 && opt.print) std::cout << "source time: " << rep.t_source << std::endl;
    imc_state.set_transported_particles(n_total);

    comm.barrier();
    replicated_transport(mesh, gpu_setup, imc_state, abs_E, track_E, comm, (int)imc_p.get_transport_algorithm(),
                         opt.tally_mode, rep);

    const double t4 = wall_now();
    {
      // Radiation balance from what the device made and tallied, with a compensated (Neumaier) sum over cells.  The
      // reference's own residual (IMC_State::print_conservation) adds abs_E serially in cell order
      // (src/mesh.h:359): with ~6e5 cells, addends below half an ulp of the running sum are dropped, which alone is
      // a relative 1e-12 (tools/debug_conservation.py) although the photons conserve energy exactly.
      double sum = 0.0, comp = 0.0;
      for (double v : abs_E) {
        const double t = sum + v;
        comp += (std::abs(sum) >= std::abs(v)) ? (sum - t) + v : (v - t) + sum;
        sum = t;
      }
      double rank_part = rep.gpu.census_E + rep.gpu.exit_E - rep.gpu.pre_census_E - st.new_photon_E;
      comm.sum(&rank_part, 1);
      rep.rad_balance_exact = (sum + comp) + rank_part;
    }
    last_abs_E = abs_E;
    last_track_E = track_E;
    mesh.update_temperature(abs_E, track_E, imc_state);
    rep.t_update_T = wall_now() - t4;

    comm.barrier();
    if (rank) {  // for replicated, just let root do conservation (:100-104)
      imc_state.set_absorbed_E(0.0);
      imc_state.set_pre_mat_E(0.0);
      imc_state.set_post_mat_E(0.0);
    }
    imc_state.print_conservation(comm, opt.print);
    comb_census(rep);
    imc_state.next_time_step();
    rep.t_cycle = wall_now() - t_begin;
    return rep;
  }

  const std::vector<double> &get_last_abs_E() const { return last_abs_E; }
  const std::vector<double> &get_last_track_E() const { return last_track_E; }

private:
  Mesh &mesh;
  IMC_State &imc_state;
  const IMC_Parameters &imc_p;
  const Comm &comm;
  GPU_Setup &gpu_setup;
  Driver_Options opt;
  std::vector<double> abs_E, track_E, last_abs_E, last_track_E;
  std::vector<bgpu_mesh_sums> rank_sums;
  std::vector<double> all_scalars;
  std::map<std::string, std::vector<double>> dev_cache;
};

// the reference's entry point: run all cycles (src/replicated_driver.h:33-122)
inline void imc_replicated_driver(Mesh &mesh, IMC_State &imc_state, const IMC_Parameters &imc_parameters,
                                  const Comm &comm, GPU_Setup &gpu_setup, const Driver_Options &opt = Driver_Options()) {
  Replicated_Driver drv(mesh, imc_state, imc_parameters, comm, gpu_setup, opt);
  while (!drv.finished()) drv.cycle();
}

}  // namespace branson
