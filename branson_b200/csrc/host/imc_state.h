// imc_state.h -- time stepping and conservation bookkeeping with the reference's IMC_State interface
// (src/imc_state.h).  Time stepping follows :112-134 and :292-296 exactly (next_dt is evaluated BEFORE time advances
// when it is used for census photons, src/replicated_transport.h:52); the conservation sums follow :207-258.
#pragma once
#include <cmath>
#include <cstdint>
#include <iostream>

#include "comm.h"
#include "constants.h"
#include "input.h"

namespace branson {

class IMC_State {
public:
  IMC_State(const Input &input, uint32_t rank_)
      : rank(rank_), m_dt(input.get_dt()), m_time(input.get_time_start()), m_time_stop(input.get_time_finish()),
        m_step(1), m_dt_mult(input.get_time_mult()), m_dt_max(input.get_dt_max()) {}

  double get_time() const { return m_time; }
  double get_dt() const { return m_dt; }
  uint32_t get_step() const { return m_step; }
  uint64_t get_transported_particles() const { return trans_particles; }
  uint64_t get_census_size() const { return census_size; }
  double get_pre_census_E() const { return pre_census_E; }
  double get_emission_E() const { return emission_E; }

  // src/imc_state.h:112-125
  double get_next_dt() const {
    double next_dt;
    if (m_dt * m_dt_mult < m_dt_max) next_dt = m_dt * m_dt_mult;
    else next_dt = m_dt_max;
    if (m_time + next_dt > m_time_stop) next_dt = m_time_stop - m_time;
    return next_dt;
  }
  // src/imc_state.h:127-134
  bool finished() const { return std::abs(m_time - m_time_stop) < 1.0e-8; }

  void print_timestep_header() const {
    std::cout << "****************************************";
    std::cout << "****************************************" << std::endl;
    std::cout << "Step: " << m_step << "  Start Time: " << m_time << "  End Time: ";
    std::cout << m_time + m_dt << "  dt: " << m_dt << std::endl;
  }
  void print_simulation_footer() const {
    std::cout << "****************************************";
    std::cout << "****************************************" << std::endl;
  }

  double get_rank_transport_runtime() const { return rank_transport_runtime; }
  double get_total_transport_time() const { return total_transport_time; }
  // src/imc_state.h:165-167 -- note: one cycle's user photon count over the transport time of ALL cycles
  double get_photons_per_second_fom(uint64_t photons) const { return (double)photons / total_transport_time; }
  uint64_t get_total_transported_particles() const { return total_trans_particles; }

  // src/imc_state.h:174-289: global sums, conservation residuals, printed by rank 0
  void print_conservation(const Comm &comm, bool print) {
    double d[8] = {absorbed_E, emission_E, source_E, pre_census_E, pre_mat_E, post_census_E, post_mat_E, exit_E};
    comm.sum(d, 8);
    double tmax = rank_transport_runtime, tmin = rank_transport_runtime;
    comm.max(&tmax, 1);
    comm.min(&tmin, 1);
    uint64_t u[2] = {trans_particles, census_size};
    comm.sum(u, 2);
    set_global_sums(d, u[0], u[1], tmax, tmin);
    finish_conservation(print);
  }

  // The same report from values that are already global: the device-mesh driver receives every rank's scalars with the
  // cycle's one all-reduce and forms these sums itself, in rank order (csrc/comm_native.cuh).
  // d = {absorbed, emission, source, pre_census, pre_mat, post_census, post_mat, exit}
  void set_global_sums(const double d[8], uint64_t trans, uint64_t census, double tmax, double tmin) {
    g_absorbed_E = d[0]; g_emission_E = d[1]; g_source_E = d[2]; g_pre_census_E = d[3];
    g_pre_mat_E = d[4]; g_post_census_E = d[5]; g_post_mat_E = d[6]; g_exit_E = d[7];
    g_trans_particles = trans;
    g_census_size = census;
    max_transport_time = tmax;
    min_transport_time = tmin;
  }
  void finish_conservation(bool print) {
    rad_conservation = (g_absorbed_E + g_post_census_E + g_exit_E) - (g_pre_census_E + g_emission_E + g_source_E);
    mat_conservation = g_post_mat_E - (g_pre_mat_E + g_absorbed_E - g_emission_E);
    total_trans_particles += g_trans_particles;
    if (rank == 0) {
      if (print) {
        using std::cout;
        using std::endl;
        cout << "Total Photons transported: " << g_trans_particles << endl;
        cout << "Emission E: " << g_emission_E << ", Source E: " << g_source_E << ", Absorption E: " << g_absorbed_E;
        cout << ", Exit E: " << g_exit_E << endl;
        cout << "Pre census E: " << g_pre_census_E << " Post census E: ";
        cout << g_post_census_E << " Post census Size: " << g_census_size << endl;
        cout << "Pre mat E: " << g_pre_mat_E << " Post mat E: " << g_post_mat_E << endl;
        cout << "Radiation conservation: " << rad_conservation << endl;
        cout << "Material conservation: " << mat_conservation << endl;
        cout << "Transport time max/min: " << max_transport_time << "/" << min_transport_time << endl;
      }
    }
    // (every rank keeps the run's transport time; the reference accumulates it on rank 0 only, src/imc_state.h:283-287)
    total_transport_time += max_transport_time;
  }

  // src/imc_state.h:292-296
  void next_time_step() {
    m_time += m_dt;
    m_dt = get_next_dt();
    m_step++;
  }

  void set_pre_census_E(double v) { pre_census_E = v; }
  void set_post_census_E(double v) { post_census_E = v; }
  void set_pre_mat_E(double v) { pre_mat_E = v; }
  void set_post_mat_E(double v) { post_mat_E = v; }
  void set_emission_E(double v) { emission_E = v; }
  void set_source_E(double v) { source_E = v; }
  void set_absorbed_E(double v) { absorbed_E = v; }
  void set_exit_E(double v) { exit_E = v; }
  void set_transported_particles(uint64_t v) { trans_particles = v; }
  void set_census_size(uint64_t v) { census_size = v; }
  void set_rank_transport_runtime(double v) { rank_transport_runtime = v; }

  // per-rank values
  double pre_census_E = 0, post_census_E = 0, pre_mat_E = 0, post_mat_E = 0, emission_E = 0, exit_E = 0,
         absorbed_E = 0, source_E = 0;
  // global values of the last print_conservation
  double g_absorbed_E = 0, g_emission_E = 0, g_source_E = 0, g_pre_census_E = 0, g_pre_mat_E = 0,
         g_post_census_E = 0, g_post_mat_E = 0, g_exit_E = 0, rad_conservation = 0, mat_conservation = 0;
  uint64_t g_trans_particles = 0, g_census_size = 0;
  double max_transport_time = 0, min_transport_time = 0;

private:
  uint32_t rank;
  double m_dt, m_time, m_time_stop;
  uint32_t m_step;
  double m_dt_mult, m_dt_max;
  uint64_t trans_particles = 0, census_size = 0, total_trans_particles = 0;
  double rank_transport_runtime = 0, total_transport_time = 0;
};

}  // namespace branson
