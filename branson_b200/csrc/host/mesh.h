// mesh.h -- replicated rectilinear mesh with the reference's Mesh interface for the hot path's callers.
//
// The reference builds a vector<Cell> of 192..656-byte records (src/proto_mesh.h:106-218, src/cell.h:318-337).  The
// mesh is a tensor-product grid, so here it is stored the way the device wants it: three per-axis face arrays plus
// per-cell structure-of-arrays state.  Every arithmetic expression that feeds the device or the conservation sums
// keeps the reference's operand order, so on the same host the doubles are bit-identical:
//   faces        lo = start + i*dx, hi = start + (i+1)*dx, last cell of a non-final division ends at the next
//                division's start (src/proto_mesh.h:132-148, dx from src/input.h:788-798)
//   calculate_photon_energy   src/mesh.h:237-323  (opacity src/region.h:44-47, Fleck factor :267, E_emission :272,
//                E_census :274-277, E_source :280, replicated redistribution :291-315)
//   update_temperature        src/mesh.h:327-362,419-422
#pragma once
#include <cmath>
#include <cstdint>
#include <iomanip>
#include <iostream>
#include <stdexcept>
#include <vector>

#include "comm.h"
#include "constants.h"
#include "imc_parameters.h"
#include "imc_state.h"
#include "input.h"

namespace branson {

class Mesh {
public:
  Mesh(const Input &input, const IMC_Parameters &, const Comm &comm_)
      : ngx(input.get_global_n_x_cells()), ngy(input.get_global_n_y_cells()), ngz(input.get_global_n_z_cells()),
        n_global(ngz * ngy * ngx), n_cell(n_global), rank(comm_.get_rank()), n_ranks(comm_.get_n_rank()),
        verbose_print(input.get_verbose_print_bool()), comm(comm_), regions(input.get_regions()) {
    using namespace Constants;
    if (input.get_dd_mode() != REPLICATED)
      throw Input_Error("only REPLICATED transport is available (PARTICLE_PASS exits in the reference as well: "
                        "src/particle_pass_transport.h:141-142)");
    replicated = true;
    replicated_factor = 1.0 / static_cast<double>(n_ranks);  // src/mesh.h:87
    for (int d = 0; d < 6; ++d) bc[d] = input.get_bc(d);

    // per-axis faces and division index of every cell column
    build_axis(input.get_n_x_divisions(), [&](uint32_t d) { return input.get_x_start(d); },
               [&](uint32_t d) { return input.get_dx(d); }, [&](uint32_t d) { return input.get_x_division_cells(d); },
               x_faces, x_div_of);
    build_axis(input.get_n_y_divisions(), [&](uint32_t d) { return input.get_y_start(d); },
               [&](uint32_t d) { return input.get_dy(d); }, [&](uint32_t d) { return input.get_y_division_cells(d); },
               y_faces, y_div_of);
    build_axis(input.get_n_z_divisions(), [&](uint32_t d) { return input.get_z_start(d); },
               [&](uint32_t d) { return input.get_dz(d); }, [&](uint32_t d) { return input.get_z_division_cells(d); },
               z_faces, z_div_of);

    region_index.resize(n_cell);
    T_e.resize(n_cell);
    T_r0.resize(n_cell);
    T_s.assign(n_cell, 0.0);
    cV.resize(n_cell);
    rho.resize(n_cell);
    op_a.assign(n_cell, 0.0);
    op_s.assign(n_cell, 0.0);
    f.assign(n_cell, 0.0);
    T_r.assign(n_cell, 0.0);
    m_census_E.assign(n_cell, 0.0);
    m_emission_E.assign(n_cell, 0.0);
    m_source_E.assign(n_cell, 0.0);
    // global index = i + ngx*j + ngx*ngy*k (src/proto_mesh.h:115-206); initialize_physical_properties (src/mesh.h:426-440)
    uint32_t g = 0;
    for (uint32_t k = 0; k < ngz; ++k)
      for (uint32_t j = 0; j < ngy; ++j)
        for (uint32_t i = 0; i < ngx; ++i, ++g) {
          const uint32_t ri = input.get_region_index(x_div_of[i], y_div_of[j], z_div_of[k]);
          const Region &r = regions[ri];
          region_index[g] = ri;
          cV[g] = r.get_cV();
          T_e[g] = r.get_T_e();
          T_r0[g] = r.get_T_r();
          rho[g] = r.get_rho();
          if (get_source_face(g) != -1) T_s[g] = input.get_source_T();
        }
  }

  // ---- geometry ----
  uint32_t get_n_local_cells() const { return n_cell; }
  uint32_t get_n_global_cells() const { return n_global; }
  uint32_t get_global_n_x() const { return ngx; }
  uint32_t get_global_n_y() const { return ngy; }
  uint32_t get_global_n_z() const { return ngz; }
  const std::vector<double> &get_x_faces() const { return x_faces; }
  const std::vector<double> &get_y_faces() const { return y_faces; }
  const std::vector<double> &get_z_faces() const { return z_faces; }
  const Constants::bc_type *get_bcs() const { return bc; }
  void get_ijk(uint32_t cell, uint32_t &i, uint32_t &j, uint32_t &k) const {
    k = cell / (ngx * ngy);
    const uint32_t rem = cell - k * ngx * ngy;
    j = rem / ngx;
    i = rem - j * ngx;
  }
  // Cell::get_node_array order x_low x_high y_low y_high z_low z_high
  void get_nodes(uint32_t cell, double n[6]) const {
    uint32_t i, j, k;
    get_ijk(cell, i, j, k);
    n[0] = x_faces[i]; n[1] = x_faces[i + 1];
    n[2] = y_faces[j]; n[3] = y_faces[j + 1];
    n[4] = z_faces[k]; n[5] = z_faces[k + 1];
  }
  // src/cell.h:188-191
  double get_volume(uint32_t cell) const {
    double n[6];
    get_nodes(cell, n);
    return (n[1] - n[0]) * (n[3] - n[2]) * (n[5] - n[4]);
  }
  // boundary condition of a cell face: the deck's bc on domain faces, ELEMENT inside (src/proto_mesh.h:150-204)
  Constants::bc_type get_bc(uint32_t cell, int face) const {
    uint32_t i, j, k;
    get_ijk(cell, i, j, k);
    const bool on[6] = {i == 0, i == ngx - 1, j == 0, j == ngy - 1, k == 0, k == ngz - 1};
    return on[face] ? bc[face] : Constants::ELEMENT;
  }
  // src/cell.h:69-76
  int get_source_face(uint32_t cell) const {
    for (int s = 0; s < 6; ++s)
      if (get_bc(cell, s) == Constants::SOURCE) return s;
    return -1;
  }
  // src/cell.h:83-100 (+ get_source_area: area of the source face, -1.0 without one)
  double get_source_area(uint32_t cell) const {
    double n[6];
    get_nodes(cell, n);
    const int face = get_source_face(cell);
    if (face == 0 || face == 1) return (n[3] - n[2]) * (n[5] - n[4]);
    if (face == 2 || face == 3) return (n[1] - n[0]) * (n[5] - n[4]);
    if (face == 4 || face == 5) return (n[1] - n[0]) * (n[3] - n[2]);
    return -1.0;
  }
  uint32_t get_region_ID(uint32_t cell) const { return regions[region_index[cell]].get_ID(); }

  // ---- per-cycle physics ----
  void calculate_photon_energy(IMC_State &imc_state, const uint32_t n_user_photons) {
    using Constants::a;
    using Constants::c;
    total_photon_E = 0.0;
    const double dt = imc_state.get_dt();
    const uint32_t step = imc_state.get_step();
    double tot_census_E = 0.0, tot_emission_E = 0.0, tot_source_E = 0.0, pre_mat_E = 0.0;
    // per-cell values in parallel (each cell's doubles are independent of the loop order) ...
    mat_E_scratch.resize(n_cell);
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < n_cell; ++i) {
      const double vol = get_volume(i);
      const double T = T_e[i], Tr = T_r0[i], Ts = T_s[i];
      const Region &region = regions[region_index[i]];
      const double opa = region.get_absorption_opacity(T);
      const double ops = region.get_scattering_opacity();
      const double fleck = 1.0 / (1.0 + dt * opa * c * (4.0 * a * std::pow(T, 3) / (cV[i] * rho[i])));
      op_a[i] = opa;
      op_s[i] = ops;
      f[i] = fleck;
      m_emission_E[i] = replicated_factor * dt * vol * fleck * opa * a * c * std::pow(T, 4);
      if (step > 1) m_census_E[i] = 0.0;
      else m_census_E[i] = replicated_factor * vol * a * std::pow(Tr, 4);
      m_source_E[i] = replicated_factor * 0.25 * a * c * get_source_area(i) * std::pow(Ts, 4) * dt;
      mat_E_scratch[i] = T * cV[i] * vol * rho[i];
    }
    // ... and the running sums serially in cell order, like the reference's single loop (src/mesh.h:282-286)
    for (uint32_t i = 0; i < n_cell; ++i) {
      pre_mat_E += mat_E_scratch[i];
      tot_emission_E += m_emission_E[i];
      tot_census_E += m_census_E[i];
      tot_source_E += m_source_E[i];
      total_photon_E += m_source_E[i] + m_census_E[i] + m_emission_E[i];
    }
    if (replicated) {
      double global_source_E{tot_emission_E + tot_census_E + tot_source_E};
      comm.sum(&global_source_E, 1);
      tot_census_E = 0.0;
      tot_emission_E = 0.0;
      tot_source_E = 0.0;
      total_photon_E = 0.0;
      m_emission_E_global.resize(n_cell);
      for (uint32_t i = 0; i < n_cell; ++i) {
        const bool mine = (int)(i % (uint32_t)n_ranks) == rank;
        {
          // What the reference's MPI_Allreduce of m_emission_E (src/mesh.h:343-345) produces, computed locally: every
          // rank holds the same pre-redistribution share, and the redistribution test below is the same on all ranks,
          // so the rank-ordered sum is either the share added n_ranks times or the one un-split value.
          const double e = m_emission_E[i];
          if (e > 0.0 && int(n_user_photons * (e / global_source_E)) == 0) {
            m_emission_E_global[i] = e / replicated_factor;
            for (int r = 1; r < n_ranks; ++r) m_emission_E_global[i] = m_emission_E_global[i] + 0.0;
          } else {
            double sum = e;
            for (int r = 1; r < n_ranks; ++r) sum = sum + e;
            m_emission_E_global[i] = sum;
          }
        }
        if (step == 1 && m_census_E[i] > 0.0 && int(n_user_photons * (m_census_E[i] / global_source_E)) == 0)
          m_census_E[i] = mine ? m_census_E[i] / replicated_factor : 0.0;
        if (m_emission_E[i] > 0.0 && int(n_user_photons * (m_emission_E[i] / global_source_E)) == 0)
          m_emission_E[i] = mine ? m_emission_E[i] / replicated_factor : 0.0;
        if (m_source_E[i] > 0.0 && int(n_user_photons * (m_source_E[i] / global_source_E)) == 0)
          m_source_E[i] = mine ? m_source_E[i] / replicated_factor : 0.0;
        tot_emission_E += m_emission_E[i];
        tot_census_E += m_census_E[i];
        tot_source_E += m_source_E[i];
        total_photon_E += m_source_E[i] + m_census_E[i] + m_emission_E[i];
      }
    }
    imc_state.set_pre_mat_E(pre_mat_E);
    imc_state.set_emission_E(tot_emission_E);
    imc_state.set_source_E(tot_source_E);
    if (imc_state.get_step() == 1) imc_state.set_pre_census_E(tot_census_E);
  }

  // The reference first all-reduces m_emission_E over ranks (src/mesh.h:343-345); the same values were computed
  // locally in calculate_photon_energy (m_emission_E_global), so no collective is needed here.
  void update_temperature(std::vector<double> &abs_E, std::vector<double> &track_E, IMC_State &imc_state) {
    using Constants::a;
    using Constants::c;
    double total_abs_E = 0.0, total_post_mat_E = 0.0;
    if (replicated) m_emission_E = m_emission_E_global;
    mat_E_scratch.resize(n_cell);
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < n_cell; ++i) {
      const Region &region = regions[region_index[i]];
      const double cv = region.get_cV();
      const double rh = region.get_rho();
      const double vol = get_volume(i);
      const double T = T_e[i];
      const double T_new = T + (abs_E[i] - m_emission_E[i]) / (cv * vol * rh);
      T_r[i] = std::pow(track_E[i] / (vol * imc_state.get_dt() * a * c), 0.25);
      T_e[i] = T_new;
      mat_E_scratch[i] = T_new * cv * vol * rh;
    }
    for (uint32_t i = 0; i < n_cell; ++i) {  // serial, cell order (src/mesh.h:359-360)
      total_abs_E += abs_E[i];
      total_post_mat_E += mat_E_scratch[i];
    }
    print_verbose_block(abs_E);
    abs_E.assign(abs_E.size(), 0.0);
    track_E.assign(track_E.size(), 0.0);
    imc_state.set_absorbed_E(total_abs_E);
    imc_state.set_post_mat_E(total_post_mat_E);
  }

  // the reference's per-cycle temperature dump (src/mesh.h:364-381), deck option <print_verbose>
  bool get_verbose_print() const { return verbose_print; }
  void print_verbose_block(const std::vector<double> &abs_E) const {
    if (!verbose_print || rank != 0) return;
    using std::setw;
    std::cout.precision(8);
    std::cout << "-------- VERBOSE PRINT BLOCK: CELL TEMPERATURE --------" << std::endl;
    std::cout << std::right << setw(12) << "cell" << " " << setw(12) << "T_e" << " " << setw(12) << "T_r" << " "
              << setw(12) << "abs_E" << " " << std::endl;
    for (uint32_t i = 0; i < n_cell; ++i)
      std::cout << std::right << setw(12) << i << " " << setw(12) << T_e[i] << " " << setw(12) << T_r[i] << " "
                << setw(12) << abs_E[i] << " " << std::endl;
    std::cout << "-------------------------------------------------------" << std::endl;
  }

  // Device-mesh runs (Driver_Options::mesh_on_device): the cell state lives in HBM and this object's T_e / T_r / f / op /
  // E arrays stop being updated.  The driver marks the mesh, after which the state getters throw instead of handing out
  // stale values, unless the driver has mirrored the device arrays back for this cycle (mirror_from_device).
  void set_device_resident() { device_resident = true; state_current = false; }
  bool is_device_resident() const { return device_resident; }
  void mirror_from_device(const std::vector<double> &T_e_dev, const std::vector<double> &T_r_dev) {
    T_e = T_e_dev;
    T_r = T_r_dev;
    state_current = true;
  }
  void invalidate_mirror() { state_current = false; }

  // ---- what the sourcing / device side reads ----
  const std::vector<double> &get_census_E() const { need_host_state(false); return m_census_E; }
  const std::vector<double> &get_emission_E() const { need_host_state(false); return m_emission_E; }
  const std::vector<double> &get_source_E() const { need_host_state(false); return m_source_E; }
  double get_total_photon_E() const { return total_photon_E; }
  const std::vector<double> &get_f() const { need_host_state(false); return f; }
  const std::vector<double> &get_op_a() const { need_host_state(false); return op_a; }
  const std::vector<double> &get_op_s() const { need_host_state(false); return op_s; }
  const std::vector<double> &get_T_e() const { need_host_state(true); return T_e; }
  const std::vector<double> &get_T_r() const { need_host_state(true); return T_r; }
  const std::vector<double> &get_T_s() const { return T_s; }
  const std::vector<Region> &get_regions() const { return regions; }
  const std::vector<uint32_t> &get_region_index() const { return region_index; }
  const std::vector<double> &get_T_r0() const { return T_r0; }
  double get_replicated_factor() const { return replicated_factor; }
  int get_rank() const { return rank; }
  int get_n_ranks() const { return n_ranks; }

private:
  void need_host_state(bool mirrored_ok) const {
    if (device_resident && !(mirrored_ok && state_current))
      throw std::logic_error("Mesh: the cell state lives on the device in this run (mesh_on_device); read it through "
                             "Replicated_Driver::device_array / bhost_get_array");
  }
  template <class Start, class Delta, class Count>
  static void build_axis(uint32_t n_div, Start start, Delta delta, Count count, std::vector<double> &faces,
                         std::vector<uint32_t> &div_of) {
    faces.clear();
    div_of.clear();
    double last_hi = 0.0;
    for (uint32_t d = 0; d < n_div; ++d) {
      const double s = start(d), dx = delta(d);
      const uint32_t n = count(d);
      for (uint32_t i = 0; i < n; ++i) {
        faces.push_back(s + i * dx);  // low face of this cell == high face of the previous one (same expression,
                                      // or the next division's start for the last cell of a division)
        div_of.push_back(d);
        last_hi = s + (i + 1) * dx;
      }
    }
    faces.push_back(last_hi);
  }

  uint32_t ngx, ngy, ngz, n_global, n_cell;
  int rank, n_ranks;
  bool verbose_print, replicated = false;
  bool device_resident = false, state_current = false;
  const Comm &comm;
  double total_photon_E = 0.0, replicated_factor = 1.0;
  std::vector<Region> regions;
  Constants::bc_type bc[6];
  std::vector<double> x_faces, y_faces, z_faces;
  std::vector<uint32_t> x_div_of, y_div_of, z_div_of;
  std::vector<uint32_t> region_index;
  std::vector<double> T_e, T_r0, T_s, cV, rho, op_a, op_s, f, T_r;
  std::vector<double> m_census_E, m_emission_E, m_source_E, mat_E_scratch, m_emission_E_global;
};

}  // namespace branson
