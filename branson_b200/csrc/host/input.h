// input.h -- deck reader with the reference's Input interface (src/input.h) and parse rules.
//
// Same getters and the same decisions as the reference: unknown dd_transport_type -> PARTICLE_PASS, one rank ->
// REPLICATED (src/input.h:176-191); default particle_storage AOS, particle_algorithm HISTORY (:193-215); batch_size
// default 10000, forced to 100000000 for REPLICATED + HISTORY (:276-283,:488-490); the number of region_map entries
// must equal the number of division triples (:466); unknown boundary names abort (:420).  The reference broadcasts the
// parsed values from rank 0 (:510-635); here every process of a multi-GPU job reads the same file, which yields the
// same values without a broadcast.
#pragma once
#include <cmath>
#include <cstdint>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "constants.h"
#include "xml_lite.h"

namespace branson {

// src/region.h: user ID, material constants, opacity model sigma_a = A + B * T^C
class Region {
public:
  uint32_t get_ID() const { return ID; }
  double get_cV() const { return cV; }
  double get_rho() const { return rho; }
  double get_opac_A() const { return opac_A; }
  double get_opac_B() const { return opac_B; }
  double get_opac_C() const { return opac_C; }
  double get_opac_S() const { return opac_S; }
  double get_T_e() const { return T_e; }
  double get_T_r() const { return T_r; }
  // src/region.h:44-47
  double get_absorption_opacity(double T) const { return opac_A + opac_B * std::pow(T, opac_C); }
  double get_scattering_opacity() const { return opac_S; }

  uint32_t ID = 0;
  double cV = 0, rho = 0, opac_A = 0, opac_B = 0, opac_C = 0, opac_S = 0, T_e = 0, T_r = 0;
};

class Input_Error : public std::runtime_error {
public:
  using std::runtime_error::runtime_error;
};

class Input {
public:
  Input(const std::string &file_name, int n_ranks_ = 1, bool quiet_ = false) : n_ranks(n_ranks_), quiet(quiet_) {
    using namespace Constants;
    std::unique_ptr<xml_lite::Node> doc;
    try {
      doc = xml_lite::parse_file(file_name);
    } catch (const std::exception &e) {
      throw Input_Error(std::string("Improperly formatted xml file: ") + e.what());
    }
    const xml_lite::Node *proto = doc->child("prototype");
    if (!proto) throw Input_Error("'prototype' root element not found!");
    const xml_lite::Node *settings = proto->child("common");
    const xml_lite::Node *debug = proto->child("debug_options");
    const xml_lite::Node *spatial = proto->child("spatial");
    const xml_lite::Node *bc_node = proto->child("boundary");
    const xml_lite::Node *region_node = proto->child("regions");
    if (!settings) throw Input_Error("'common' section not found!");
    if (!spatial) throw Input_Error("'spatial' section not found!");
    if (!bc_node) throw Input_Error("'boundary' section not found!");
    if (!region_node) throw Input_Error("'regions' section not found!");

    tFinish = settings->as_double("t_stop");
    dt = settings->as_double("dt_start");
    tStart = settings->as_double("t_start");
    tMult = settings->as_double("t_mult");
    dtMax = settings->as_double("dt_max");
    n_photons = (uint64_t)settings->as_llong("photons");
    seed = (uint32_t)settings->as_int("seed");
    output_freq = (uint32_t)settings->as_int("output_frequency");

    std::string s = settings->child_value("use_gpu_transporter");
    if (s == "TRUE") use_gpu_transporter = true;
    else if (s == "FALSE") use_gpu_transporter = false;
    else { warn("\"use_gpu_transporter\" not found or recognized, defaulting to FALSE"); use_gpu_transporter = false; }

    s = settings->child_value("use_combing");
    if (s == "FALSE") use_comb = false;
    else if (s == "TRUE") use_comb = true;
    else { warn("\"use_combing\" not found or recognized, defaulting to TRUE"); use_comb = true; }

    write_silo = settings->child_value("write_silo") == "TRUE";

    s = settings->child_value("dd_transport_type");
    if (s == "PARTICLE_PASS") dd_mode = PARTICLE_PASS;
    else if (s == "REPLICATED") dd_mode = REPLICATED;
    else {
      warn("WARNING: Domain decomposition method not recognized or not set... setting to PARTICLE PASSING method");
      dd_mode = PARTICLE_PASS;
    }
    if (n_ranks == 1 && dd_mode == PARTICLE_PASS) {
      warn("WARNING: Domain decomposition method set to PARTICLE_PASS but there is only one rank, setting to REPLICATED");
      dd_mode = REPLICATED;
    }

    s = settings->child_value("particle_storage");
    if (s == "AOS") particle_storage = AOS;
    else if (s == "SOA") particle_storage = SOA;
    else { warn("WARNING: Particle storage type not recognized or not set... setting to AOS"); particle_storage = AOS; }

    s = settings->child_value("particle_algorithm");
    if (s == "EVENT") particle_algorithm = EVENT;
    else if (s == "HISTORY") particle_algorithm = HISTORY;
    else { warn("WARNING: Particle algorithm not recognized or not set... setting to HISTORY"); particle_algorithm = HISTORY; }

    n_omp_threads = settings->child("n_omp_threads") ? (uint32_t)settings->as_int("n_omp_threads") : 1u;

    s = settings->child_value("mesh_decomposition");
    if (s == "METIS") decomp_mode = METIS;
    else if (s == "CUBE") decomp_mode = CUBE;
    else if (dd_mode == REPLICATED) decomp_mode = NO_DECOMP;
    else { warn("WARNING: Mesh decomposition method is required but not recognized... setting to METIS method"); decomp_mode = METIS; }

    s = settings->child_value("particle_message_size");
    if (s.empty() && dd_mode == PARTICLE_PASS) particle_message_size = 10000;
    else particle_message_size = (uint32_t)settings->as_double("particle_message_size");

    s = settings->child_value("batch_size");
    if (!s.empty()) batch_size = (uint32_t)settings->as_llong("batch_size");
    else { warn("batch_size not found in settings, defaulting to 10000"); batch_size = 10000; }

    if (debug) {
      print_verbose = debug->child_value("print_verbose") == "TRUE";
      print_mesh_info = debug->child_value("print_mesh_info") == "TRUE";
    }

    for (const auto &it : spatial->children) {
      if (it->name == "x_division") {
        x_start.push_back(it->as_double("x_start"));
        x_end.push_back(it->as_double("x_end"));
        n_x_cells.push_back((uint32_t)it->as_int("n_x_cells"));
      } else if (it->name == "y_division") {
        y_start.push_back(it->as_double("y_start"));
        y_end.push_back(it->as_double("y_end"));
        n_y_cells.push_back((uint32_t)it->as_int("n_y_cells"));
      } else if (it->name == "z_division") {
        z_start.push_back(it->as_double("z_start"));
        z_end.push_back(it->as_double("z_end"));
        n_z_cells.push_back((uint32_t)it->as_int("n_z_cells"));
      } else if (it->name == "region_map") {
        // key = z*1e6 + y*1e3 + x (src/input.h:333-337)
        const uint32_t key = (uint32_t)it->as_int("z_div_ID") * 1000000u + (uint32_t)it->as_int("y_div_ID") * 1000u +
                             (uint32_t)it->as_int("x_div_ID");
        region_map[key] = (uint32_t)it->as_int("region_ID");
      }
    }

    bool source_on = false;
    const char *tags[6] = {"bc_left", "bc_right", "bc_down", "bc_up", "bc_bottom", "bc_top"};  // X_NEG .. Z_POS
    for (int d = 0; d < 6; ++d) {
      s = bc_node->child_value(tags[d]);
      if (s == "REFLECT") bc[d] = REFLECT;
      else if (s == "VACUUM") bc[d] = VACUUM;
      else if (s == "SOURCE") { bc[d] = SOURCE; source_on = true; }
      else throw Input_Error("ERROR: Boundary type not reconginzed. Exiting...");
    }
    if (source_on) T_source = bc_node->as_double("T_source");

    for (const auto &it : region_node->children) {
      if (it->name != "region") continue;
      Region r;
      r.ID = (uint32_t)it->as_int("ID");
      r.cV = it->as_double("CV");
      r.rho = it->as_double("density");
      r.opac_A = it->as_double("opacA");
      r.opac_B = it->as_double("opacB");
      r.opac_C = it->as_double("opacC");
      r.opac_S = it->as_double("opacS");
      r.T_e = it->as_double("initial_T_e");
      r.T_r = it->as_double("initial_T_r");
      region_ID_to_index[r.ID] = (uint32_t)regions.size();
      regions.push_back(r);
    }

    n_divisions = (uint32_t)(n_x_cells.size() * n_y_cells.size() * n_z_cells.size());
    n_global_x_cells = std::accumulate(n_x_cells.begin(), n_x_cells.end(), 0u);
    n_global_y_cells = std::accumulate(n_y_cells.begin(), n_y_cells.end(), 0u);
    n_global_z_cells = std::accumulate(n_z_cells.begin(), n_z_cells.end(), 0u);
    if (regions.empty()) throw Input_Error("ERROR: No regions were specified. Exiting...");
    if (n_divisions != region_map.size())
      throw Input_Error("ERROR: Number of total divisions must match the number of unique region maps. Exiting...");
    for (const auto &kv : region_map)
      if (!region_ID_to_index.count(kv.second))
        throw Input_Error("ERROR: region_map names a region ID that is not defined. Exiting...");
    if (dd_mode == REPLICATED && particle_algorithm == HISTORY) batch_size = 100000000;
  }

  // ---- getters, names as in the reference (src/input.h:700-900) ----
  uint32_t get_global_n_x_cells() const { return n_global_x_cells; }
  uint32_t get_global_n_y_cells() const { return n_global_y_cells; }
  uint32_t get_global_n_z_cells() const { return n_global_z_cells; }
  uint32_t get_n_x_divisions() const { return (uint32_t)n_x_cells.size(); }
  uint32_t get_n_y_divisions() const { return (uint32_t)n_y_cells.size(); }
  uint32_t get_n_z_divisions() const { return (uint32_t)n_z_cells.size(); }
  // src/input.h:788-798
  double get_dx(uint32_t div) const { return (x_end[div] - x_start[div]) / n_x_cells[div]; }
  double get_dy(uint32_t div) const { return (y_end[div] - y_start[div]) / n_y_cells[div]; }
  double get_dz(uint32_t div) const { return (z_end[div] - z_start[div]) / n_z_cells[div]; }
  double get_x_start(uint32_t div) const { return x_start[div]; }
  double get_y_start(uint32_t div) const { return y_start[div]; }
  double get_z_start(uint32_t div) const { return z_start[div]; }
  uint32_t get_x_division_cells(uint32_t div) const { return n_x_cells[div]; }
  uint32_t get_y_division_cells(uint32_t div) const { return n_y_cells[div]; }
  uint32_t get_z_division_cells(uint32_t div) const { return n_z_cells[div]; }
  const std::vector<Region> &get_regions() const { return regions; }
  const Region &get_region(uint32_t region_ID) const { return regions[region_ID_to_index.at(region_ID)]; }
  // src/input.h:871-875
  uint32_t get_region_index(uint32_t x_div, uint32_t y_div, uint32_t z_div) const {
    return region_ID_to_index.at(region_map.at(z_div * 1000000u + y_div * 1000u + x_div));
  }
  Constants::bc_type get_bc(int direction) const { return bc[direction]; }
  double get_source_T() const { return T_source; }
  double get_dt() const { return dt; }
  double get_time_start() const { return tStart; }
  double get_time_finish() const { return tFinish; }
  double get_time_mult() const { return tMult; }
  double get_dt_max() const { return dtMax; }
  uint64_t get_number_photons() const { return n_photons; }
  uint32_t get_rng_seed() const { return seed; }
  uint32_t get_output_freq() const { return output_freq; }
  uint32_t get_dd_mode() const { return dd_mode; }
  uint32_t get_decomposition_mode() const { return decomp_mode; }
  uint32_t get_particle_storage() const { return particle_storage; }
  uint32_t get_particle_algorithm() const { return particle_algorithm; }
  uint32_t get_n_omp_threads() const { return n_omp_threads; }
  uint32_t get_batch_size() const { return batch_size; }
  uint32_t get_particle_message_size() const { return particle_message_size; }
  bool get_write_silo_bool() const { return write_silo; }
  bool get_use_gpu_transporter_bool() const { return use_gpu_transporter; }
  bool get_comb_bool() const { return use_comb; }
  bool get_verbose_print_bool() const { return print_verbose; }
  bool get_print_mesh_info_bool() const { return print_mesh_info; }

  // test / bench overrides (the reference has none; its decks are edited instead)
  void set_number_photons(uint64_t n) { n_photons = n; }
  void set_time_finish(double t) { tFinish = t; }
  void set_dd_mode(uint32_t m) { dd_mode = m; }

private:
  void warn(const char *msg) const {
    if (!quiet) std::cout << msg << std::endl;
  }

  int n_ranks;
  bool quiet;
  double tStart = 0, dt = 0, tFinish = 0, tMult = 0, dtMax = 0, T_source = 0;
  uint64_t n_photons = 0;
  uint32_t seed = 0, output_freq = 0, dd_mode = 0, decomp_mode = 0, particle_storage = 0, particle_algorithm = 0;
  uint32_t n_omp_threads = 1, batch_size = 10000, particle_message_size = 0;
  bool use_gpu_transporter = false, use_comb = true, write_silo = false, print_verbose = false, print_mesh_info = false;
  std::vector<double> x_start, x_end, y_start, y_end, z_start, z_end;
  std::vector<uint32_t> n_x_cells, n_y_cells, n_z_cells;
  std::map<uint32_t, uint32_t> region_map, region_ID_to_index;
  Constants::bc_type bc[6] = {};
  std::vector<Region> regions;
  uint32_t n_divisions = 0, n_global_x_cells = 0, n_global_y_cells = 0, n_global_z_cells = 0;
};

}  // namespace branson
