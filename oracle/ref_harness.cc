// ref_harness.cc -- PARITY ORACLE, test infrastructure only.
//
// Drives the UNMODIFIED lanl/branson reference (headers included from
// /root/reference/src at build time; nothing is copied into this repo) through
// the same sequence of calls as imc_replicated_driver
// (reference src/replicated_driver.h:47-121) and dumps, per cycle and per
// rank, everything a parity test needs as raw binary doubles / integers -- the
// stock binary prints only 6-8 significant digits (src/mesh.h:365-416).
//
// Build: see oracle/Makefile (needs oracle/refshim/mpi.h and a generated
// config.h).  Output goes to oracle/_ref/ only.
//
// Usage: ref_harness <input.xml> <out_prefix> [max_cycles] [photon_dump_limit]
//   REF_DUMP_LEVEL=1 (full-size parity runs): only the per-photon integers of the post-transport list (cell, group, RNG
//   counter, descriptor), abs_E / T_e / T_r and the scalars are dumped -- no mesh tables, no pre-transport list, no
//   per-photon doubles
//   writes <out_prefix>.rank<r>.bin ; records are
//   [u32 name_len][name][u8 dtype][u64 count][payload]
//   dtype: 0=f64 1=u32 2=u64 3=u8
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// legs may execute this binary.

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <numeric>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <mpi.h>
#ifdef USE_OPENMP_IN_HARNESS
#include <omp.h>
#endif

// Read-only access to the reference's private diagnostics (IMC_State energies,
// RNG counter words).  Standard headers are all included above, so this only
// affects the reference's own class definitions; layouts are unchanged.
#define private public
#include "config.h"
#include "constants.h"
#include "imc_parameters.h"
#include "imc_state.h"
#include "info.h"
#include "input.h"
#include "mesh.h"
#include "mpi_types.h"
#include "census_functions.h"
#include "replicated_driver.h"
#include "timer.h"
#undef private

namespace {

struct Dump {
  FILE *f = nullptr;
  void open(const std::string &path) {
    f = std::fopen(path.c_str(), "wb");
    if (!f) {
      std::perror(path.c_str());
      std::exit(2);
    }
  }
  void rec(const std::string &name, uint8_t dtype, uint64_t count, const void *data, size_t elem) {
    uint32_t nl = (uint32_t)name.size();
    std::fwrite(&nl, 4, 1, f);
    std::fwrite(name.data(), 1, nl, f);
    std::fwrite(&dtype, 1, 1, f);
    std::fwrite(&count, 8, 1, f);
    if (count) std::fwrite(data, elem, count, f);
  }
  void f64(const std::string &n, const std::vector<double> &v) { rec(n, 0, v.size(), v.data(), 8); }
  void f64(const std::string &n, double v) { rec(n, 0, 1, &v, 8); }
  void u32(const std::string &n, const std::vector<uint32_t> &v) { rec(n, 1, v.size(), v.data(), 4); }
  void u64(const std::string &n, const std::vector<uint64_t> &v) { rec(n, 2, v.size(), v.data(), 8); }
  void u64(const std::string &n, uint64_t v) { rec(n, 2, 1, &v, 8); }
  void u8(const std::string &n, const std::vector<uint8_t> &v) { rec(n, 3, v.size(), v.data(), 1); }
  void close() {
    if (f) std::fclose(f);
    f = nullptr;
  }
};

static_assert(sizeof(Photon) == 120, "reference Photon layout changed");

int g_dump_level = 0;

void dump_photons(Dump &d, const std::string &pfx, const std::vector<Photon> &p, size_t limit, bool post) {
  size_t n = std::min(p.size(), limit);
  if (g_dump_level >= 1) {
    if (!post) return;
    std::vector<uint32_t> cell(n), group(n);
    std::vector<uint8_t> desc(n);
    std::vector<uint64_t> ctr(n);
    for (size_t i = 0; i < n; ++i) {
      cell[i] = p[i].m_cell_ID;
      group[i] = p[i].group;
      desc[i] = p[i].descriptors[0];
      ctr[i] = p[i].m_rng.data[0];
    }
    d.u32(pfx + "cell", cell);
    d.u32(pfx + "group", group);
    d.u64(pfx + "ctr", ctr);
    d.u8(pfx + "descriptor", desc);
    return;
  }
  std::vector<uint32_t> cell(n), group(n), stype(n);
  std::vector<uint8_t> desc(n);
  std::vector<uint64_t> ctr(n), stream(n);
  std::vector<double> pos(3 * n), ang(3 * n), E(n), E0(n), life(n);
  for (size_t i = 0; i < n; ++i) {
    const Photon &q = p[i];
    cell[i] = q.m_cell_ID;
    group[i] = q.group;
    stype[i] = q.source_type;
    desc[i] = q.descriptors[0];
    for (int k = 0; k < 3; ++k) {
      pos[3 * i + k] = q.m_pos[k];
      ang[3 * i + k] = q.m_angle[k];
    }
    E[i] = q.m_E;
    E0[i] = q.m_E0;
    life[i] = q.m_life_dx;
    ctr[i] = q.m_rng.data[0];
    stream[i] = q.m_rng.data[2];
  }
  d.u32(pfx + "cell", cell);
  d.u32(pfx + "group", group);
  d.u64(pfx + "ctr", ctr);
  d.f64(pfx + "pos", pos);
  d.f64(pfx + "angle", ang);
  d.f64(pfx + "E", E);
  d.f64(pfx + "life_dx", life);
  if (post) {
    d.u8(pfx + "descriptor", desc);
  } else {
    d.u32(pfx + "source_type", stype);
    d.u64(pfx + "stream", stream);
    d.f64(pfx + "E0", E0);
  }
}

} // namespace

int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  if (argc < 3) {
    std::cout << "usage: ref_harness <input.xml> <out_prefix> [max_cycles] [photon_dump_limit]" << std::endl;
    return 1;
  }
  const std::string filename(argv[1]);
  const std::string out_prefix(argv[2]);
  const uint32_t max_cycles = argc > 3 ? (uint32_t)std::atol(argv[3]) : 0xffffffffu;
  const size_t photon_limit = argc > 4 ? (size_t)std::atoll(argv[4]) : ~size_t(0);
  if (const char *lv = std::getenv("REF_DUMP_LEVEL")) g_dump_level = std::atoi(lv);
  const bool lean = g_dump_level >= 1;
  {
    const Info mpi_info;
    const int rank = mpi_info.get_rank();
    const int n_ranks = mpi_info.get_n_rank();
    MPI_Types mpi_types;
    Input input(filename, mpi_types);
    IMC_Parameters imc_p(input);
    IMC_State imc_state(input, rank);
    Mesh mesh(input, mpi_types, mpi_info, imc_p);
    mesh.initialize_physical_properties(input);
    MPI_Barrier(MPI_COMM_WORLD);
#ifdef USE_OPENMP
    omp_set_num_threads(input.get_n_omp_threads());
#endif
    if (input.get_dd_mode() != Constants::REPLICATED) {
      std::cout << "ref_harness only drives the replicated path" << std::endl;
      MPI_Abort(MPI_COMM_WORLD, 3);
    }

    Dump d;
    d.open(out_prefix + ".rank" + std::to_string(rank) + ".bin");
    const uint32_t n_cells = mesh.get_n_local_cells();
    d.u64("n_cells", n_cells);
    d.u64("n_ranks", (uint64_t)n_ranks);
    d.u64("n_groups", (uint64_t)BRANSON_N_GROUPS);
    d.u64("n_user_photons", imc_p.get_n_user_photons());
    d.u64("seed", imc_p.get_rng_seed());
    if (!lean) {
      std::vector<double> nodes(6 * (size_t)n_cells);
      std::vector<uint32_t> region(n_cells), enext(6 * (size_t)n_cells), bc(6 * (size_t)n_cells);
      for (uint32_t i = 0; i < n_cells; ++i) {
        const Cell &c = mesh.get_cell_ref(i);
        for (int k = 0; k < 6; ++k) {
          nodes[6 * (size_t)i + k] = c.nodes[k];
          enext[6 * (size_t)i + k] = c.e_next[k];
          bc[6 * (size_t)i + k] = (uint32_t)c.bc[k];
        }
        region[i] = c.get_region_ID();
      }
      d.f64("mesh/nodes", nodes);
      d.u32("mesh/region", region);
      d.u32("mesh/e_next", enext);
      d.u32("mesh/bc", bc);
    }

    // ---- the cycle loop: same calls, same order as
    // imc_replicated_driver<std::vector<Photon>> (replicated_driver.h:47-121)
    typedef std::vector<Photon> Census_T;
    std::vector<double> abs_E(mesh.get_n_global_cells(), 0.0);
    std::vector<double> track_E(mesh.get_n_global_cells(), 0.0);
    Census_T census_photons;
    const uint64_t n_user_photons = imc_p.get_n_user_photons();
    const uint32_t seed = imc_p.get_rng_seed();
    uint32_t cycles_done = 0;
    double total_transport = 0.0, total_source = 0.0;
    uint64_t total_histories = 0;

    while (!imc_state.finished() && cycles_done < max_cycles) {
      const std::string c = "c" + std::to_string(imc_state.get_step()) + "/";
      d.f64(c + "dt", imc_state.get_dt());
      d.f64(c + "time", imc_state.get_time());
      if (!lean) {
        std::vector<double> Te(n_cells);
        for (uint32_t i = 0; i < n_cells; ++i) Te[i] = mesh.get_cell_ref(i).get_T_e();
        d.f64(c + "T_e_pre", Te);
      }
      mesh.calculate_photon_energy(imc_state, n_user_photons);
      double global_source_energy = mesh.get_total_photon_E();
      d.f64(c + "rank_total_photon_E", global_source_energy);
      MPI_Allreduce(MPI_IN_PLACE, &global_source_energy, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
      d.f64(c + "global_source_energy", global_source_energy);
      if (!lean) {
        std::vector<double> f(n_cells), opa(n_cells), ops(n_cells);
        for (uint32_t i = 0; i < n_cells; ++i) {
          const Cell &cell = mesh.get_cell_ref(i);
          f[i] = cell.get_f();
          opa[i] = cell.get_op_a();
          ops[i] = cell.get_op_s();
        }
        d.f64(c + "f", f);
        d.f64(c + "op_a", opa);
        d.f64(c + "op_s", ops);
        d.f64(c + "E_emission", mesh.get_emission_E());
        d.f64(c + "E_census", mesh.get_census_E());
        d.f64(c + "E_source", mesh.get_source_E());
      }
      imc_state.set_pre_census_E(get_photon_list_E(census_photons));
#ifdef BRANSON_B200_DROPIN
      // the patched reference (oracle/dropin.patch): GPU_Setup holds the B200 device context, replicated_transport takes
      // its gpu_transport_photons branch (bgpu_transport_photons_aos) when the deck says use_gpu_transporter TRUE
      GPU_Setup gpu_setup(rank, n_ranks, imc_p.get_use_gpu_transporter_flag(), mesh.get_cells(), seed, n_user_photons);
#else
      GPU_Setup gpu_setup(rank, n_ranks, false, mesh.get_cells());
#endif

      auto t0 = std::chrono::high_resolution_clock::now();
      if (imc_state.get_step() == 1)
        census_photons = make_initial_census_photons<Census_T>(imc_state.get_dt(), mesh, rank, seed,
                                                               n_user_photons, global_source_energy);
      imc_state.set_pre_census_E(get_photon_list_E(census_photons));
      auto all_photons = make_photons<Census_T>(imc_state.get_dt(), mesh, rank, imc_state.get_step(), seed,
                                                n_user_photons, global_source_energy);
      const uint64_t n_new = all_photons.size();
      join_photon_arrays(all_photons, census_photons);
      auto t1 = std::chrono::high_resolution_clock::now();
      total_source += std::chrono::duration<double>(t1 - t0).count();
      imc_state.set_transported_particles(all_photons.size());
      d.u64(c + "n_new", n_new);
      d.u64(c + "n_photons", (uint64_t)all_photons.size());
      d.f64(c + "pre_census_E", imc_state.pre_census_E);
      dump_photons(d, c + "pre/", all_photons, photon_limit, false);
      MPI_Barrier(MPI_COMM_WORLD);

      d.f64(c + "next_dt", imc_state.get_next_dt());
      census_photons = replicated_transport<Census_T>(mesh, gpu_setup, imc_state, abs_E, track_E, all_photons,
                                                      imc_p.get_n_omp_threads(), imc_p.get_batch_size(),
                                                      imc_p.get_transport_algorithm());
      dump_photons(d, c + "post/", all_photons, photon_limit, true);
      d.u64(c + "n_census", (uint64_t)census_photons.size());
      d.f64(c + "exit_E", imc_state.exit_E);
      d.f64(c + "post_census_E", imc_state.post_census_E);
      d.f64(c + "transport_seconds", imc_state.rank_transport_runtime);
      total_transport += imc_state.rank_transport_runtime;
      total_histories += all_photons.size();
      if (!lean) {
        d.f64(c + "rank_abs_E", abs_E);
        d.f64(c + "rank_track_E", track_E);
      }

      MPI_Allreduce(MPI_IN_PLACE, &abs_E[0], mesh.get_n_global_cells(), MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
      MPI_Allreduce(MPI_IN_PLACE, &track_E[0], mesh.get_n_global_cells(), MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
      d.f64(c + "abs_E", abs_E);
      d.f64(c + "track_E", track_E);

      mesh.update_temperature(abs_E, track_E, imc_state);
      {
        std::vector<double> Te(n_cells), Tr(n_cells);
        for (uint32_t i = 0; i < n_cells; ++i) {
          Te[i] = mesh.get_cell_ref(i).get_T_e();
          Tr[i] = mesh.get_T_r(i);
        }
        d.f64(c + "T_e", Te);
        d.f64(c + "T_r", Tr);
      }
      MPI_Barrier(MPI_COMM_WORLD);
      if (rank) {
        imc_state.set_absorbed_E(0.0);
        imc_state.set_pre_mat_E(0.0);
        imc_state.set_post_mat_E(0.0);
      }
      d.f64(c + "emission_E", imc_state.emission_E);
      d.f64(c + "source_E", imc_state.source_E);
      d.f64(c + "absorbed_E", imc_state.absorbed_E);
      d.f64(c + "pre_mat_E", imc_state.pre_mat_E);
      d.f64(c + "post_mat_E", imc_state.post_mat_E);
      imc_state.print_conservation(imc_p.get_dd_mode());
      imc_state.next_time_step();
      ++cycles_done;
    }
    // Optional: the reference's population control, which exists as a function (src/census_functions.h:48-93) but has
    // no call site in this snapshot.  Run it once on the final census with RNG(seed, REF_COMB_STREAM) and dump the list
    // before and after, as the function-level oracle of the device comb.
    if (const char *cm = std::getenv("REF_COMB_MAX")) {
      const int64_t max_census = std::atoll(cm);
      const char *cs = std::getenv("REF_COMB_STREAM");
      const uint64_t stream = cs ? std::strtoull(cs, nullptr, 10) : 0ull;
      RNG rng(seed, stream);
      dump_photons(d, "comb/pre/", census_photons, ~size_t(0), false);
      {
        std::vector<uint64_t> st(census_photons.size());
        std::vector<double> e0(census_photons.size());
        for (size_t i = 0; i < st.size(); ++i) { st[i] = census_photons[i].m_rng.data[2]; e0[i] = census_photons[i].m_E0; }
        d.u64("comb/pre/stream", st);
        d.f64("comb/pre/E0", e0);
      }
      d.f64("comb/local_census_E", get_photon_list_E(census_photons));
      d.u64("comb/max_census_photons", (uint64_t)max_census);
      d.u64("comb/rng_stream", stream);
      comb_photons(census_photons, max_census, &rng);
      dump_photons(d, "comb/post/", census_photons, ~size_t(0), false);
      {
        std::vector<uint64_t> st(census_photons.size());
        for (size_t i = 0; i < st.size(); ++i) st[i] = census_photons[i].m_rng.data[2];
        d.u64("comb/post/stream", st);
      }
      d.u64("comb/rng_draws", rng.data[0]);
    }
    d.u64("cycles_done", cycles_done);
    d.close();
    if (rank == 0) {
      std::cout << "HARNESS_SUMMARY histories " << total_histories << " transport_s " << std::setprecision(9)
                << total_transport << " source_s " << total_source << " cycles " << cycles_done << std::endl;
    }
  }
  MPI_Barrier(MPI_COMM_WORLD);
  MPI_Finalize();
  return 0;
}
