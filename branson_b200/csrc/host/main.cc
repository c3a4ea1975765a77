// main.cc -- command-line driver with the reference's shape (src/main.cc:39-154):
//
//   BRANSON <deck.xml> [n_groups] [host-mesh] [--ranks N]
//
// The reference is started as `mpirun -n N BRANSON deck.xml`, one MPI rank per process, N ranks sharing the photons of
// a REPLICATED run (src/main.cc:39-60, src/replicated_driver.h).  Here `--ranks N` plays the N ranks inside ONE process:
// one host thread per rank, rank r on GPU r % n_devices (the reference's map, src/gpu_setup.h:68-78), and the
// collectives native -- NCCL over NVLink when every rank has its own GPU, the in-process rank-ordered sum when ranks
// share one (csrc/comm_native.cuh).  Rank 0 prints, like the reference.  (bench.py and the multi-process tests start one
// process per GPU with torchrun instead and hand the NCCL unique id over; same library calls underneath.)
// Mesh physics run on the device by default; "host-mesh" keeps Mesh::calculate_photon_energy / update_temperature on
// the host.
#include <cstdlib>
#include <exception>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "comm.h"
#include "gpu_setup.h"
#include "imc_parameters.h"
#include "imc_state.h"
#include "input.h"
#include "mesh.h"
#include "replicated_driver.h"

namespace branson {
void Comm::check(bool ok) {
  if (!ok) throw std::runtime_error("collective failed or not provided");
}

// everything one rank owns (the reference's per-process objects, src/main.cc:62-128)
struct Rank {
  Rank(const std::string &deck, int rank, int n_ranks, uint32_t n_groups, bool force_replicated)
      : comm(nullptr, rank, n_ranks), input(deck, n_ranks, rank != 0), imc_p(prepare(input, force_replicated)),
        imc_state(input, (uint32_t)rank), mesh(input, imc_p, comm),
        gpu_setup(rank, n_ranks, imc_p.get_use_gpu_transporter_flag(), mesh, imc_p, n_groups) {}
  static Input &prepare(Input &in, bool force_replicated) {
    if (force_replicated) in.set_dd_mode(Constants::REPLICATED);
    return in;
  }
  Comm comm;
  Input input;
  IMC_Parameters imc_p;
  IMC_State imc_state;
  Mesh mesh;
  GPU_Setup gpu_setup;
};
}  // namespace branson

int main(int argc, char **argv) {
  using namespace branson;
  if (argc < 2) {
    std::cout << "Usage: BRANSON <path_to_input_file> [n_groups] [host-mesh] [--ranks N] [--replicated]" << std::endl;
    return EXIT_FAILURE;
  }
  uint32_t n_groups = 1u;
  int n_ranks = 1;
  bool force_replicated = false;
  Driver_Options opt;
  opt.mesh_on_device = true;
  for (int a = 2; a < argc; ++a) {
    const std::string s(argv[a]);
    if (s == "host-mesh") opt.mesh_on_device = false;
    else if (s == "--replicated") force_replicated = true;  // the multi-node deck names PARTICLE_PASS (exits in the reference)
    else if (s == "--ranks" && a + 1 < argc) n_ranks = std::atoi(argv[++a]);
    else if (a == 2) n_groups = (uint32_t)std::atoi(argv[a]);
    else {
      std::cout << "unknown argument: " << s << std::endl;
      return EXIT_FAILURE;
    }
  }
  if (n_ranks < 1 || n_groups < 1) {
    std::cout << "--ranks and n_groups must be positive" << std::endl;
    return EXIT_FAILURE;
  }
  try {
    std::cout << "----- Branson (B200 hot path), replicated IMC, " << n_ranks << " rank(s) -----" << std::endl;
    std::vector<std::unique_ptr<Rank>> ranks;
    for (int r = 0; r < n_ranks; ++r)
      ranks.push_back(std::make_unique<Rank>(argv[1], r, n_ranks, n_groups, force_replicated));
    if (n_ranks > 1) {
      std::vector<bgpu_ctx *> ctxs;
      for (auto &rk : ranks) ctxs.push_back(rk->gpu_setup.get_ctx());
      if (bgpu_comm_init_local(ctxs.data(), n_ranks)) {
        const char *e = bgpu_last_error(ctxs[0]);
        throw std::runtime_error(std::string("bgpu_comm_init_local: ") + ((e && *e) ? e : bgpu_last_error(nullptr)));
      }
      for (auto &rk : ranks) rk->comm.attach_native(rk->gpu_setup.get_ctx());
      int kind = 0;
      bgpu_comm_info(ctxs[0], &kind, nullptr, nullptr);
      std::cout << "collectives: " << (kind == BGPU_COMM_NCCL ? "NCCL, one GPU per rank" : "in-process, ranks share a GPU")
                << std::endl;
    }
    const double t0 = wall_now();
    std::mutex err_mutex;
    auto run = [&](int r) {
      try {
        Rank &rk = *ranks[(size_t)r];
        Driver_Options o = opt;
        o.print = r == 0;
        imc_replicated_driver(rk.mesh, rk.imc_state, rk.imc_p, rk.comm, rk.gpu_setup, o);
      } catch (const std::exception &e) {
        // the reference prints and MPI_Aborts (src/config.h.in:86-91): a rank that fails would leave the others waiting
        std::lock_guard<std::mutex> lk(err_mutex);
        std::cout << "rank " << r << ": " << e.what() << std::endl;
        std::exit(EXIT_FAILURE);
      }
    };
    std::vector<std::thread> threads;
    for (int r = 1; r < n_ranks; ++r) threads.emplace_back(run, r);
    run(0);
    for (auto &t : threads) t.join();
    IMC_State &imc_state = ranks[0]->imc_state;
    imc_state.print_simulation_footer();
    std::cout << "Total transport: " << imc_state.get_total_transport_time() << std::endl;
    std::cout << "Total time: " << wall_now() - t0 << std::endl;
    // src/main.cc:145-146, src/imc_state.h:165-167
    std::cout << "Photons Per Second (FOM): "
              << imc_state.get_photons_per_second_fom(ranks[0]->imc_p.get_n_user_photons()) << std::endl;
    std::cout << "Photon histories per second: "
              << (double)imc_state.get_total_transported_particles() / imc_state.get_total_transport_time() << std::endl;
  } catch (const std::exception &e) {
    std::cout << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return 0;
}
