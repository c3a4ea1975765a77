// main.cc -- command-line driver with the reference's shape (src/main.cc:39-154): BRANSON <deck.xml> [n_groups] [host-mesh]
// (mesh physics on the device by default; "host-mesh" keeps Mesh::calculate_photon_energy / update_temperature on the host)
// Single process, one GPU (multi-GPU runs are launched one process per GPU by torchrun, see bench.py).
#include <cstdlib>
#include <iostream>
#include <string>

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
}  // namespace branson

int main(int argc, char **argv) {
  using namespace branson;
  if (argc < 2) {
    std::cout << "Usage: BRANSON <path_to_input_file> [n_groups] [host-mesh]" << std::endl;
    return EXIT_FAILURE;
  }
  const uint32_t n_groups = argc > 2 ? (uint32_t)std::atoi(argv[2]) : 1u;
  Driver_Options opt;
  opt.mesh_on_device = !(argc > 3 && std::string(argv[3]) == "host-mesh");
  try {
    Comm comm;
    std::cout << "----- Branson (B200 hot path), replicated IMC -----" << std::endl;
    Input input(argv[1], 1);
    IMC_Parameters imc_p(input);
    IMC_State imc_state(input, 0);
    Mesh mesh(input, imc_p, comm);
    GPU_Setup gpu_setup(0, 1, imc_p.get_use_gpu_transporter_flag(), mesh, imc_p, n_groups);
    const double t0 = wall_now();
    imc_replicated_driver(mesh, imc_state, imc_p, comm, gpu_setup, opt);
    imc_state.print_simulation_footer();
    std::cout << "Total transport: " << imc_state.get_total_transport_time() << std::endl;
    std::cout << "Total time: " << wall_now() - t0 << std::endl;
    // src/main.cc:145-146, src/imc_state.h:165-167
    std::cout << "Photons Per Second (FOM): " << imc_state.get_photons_per_second_fom(imc_p.get_n_user_photons())
              << std::endl;
    std::cout << "Photon histories per second: "
              << (double)imc_state.get_total_transported_particles() / imc_state.get_total_transport_time() << std::endl;
  } catch (const std::exception &e) {
    std::cout << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return 0;
}
