// gpu_setup.h -- host owner of the device context, with the role of the reference's GPU_Setup (src/gpu_setup.h:19-90):
// pick the device for this rank, put the mesh on it.  Unlike the reference, which re-uploads every 192..656-byte Cell
// each cycle and mallocs / frees photon storage around every launch (src/history_based_transport.h:368-410), the
// context is created once: geometry is static, only f / op_a / op_s (three doubles per cell) move per cycle and the
// census never leaves HBM.
#pragma once
#include <stdexcept>
#include <string>

#include "../../../include/branson_gpu.h"
#include "imc_parameters.h"
#include "mesh.h"

namespace branson {

class GPU_Error : public std::runtime_error {
public:
  using std::runtime_error::runtime_error;
};

class GPU_Setup {
public:
  GPU_Setup(const int rank, const int n_ranks, const bool use_gpu_transporter, const Mesh &mesh,
            const IMC_Parameters &imc_p, const uint32_t n_groups, const int device = -1)
      : m_use_gpu_transporter(use_gpu_transporter) {
    // there is no CPU transport in this code base: the flag is recorded (the decks set it) but the device is always used
    bgpu_mesh_desc d{};
    d.abi_version = BGPU_ABI_VERSION;
    d.n_groups = n_groups;
    d.nx = mesh.get_global_n_x();
    d.ny = mesh.get_global_n_y();
    d.nz = mesh.get_global_n_z();
    d.x_faces = mesh.get_x_faces().data();
    d.y_faces = mesh.get_y_faces().data();
    d.z_faces = mesh.get_z_faces().data();
    for (int i = 0; i < 6; ++i) d.bc[i] = (int32_t)mesh.get_bcs()[i];
    d.seed = imc_p.get_rng_seed();
    d.n_user_photons = imc_p.get_n_user_photons();
    d.rank = rank;
    d.n_ranks = n_ranks;
    d.device = device;
    d.photon_capacity = 0;
    if (bgpu_create(&ctx, &d)) throw GPU_Error(std::string("GPU_Setup: ") + bgpu_last_error(nullptr));
  }
  ~GPU_Setup() { bgpu_destroy(ctx); }
  GPU_Setup(const GPU_Setup &) = delete;
  GPU_Setup &operator=(const GPU_Setup &) = delete;

  bgpu_ctx *get_ctx() const { return ctx; }
  bool use_gpu_transporter() const { return m_use_gpu_transporter; }
  void check(int rc, const char *what) const {
    if (rc) throw GPU_Error(std::string(what) + ": " + bgpu_last_error(ctx));
  }

private:
  bool m_use_gpu_transporter;
  bgpu_ctx *ctx = nullptr;
};

}  // namespace branson
