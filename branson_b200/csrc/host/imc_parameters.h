// imc_parameters.h -- run parameters with the reference's IMC_Parameters interface (src/imc_parameters.h:25-95).
#pragma once
#include <cstdint>

#include "input.h"

namespace branson {

class IMC_Parameters {
public:
  explicit IMC_Parameters(const Input &input)
      : n_user_photons(input.get_number_photons()), seed(input.get_rng_seed()), dd_mode(input.get_dd_mode()),
        batch_size(input.get_batch_size()), particle_message_size(input.get_particle_message_size()),
        output_frequency(input.get_output_freq()), transport_algorithm(input.get_particle_algorithm()),
        n_omp_threads(input.get_n_omp_threads()), write_silo_flag(input.get_write_silo_bool()),
        use_gpu_transporter_flag(input.get_use_gpu_transporter_bool()), use_comb_flag(input.get_comb_bool()) {}

  uint64_t get_n_user_photons() const { return n_user_photons; }
  uint32_t get_rng_seed() const { return seed; }
  uint32_t get_dd_mode() const { return dd_mode; }
  uint32_t get_batch_size() const { return batch_size; }
  uint32_t get_particle_message_size() const { return particle_message_size; }
  bool get_write_silo_flag() const { return write_silo_flag; }
  bool get_use_gpu_transporter_flag() const { return use_gpu_transporter_flag; }
  bool get_use_comb_flag() const { return use_comb_flag; }
  uint32_t get_output_frequency() const { return output_frequency; }
  uint32_t get_n_omp_threads() const { return n_omp_threads; }
  uint32_t get_transport_algorithm() const { return transport_algorithm; }
  void set_transport_algorithm(uint32_t a) { transport_algorithm = a; }

private:
  uint64_t n_user_photons;
  uint32_t seed, dd_mode, batch_size, particle_message_size, output_frequency, transport_algorithm, n_omp_threads;
  bool write_silo_flag, use_gpu_transporter_flag, use_comb_flag;
};

}  // namespace branson
