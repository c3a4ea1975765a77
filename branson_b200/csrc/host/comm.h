// comm.h -- the handful of collectives replicated mode needs, behind function pointers.
//
// The reference calls MPI directly (per cycle one scalar MPI_Allreduce in replicated_driver.h:58,
// one in mesh.h:293, abs_E / track_E in replicated_driver.h:91-94, m_emission_E in mesh.h:344 and the scalar
// reductions of imc_state.h:207-252).  There is no MPI on the B200 boxes: one process drives one GPU, the processes
// are started by torchrun and the collectives run over NCCL (NVLink) -- or gloo in the CPU tests.  The embedding
// harness supplies the callbacks; with none set the job is a single rank and every collective is the identity.
#pragma once
#include <cstddef>
#include <cstdint>

extern "C" {
typedef struct {
  void *user;
  // in-place sum of n host doubles over all ranks
  int (*allreduce_sum_f64)(void *user, double *buf, uint64_t n);
  // in-place sum of n DEVICE doubles (the packed tally buffer of bgpu_tally_buffer); stream-ordered after `stream`
  int (*allreduce_sum_f64_device)(void *user, void *device_ptr, uint64_t n, void *stream);
  int (*allreduce_sum_u64)(void *user, uint64_t *buf, uint64_t n);
  int (*allreduce_max_f64)(void *user, double *buf, uint64_t n);
  int (*allreduce_min_f64)(void *user, double *buf, uint64_t n);
  int (*barrier)(void *user);
} bhost_comm;
}

namespace branson {

class Comm {
public:
  Comm() : c{} {}
  Comm(const bhost_comm *cb, int rank_, int n_ranks_) : c{}, rank(rank_), n_ranks(n_ranks_) {
    if (cb) c = *cb;
  }
  int get_rank() const { return rank; }
  int get_n_rank() const { return n_ranks; }
  bool single() const { return n_ranks == 1; }
  bool has_device_allreduce() const { return c.allreduce_sum_f64_device != nullptr; }

  void sum(double *buf, uint64_t n) const {
    if (n_ranks > 1) check(c.allreduce_sum_f64 && !c.allreduce_sum_f64(c.user, buf, n));
  }
  void sum_device(void *dptr, uint64_t n, void *stream) const {
    if (n_ranks > 1) check(c.allreduce_sum_f64_device && !c.allreduce_sum_f64_device(c.user, dptr, n, stream));
  }
  void sum(uint64_t *buf, uint64_t n) const {
    if (n_ranks > 1) check(c.allreduce_sum_u64 && !c.allreduce_sum_u64(c.user, buf, n));
  }
  void max(double *buf, uint64_t n) const {
    if (n_ranks > 1) check(c.allreduce_max_f64 && !c.allreduce_max_f64(c.user, buf, n));
  }
  void min(double *buf, uint64_t n) const {
    if (n_ranks > 1) check(c.allreduce_min_f64 && !c.allreduce_min_f64(c.user, buf, n));
  }
  void barrier() const {
    if (n_ranks > 1 && c.barrier) check(!c.barrier(c.user));
  }

private:
  static void check(bool ok);
  bhost_comm c;
  int rank = 0, n_ranks = 1;
};

}  // namespace branson
