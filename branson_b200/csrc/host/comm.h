// comm.h -- the handful of collectives replicated mode needs, behind function pointers.
//
// The reference calls MPI directly (per cycle one scalar MPI_Allreduce in replicated_driver.h:58,
// one in mesh.h:293, abs_E / track_E in replicated_driver.h:91-94, m_emission_E in mesh.h:344 and the scalar
// reductions of imc_state.h:207-252).  There is no MPI on the B200 boxes.  A rank is a device context; its collectives
// are native (csrc/comm_native.cuh behind bgpu_comm_*: NCCL over NVLink between GPUs, an in-process sum between ranks
// that share a GPU).  The callback table below remains for harnesses WITHOUT a device -- the CPU-only gloo tests of the
// host-side rank partitioning.
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../../include/branson_gpu.h"

#ifndef BHOST_COMM_DEFINED
#define BHOST_COMM_DEFINED
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
#endif

namespace branson {

// Collectives of the replicated cycle.  Three providers, in this order of preference:
//   native   the device context's own communicator (bgpu_comm_*: NCCL over NVLink, or the in-process sum of ranks that
//            share a device) -- attached once the GPU_Setup of this rank exists; what every GPU run uses;
//   callbacks  supplied by an embedding harness (the CPU-only gloo tests of the host partitioning, which have no
//            device context at all);
//   none     a single rank: every collective is the identity.
class Comm {
public:
  Comm() : c{} {}
  Comm(const bhost_comm *cb, int rank_, int n_ranks_) : c{}, rank(rank_), n_ranks(n_ranks_) {
    if (cb) c = *cb;
  }
  int get_rank() const { return rank; }
  int get_n_rank() const { return n_ranks; }
  bool single() const { return n_ranks == 1; }
  void attach_native(bgpu_ctx *ctx) { native = ctx; }
  bool has_native() const { return native != nullptr; }
  bgpu_ctx *native_ctx() const { return native; }
  bool has_device_allreduce() const { return native != nullptr || c.allreduce_sum_f64_device != nullptr; }

  void sum(double *buf, uint64_t n) const {
    if (n_ranks == 1) return;
    if (native) check(!bgpu_comm_allreduce_host(native, buf, n, BGPU_OP_SUM));
    else check(c.allreduce_sum_f64 && !c.allreduce_sum_f64(c.user, buf, n));
  }
  void sum_device(void *dptr, uint64_t n, void *stream) const {
    if (n_ranks > 1) check(c.allreduce_sum_f64_device && !c.allreduce_sum_f64_device(c.user, dptr, n, stream));
  }
  void sum(uint64_t *buf, uint64_t n) const {
    if (n_ranks == 1) return;
    if (native) {
      // photon counts: far below 2^53, exact as doubles
      double tmp[16];
      for (uint64_t off = 0; off < n; off += 16) {
        const uint64_t m = n - off < 16 ? n - off : 16;
        for (uint64_t i = 0; i < m; ++i) tmp[i] = (double)buf[off + i];
        check(!bgpu_comm_allreduce_host(native, tmp, m, BGPU_OP_SUM));
        for (uint64_t i = 0; i < m; ++i) buf[off + i] = (uint64_t)tmp[i];
      }
    } else {
      check(c.allreduce_sum_u64 && !c.allreduce_sum_u64(c.user, buf, n));
    }
  }
  void max(double *buf, uint64_t n) const {
    if (n_ranks == 1) return;
    if (native) check(!bgpu_comm_allreduce_host(native, buf, n, BGPU_OP_MAX));
    else check(c.allreduce_max_f64 && !c.allreduce_max_f64(c.user, buf, n));
  }
  void min(double *buf, uint64_t n) const {
    if (n_ranks == 1) return;
    if (native) check(!bgpu_comm_allreduce_host(native, buf, n, BGPU_OP_MIN));
    else check(c.allreduce_min_f64 && !c.allreduce_min_f64(c.user, buf, n));
  }
  void barrier() const {
    // (native: every collective already orders the ranks; the reference's MPI_Barrier calls only fence its timers)
    if (n_ranks > 1 && !native && c.barrier) check(!c.barrier(c.user));
  }

private:
  static void check(bool ok);
  bhost_comm c;
  int rank = 0, n_ranks = 1;
  bgpu_ctx *native = nullptr;
};

}  // namespace branson
