// comm_native.cuh -- the collectives of replicated mode, native: NCCL over NVLink / NVSwitch, or an in-process
// rank-ordered device sum where several ranks share one device.
//
// The reference calls MPI directly: per cycle MPI_Allreduce of the source energy (src/replicated_driver.h:56-59 and
// src/mesh.h:291-294), of abs_E and track_E (src/replicated_driver.h:91-94), of m_emission_E (src/mesh.h:343-345) and
// the eight scalar reductions of IMC_State::print_conservation (src/imc_state.h:207-252).  Here a rank is a bgpu_ctx
// (one per GPU), and a cycle needs ONE collective: the in-place sum of the packed tally buffer
//   {abs_E, track_E}[n_cells]  +  tail[n_ranks][BGPU_RANK_SCALARS]
// where every rank fills only its own row of the tail, so that the sum hands every rank every other rank's scalars --
// sums, maxima and minima over ranks are then formed locally in rank order (what the n-rank oracle's MPI shim does).
// The source-energy reductions need no collective at all: every rank holds the same cell state, so it can form every
// rank's share itself (mesh_dev.cuh, k_mesh_redistribute).
//
// Two back ends behind one call:
//   NCCL   one communicator per ctx.  Multi-process (torchrun: one process per GPU): rank 0 makes the unique id
//          (bgpu_comm_unique_id), the launcher hands it to every process, bgpu_comm_init_rank.  One process, one thread
//          per GPU (bin/branson --ranks N): bgpu_comm_init_local -> ncclCommInitAll.  libnccl.so.2 is dlopen'ed on
//          first use, so single-rank runs do not need it, and a process that already holds a copy (torch's) shares it.
//   LOCAL  ranks that share a device (more ranks than GPUs: the reference's rank % n_devices map, src/gpu_setup.h:68-78;
//          and the N-rank parity tests on a one-GPU box): the ranks' host threads meet at a barrier, the last one to
//          arrive launches k_reduce_ranks -- out = ((p0 + p1) + p2) + ... per element, written back to every rank's
//          buffer -- and every rank's stream waits for it.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

namespace bg {

// ---- libnccl, loaded on demand -----------------------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  std::string error;

  static NcclApi &get() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] { api.load(); });
    return api;
  }
  bool ok() const { return handle != nullptr; }

 private:
  void load() {
    // a copy the process already holds (torch's bundled libnccl.so.2) first, then the loader's search path
    handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) {
      const char *e = dlerror();
      error = std::string("cannot load libnccl.so.2: ") + (e ? e : "?");
      return;
    }
    bool all = true;
    auto sym = [&](auto &fn, const char *name) {
      fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(handle, name));
      if (!fn) {
        all = false;
        error = std::string("libnccl.so.2 lacks ") + name;
      }
    };
    sym(GetUniqueId, "ncclGetUniqueId");
    sym(CommInitRank, "ncclCommInitRank");
    sym(CommInitAll, "ncclCommInitAll");
    sym(CommDestroy, "ncclCommDestroy");
    sym(AllReduce, "ncclAllReduce");
    sym(GetErrorString, "ncclGetErrorString");
    sym(GetVersion, "ncclGetVersion");
    if (!all) handle = nullptr;
  }
};

// ---- LOCAL back end: ranks on one device ------------------------------------------------------------------------
constexpr int LOCAL_MAX_RANKS = 16;
struct RankPtrs {
  double *p[LOCAL_MAX_RANKS];
};
// every element reduced over the ranks in rank order, the result stored to every rank's buffer (one thread reads all
// of an element's addends before it writes any of them: in place is safe).  OP: 0 sum, 2 max, 3 min (ncclRedOp_t values)
template <int OP>
__global__ void k_reduce_ranks(const RankPtrs ptrs, const int n_ranks, const uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    double s = ptrs.p[0][i];
    for (int r = 1; r < n_ranks; ++r) {
      const double v = ptrs.p[r][i];
      if (OP == 0) s += v;
      else if (OP == 2) s = (v > s) ? v : s;
      else s = (v < s) ? v : s;
    }
    for (int r = 0; r < n_ranks; ++r) ptrs.p[r][i] = s;
  }
}

struct LocalGroup {
  int n_ranks = 0, device = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t generation = 0;
  RankPtrs ptrs{};
  uint64_t count[LOCAL_MAX_RANKS] = {};
  cudaEvent_t ev_ready[LOCAL_MAX_RANKS] = {};  // rank r's buffer is final on its stream
  cudaEvent_t ev_done = nullptr;               // the summed values are in every buffer
  cudaError_t status = cudaSuccess;            // of the launch, seen by every rank
  std::string error;
  ~LocalGroup() {
    for (cudaEvent_t e : ev_ready)
      if (e) cudaEventDestroy(e);
    if (ev_done) cudaEventDestroy(ev_done);
  }
};

enum : int { COMM_NONE = 0, COMM_NCCL = 1, COMM_LOCAL = 2 };

// what a ctx holds
struct CommHandle {
  int kind = COMM_NONE;
  ncclComm_t nccl = nullptr;
  std::shared_ptr<LocalGroup> local;
  uint64_t bytes = 0, calls = 0;  // all-reduced through this handle since creation
  void reset() {
    if (nccl) {
      NcclApi &api = NcclApi::get();
      if (api.ok()) api.CommDestroy(nccl);
      nccl = nullptr;
    }
    local.reset();
    kind = COMM_NONE;
  }
};

// in-place reduction of n doubles at `ptr` (device memory of this rank) over the group, ordered on `stream`; returns an
// error text, empty on success (a copy: the group's own message may be rewritten by the next collective)
inline std::string local_allreduce(LocalGroup &g, int rank, double *ptr, uint64_t n, int op, cudaStream_t stream) {
  cudaError_t e = cudaEventRecord(g.ev_ready[rank], stream);
  if (e != cudaSuccess) return cudaGetErrorString(e);
  {
    std::unique_lock<std::mutex> lk(g.m);
    g.ptrs.p[rank] = ptr;
    g.count[rank] = n;
    const uint64_t gen = g.generation;
    if (++g.arrived == g.n_ranks) {
      // the last rank to arrive reduces on its own stream, behind every rank's pending work
      g.status = cudaSuccess;
      g.error.clear();
      for (int r = 0; r < g.n_ranks && g.status == cudaSuccess; ++r) {
        if (g.count[r] != n) {
          g.status = cudaErrorInvalidValue;
          g.error = "ranks disagree on the all-reduce length";
        } else if (r != rank) {
          g.status = cudaStreamWaitEvent(stream, g.ev_ready[r], 0);
        }
      }
      if (g.status == cudaSuccess) {
        const unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8);
        if (op == 0) k_reduce_ranks<0><<<blocks ? blocks : 1, 256, 0, stream>>>(g.ptrs, g.n_ranks, n);
        else if (op == 2) k_reduce_ranks<2><<<blocks ? blocks : 1, 256, 0, stream>>>(g.ptrs, g.n_ranks, n);
        else k_reduce_ranks<3><<<blocks ? blocks : 1, 256, 0, stream>>>(g.ptrs, g.n_ranks, n);
        g.status = cudaGetLastError();
      }
      if (g.status == cudaSuccess) g.status = cudaEventRecord(g.ev_done, stream);
      if (g.status != cudaSuccess && g.error.empty()) g.error = cudaGetErrorString(g.status);
      g.arrived = 0;
      ++g.generation;
      g.cv.notify_all();
    } else {
      g.cv.wait(lk, [&] { return g.generation != gen; });
    }
    if (g.status != cudaSuccess) return g.error;
  }
  e = cudaStreamWaitEvent(stream, g.ev_done, 0);
  return e == cudaSuccess ? std::string() : std::string(cudaGetErrorString(e));
}

}  // namespace bg
