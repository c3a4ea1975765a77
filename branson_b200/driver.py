"""ctypes binding of libbranson_host.so (include/branson_host.h): the C++ host layer -- Input, IMC_Parameters,
IMC_State, Mesh and the replicated cycle driver -- stepped one cycle at a time.

Harness plumbing only.  Multi-GPU: one process per GPU (torchrun).  The collectives of the cycle are native
(csrc/comm_native.cuh: ncclAllReduce on the ctx stream, called from C++); Python only carries the NCCL unique id from
rank 0 to the other processes (`init_nccl`), or groups several in-process ranks (`init_local`: one thread per rank).
`TorchComm` (gloo callbacks) remains for the CPU-only tests of the host-side rank partitioning.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import gpu

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(gpu.LIB_DIR, "libbranson_host.so")

EXPORTS = ["bhost_create", "bhost_destroy", "bhost_last_error", "bhost_finished", "bhost_calculate_photon_energy",
           "bhost_cycle", "bhost_next_time_step", "bhost_get_array", "bhost_get_param", "bhost_gpu_ctx", "bhost_total_transport_time",
           "bhost_comm_unique_id", "bhost_comm_init_rank", "bhost_comm_init_local"]
COMM_ID_BYTES = 128

_F_SUM = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_uint64)
_F_SUMDEV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p)
_F_SUMU = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_uint64)
_F_BAR = C.CFUNCTYPE(C.c_int, C.c_void_p)


class CommStruct(C.Structure):
    _fields_ = [("user", C.c_void_p), ("allreduce_sum_f64", _F_SUM), ("allreduce_sum_f64_device", _F_SUMDEV),
                ("allreduce_sum_u64", _F_SUMU), ("allreduce_max_f64", _F_SUM), ("allreduce_min_f64", _F_SUM),
                ("barrier", _F_BAR)]


class Options(C.Structure):
    _fields_ = [("n_groups", C.c_uint32), ("device", C.c_int32), ("tally_mode", C.c_int32), ("algorithm", C.c_int32),
                ("print", C.c_int32), ("validate", C.c_int32), ("no_gpu", C.c_int32),
                ("photons_override", C.c_uint64), ("t_stop_override", C.c_double), ("force_replicated", C.c_int32),
                ("mesh_on_device", C.c_int32), ("comb_max_census", C.c_uint64), ("sort_census", C.c_int32)]


class CycleReport(C.Structure):
    _fields_ = [("step", C.c_uint32), ("dt", C.c_double), ("time", C.c_double), ("next_dt", C.c_double),
                ("global_source_energy", C.c_double), ("gpu", gpu.CycleStats),
                ("t_calc_energy", C.c_double), ("t_cell_upload", C.c_double), ("t_source", C.c_double),
                ("t_transport", C.c_double), ("t_allreduce", C.c_double), ("t_tally_download", C.c_double),
                ("t_update_T", C.c_double), ("t_cycle", C.c_double),
                ("absorbed_E", C.c_double), ("emission_E", C.c_double), ("source_E", C.c_double),
                ("pre_census_E", C.c_double), ("post_census_E", C.c_double), ("pre_mat_E", C.c_double),
                ("post_mat_E", C.c_double), ("exit_E", C.c_double), ("rad_conservation", C.c_double),
                ("mat_conservation", C.c_double), ("rad_balance_exact", C.c_double), ("trans_particles", C.c_uint64), ("census_size", C.c_uint64),
                ("comb_n_before", C.c_uint64), ("comb_n_after", C.c_uint64)]

    def as_dict(self):
        d = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            d[k] = v.as_dict() if isinstance(v, gpu.CycleStats) else v
        return d


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C branson_b200/csrc`")
        gpu.lib()  # libbranson_gpu.so first (rpath $ORIGIN also finds it)
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.bhost_create.restype = vp
        L.bhost_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(Options), C.POINTER(CommStruct), C.c_char_p,
                                   C.c_size_t]
        L.bhost_destroy.argtypes = [vp]
        L.bhost_destroy.restype = None
        L.bhost_last_error.argtypes = [vp]
        L.bhost_last_error.restype = C.c_char_p
        L.bhost_finished.argtypes = [vp]
        L.bhost_calculate_photon_energy.argtypes = [vp, C.POINTER(C.c_double)]
        L.bhost_cycle.argtypes = [vp, C.POINTER(CycleReport)]
        L.bhost_next_time_step.argtypes = [vp]
        L.bhost_get_array.argtypes = [vp, C.c_char_p, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_uint64)]
        L.bhost_get_param.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double)]
        L.bhost_gpu_ctx.argtypes = [vp]
        L.bhost_gpu_ctx.restype = vp
        L.bhost_total_transport_time.argtypes = [vp]
        L.bhost_total_transport_time.restype = C.c_double
        L.bhost_comm_unique_id.argtypes = [C.c_char_p]
        L.bhost_comm_init_rank.argtypes = [vp, C.c_char_p]
        L.bhost_comm_init_local.argtypes = [C.POINTER(vp), C.c_int]
        _LIB = L
    return _LIB


class HostError(RuntimeError):
    pass


class TorchComm:
    """torch.distributed (gloo) collectives as comm.h callbacks, for host-only drivers (`no_gpu=True`): the CPU tests
    of the rank partitioning.  GPU runs do not use it: their collectives are native (init_nccl / init_local)."""

    def __init__(self, device="cpu"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.device = torch.device(device)
        if self.device.type != "cpu":
            raise ValueError("TorchComm carries host scalars of CPU-only runs; GPU ranks use driver.init_nccl")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self._cbs = [_F_SUM(self._sum), _F_SUMU(self._sum_u64), _F_SUM(self._max), _F_SUM(self._min),
                     _F_BAR(self._barrier)]
        self.struct = CommStruct(None, self._cbs[0], C.cast(None, _F_SUMDEV), self._cbs[1], self._cbs[2],
                                 self._cbs[3], self._cbs[4])

    def _host_reduce(self, buf, n, np_dtype, op):
        try:
            a = np.ctypeslib.as_array(buf, shape=(n,))
            t = self.torch.from_numpy(a.astype(np_dtype, copy=True))
            self.dist.all_reduce(t, op=op)
            a[:] = t.numpy().astype(a.dtype)
            return 0
        except Exception as e:  # pragma: no cover
            print(f"TorchComm: {e}", flush=True)
            return 1

    def _sum(self, _u, buf, n):
        return self._host_reduce(buf, n, np.float64, self.dist.ReduceOp.SUM)

    def _max(self, _u, buf, n):
        return self._host_reduce(buf, n, np.float64, self.dist.ReduceOp.MAX)

    def _min(self, _u, buf, n):
        return self._host_reduce(buf, n, np.float64, self.dist.ReduceOp.MIN)

    def _sum_u64(self, _u, buf, n):
        # counts stay far below 2^63: carried as int64
        return self._host_reduce(buf, n, np.int64, self.dist.ReduceOp.SUM)

    def _barrier(self, _u):
        try:
            self.dist.barrier()
            return 0
        except Exception as e:  # pragma: no cover
            print(f"TorchComm barrier: {e}", flush=True)
            return 1


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 of a multi-process run)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    if lib().bhost_comm_unique_id(buf):
        raise HostError("bhost_comm_unique_id: " + gpu.last_create_error())
    return buf.raw


def init_nccl(drv: "Driver", dist) -> None:
    """Multi-process bootstrap: rank 0's NCCL unique id travels through torch.distributed's store (any backend), every
    rank then creates its communicator natively (ncclCommInitRank on the driver's device context).  From here on the
    cycle's collectives never touch Python."""
    box = [comm_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    drv.comm_init_rank(box[0])


def init_local(drivers) -> None:
    """All ranks in this process (drivers[r] = rank r; step each from its own thread): NCCL when every rank has its own
    GPU, the in-process rank-ordered device sum when they share one."""
    arr = (C.c_void_p * len(drivers))(*[d._h for d in drivers])
    if lib().bhost_comm_init_local(arr, len(drivers)):
        raise HostError("bhost_comm_init_local: " + lib().bhost_last_error(drivers[0]._h).decode())


def run_ranks(drivers, fn):
    """fn(rank, driver) on one thread per in-process rank (the collectives rendezvous across the threads; ctypes calls
    release the GIL); returns the list of results, re-raising the first failure."""
    import threading
    out, errs = [None] * len(drivers), [None] * len(drivers)

    def work(r):
        try:
            out[r] = fn(r, drivers[r])
        except BaseException as e:  # noqa: BLE001 -- re-raised below
            errs[r] = e

    ts = [threading.Thread(target=work, args=(r,)) for r in range(len(drivers))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in errs:
        if e is not None:
            raise e
    return out


def device_tensor_f64(torch, dptr: int, n: int, device):
    """A torch float64 view of `n` doubles of device memory owned by the bgpu ctx (no copy)."""

    class _Arr:
        __cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(dptr), False), "version": 3,
                                    "strides": None}

    return torch.as_tensor(_Arr(), device=device)


class Driver:
    """Replicated cycle driver for one rank / one GPU."""

    def __init__(self, xml_path, n_groups=1, rank=0, n_ranks=1, device=-1, tally_mode=gpu.TALLY_ATOMIC, algorithm=-1,
                 print_report=False, validate=False, no_gpu=False, photons=0, t_stop=0.0, force_replicated=False,
                 comm: TorchComm | None = None, mesh_on_device=False, comb_max_census=0, sort_census=False):
        L = lib()
        o = Options(n_groups, device, tally_mode, algorithm, 1 if print_report else 0, 1 if validate else 0,
                    1 if no_gpu else 0, photons, t_stop, 1 if force_replicated else 0, 1 if mesh_on_device else 0,
                    int(comb_max_census), 1 if sort_census else 0)
        err = C.create_string_buffer(1024)
        self._comm = comm
        self._h = L.bhost_create(str(xml_path).encode(), rank, n_ranks, C.byref(o),
                                 C.byref(comm.struct) if comm is not None else None, err, 1024)
        if not self._h:
            raise HostError(err.value.decode())
        self.n_groups = n_groups

    def close(self):
        if getattr(self, "_h", None):
            lib().bhost_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init_rank(self, unique_id: bytes):
        if len(unique_id) != COMM_ID_BYTES:
            raise ValueError("NCCL unique id must be 128 bytes")
        if lib().bhost_comm_init_rank(self._h, unique_id):
            raise HostError(lib().bhost_last_error(self._h).decode())

    def finished(self) -> bool:
        return bool(lib().bhost_finished(self._h))

    def calculate_photon_energy(self) -> float:
        g = C.c_double()
        if lib().bhost_calculate_photon_energy(self._h, C.byref(g)):
            raise HostError(lib().bhost_last_error(self._h).decode())
        return g.value

    def cycle(self) -> dict:
        r = CycleReport()
        if lib().bhost_cycle(self._h, C.byref(r)):
            raise HostError(lib().bhost_last_error(self._h).decode())
        return r.as_dict()

    def next_time_step(self):
        lib().bhost_next_time_step(self._h)

    def array(self, name: str) -> np.ndarray:
        p, n = C.POINTER(C.c_double)(), C.c_uint64()
        if lib().bhost_get_array(self._h, name.encode(), C.byref(p), C.byref(n)):
            raise KeyError(name)
        if n.value == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def param(self, name: str) -> float:
        v = C.c_double()
        if lib().bhost_get_param(self._h, name.encode(), C.byref(v)):
            raise KeyError(name)
        return v.value

    def gpu_context(self) -> "GpuView":
        h = lib().bhost_gpu_ctx(self._h)
        if not h:
            raise HostError("driver was created without a GPU context")
        return GpuView(h, int(self.param("n_cells")), self.n_groups)

    def total_transport_time(self) -> float:
        return lib().bhost_total_transport_time(self._h)


class GpuView(gpu.Context):
    """A non-owning gpu.Context over the driver's bgpu_ctx (photon dumps in validation runs)."""

    def __init__(self, handle, n_cells, n_groups):  # noqa: super().__init__ intentionally not called
        self._h = C.c_void_p(handle)
        self.n_cells = n_cells
        self.n_groups = n_groups

    def close(self):
        self._h = None
