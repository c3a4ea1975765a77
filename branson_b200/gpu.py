"""ctypes binding of libbranson_gpu.so -- the C ABI declared in include/branson_gpu.h.

This is harness plumbing for tests / bench: the product is the CUDA library and the C++ host driver.  There is no
fallback of any kind: if the shared library is missing or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BRANSON_LIB_DIR: developer knob to load an experimental build of the same ABI (both libraries, built with
# `make -C branson_b200/csrc OUT=<dir> EXTRA=...`) for A/B timing; the tests and the bench use the in-tree build
LIB_DIR = os.environ.get("BRANSON_LIB_DIR") or _HERE
LIB_PATH = os.path.join(LIB_DIR, "libbranson_gpu.so")

ABI_VERSION = 2
HISTORY, EVENT = 0, 1
TALLY_ATOMIC, TALLY_DETERMINISTIC = 0, 1
LIST_WORK, LIST_CENSUS = 0, 1
BC = {"REFLECT": 0, "VACUUM": 1, "ELEMENT": 2, "SOURCE": 3, "PROCESSOR": 4}
EXIT, PASS, CENSUS, SCATTER, KILLED, BOUND = range(6)


class MeshDesc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("n_groups", C.c_uint32), ("nx", C.c_uint32), ("ny", C.c_uint32),
                ("nz", C.c_uint32), ("x_faces", C.c_void_p), ("y_faces", C.c_void_p), ("z_faces", C.c_void_p),
                ("bc", C.c_int32 * 6), ("seed", C.c_uint32), ("n_user_photons", C.c_uint64), ("rank", C.c_int32),
                ("n_ranks", C.c_int32), ("device", C.c_int32), ("photon_capacity", C.c_uint64)]


class CycleStats(C.Structure):
    _fields_ = [("census_E", C.c_double), ("exit_E", C.c_double), ("pre_census_E", C.c_double),
                ("new_photon_E", C.c_double),
                ("n_new", C.c_uint64), ("n_transported", C.c_uint64), ("n_census", C.c_uint64),
                ("n_killed", C.c_uint64), ("n_exit", C.c_uint64), ("n_events", C.c_uint64),
                ("n_scatters", C.c_uint64), ("n_crossings", C.c_uint64), ("n_reflections", C.c_uint64),
                ("n_deposits", C.c_uint64), ("n_group_lookups", C.c_uint64), ("n_launches", C.c_uint64),
                ("ms_source", C.c_float),
                ("ms_transport", C.c_float), ("ms_census", C.c_float), ("ms_total", C.c_float),
                ("transport_kernel", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CombStats(C.Structure):
    _fields_ = [("n_before", C.c_uint64), ("n_after", C.c_uint64), ("E_before", C.c_double), ("E_after", C.c_double),
                ("comb_photon_E", C.c_double), ("rng_draws", C.c_uint64)]


class PhotonSoA(C.Structure):
    _fields_ = [("n", C.c_uint64), ("cell", C.c_void_p), ("group", C.c_void_p), ("pos", C.c_void_p),
                ("angle", C.c_void_p), ("E", C.c_void_p), ("E0", C.c_void_p), ("life_dx", C.c_void_p),
                ("ctr", C.c_void_p), ("stream", C.c_void_p), ("descriptor", C.c_void_p), ("counters", C.c_void_p)]


_LIB = None

EXPORTS = [
    "bgpu_device_count", "bgpu_last_error", "bgpu_create", "bgpu_destroy", "bgpu_set_cell_data",
    "bgpu_set_cell_groups", "bgpu_source", "bgpu_transport", "bgpu_get_tallies", "bgpu_tally_buffer", "bgpu_sync",
    "bgpu_stream", "bgpu_device", "bgpu_transport_photons_aos", "bgpu_upload_photons", "bgpu_download_photons",
    "bgpu_list_size", "bgpu_enable_counters", "bgpu_set_launch", "bgpu_set_divergence", "bgpu_census_energy", "bgpu_comb_census", "bgpu_sort_census_by_cell", "bgpu_set_tally_copies", "bgpu_set_event_tail", "bgpu_set_group_walk", "bgpu_test_rng_draws", "bgpu_test_threefry", "bgpu_test_fastmath",
    "bgpu_mesh_init", "bgpu_mesh_calculate_photon_energy", "bgpu_mesh_redistribute", "bgpu_mesh_source",
    "bgpu_mesh_update_temperature", "bgpu_mesh_get", "bgpu_mesh_calculate_photon_energy_replicated",
    "bgpu_mesh_finish_cycle", "bgpu_comm_unique_id", "bgpu_comm_init_rank", "bgpu_comm_init_local", "bgpu_comm_info",
    "bgpu_comm_allreduce_host", "bgpu_comm_allreduce_tallies", "bgpu_set_event_mode", "bgpu_set_kernel",
]
KERNEL_AUTO, KERNEL_HISTORY, KERNEL_QUEUES = 0, 1, 2
COMM_NONE, COMM_NCCL, COMM_LOCAL = 0, 1, 2


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C branson_b200/csrc` "
                               "(__graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, u64, u32, i32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_double
        L.bgpu_device_count.restype = i32
        L.bgpu_last_error.restype = C.c_char_p
        L.bgpu_last_error.argtypes = [vp]
        L.bgpu_create.argtypes = [C.POINTER(vp), C.POINTER(MeshDesc)]
        L.bgpu_destroy.argtypes = [vp]
        L.bgpu_destroy.restype = None
        L.bgpu_set_cell_data.argtypes = [vp, vp, vp, vp]
        L.bgpu_set_cell_groups.argtypes = [vp, vp, vp, vp]
        L.bgpu_source.argtypes = [vp, u32, dbl, vp, vp, vp, dbl, C.POINTER(u64), C.POINTER(u64)]
        L.bgpu_transport.argtypes = [vp, dbl, i32, i32]
        L.bgpu_get_tallies.argtypes = [vp, vp, vp, C.POINTER(CycleStats)]
        L.bgpu_tally_buffer.argtypes = [vp, u64, C.POINTER(vp), C.POINTER(u64)]
        L.bgpu_sync.argtypes = [vp]
        L.bgpu_stream.argtypes = [vp]
        L.bgpu_stream.restype = vp
        L.bgpu_device.argtypes = [vp]
        L.bgpu_transport_photons_aos.argtypes = [vp, vp, u64, vp, i32, i32]
        L.bgpu_upload_photons.argtypes = [vp, i32, C.POINTER(PhotonSoA)]
        L.bgpu_download_photons.argtypes = [vp, i32, C.POINTER(PhotonSoA)]
        L.bgpu_list_size.argtypes = [vp, i32]
        L.bgpu_list_size.restype = u64
        L.bgpu_enable_counters.argtypes = [vp, i32]
        L.bgpu_set_launch.argtypes = [vp, i32, i32, i32]
        L.bgpu_set_divergence.argtypes = [vp, i32, i32]
        L.bgpu_census_energy.argtypes = [vp, C.POINTER(C.c_double)]
        L.bgpu_comb_census.argtypes = [vp, C.c_uint64, C.c_double, C.c_uint64, C.POINTER(CombStats)]
        L.bgpu_sort_census_by_cell.argtypes = [vp]
        L.bgpu_set_tally_copies.argtypes = [vp, i32]
        L.bgpu_set_event_tail.argtypes = [vp, u64]
        L.bgpu_set_event_mode.argtypes = [vp, i32, i32, i32]
        L.bgpu_set_kernel.argtypes = [vp, i32]
        L.bgpu_set_group_walk.argtypes = [vp, i32]
        L.bgpu_test_rng_draws.argtypes = [u32, u64, u32, vp]
        L.bgpu_test_threefry.argtypes = [vp, vp]
        L.bgpu_test_fastmath.argtypes = [C.c_int, u64, vp, vp, vp]
        L.bgpu_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(u64), C.POINTER(u64)]
        L.bgpu_comm_init_local.argtypes = [C.POINTER(vp), i32]
        L.bgpu_comm_allreduce_tallies.argtypes = [vp, vp, vp]
        L.bgpu_comm_allreduce_host.argtypes = [vp, vp, u64, i32]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class GpuError(RuntimeError):
    pass


def last_create_error() -> str:
    return (lib().bgpu_last_error(None) or b"").decode()


def comm_info(handle) -> dict:
    """back end (COMM_NONE / COMM_NCCL / COMM_LOCAL), number of collectives and bytes reduced through a ctx so far"""
    k, n, b = C.c_int(), C.c_uint64(), C.c_uint64()
    lib().bgpu_comm_info(handle, C.byref(k), C.byref(n), C.byref(b))
    return {"kind": k.value, "calls": n.value, "bytes": b.value}


def faces_from_nodes(nodes: np.ndarray, nx: int, ny: int, nz: int):
    """Per-axis face arrays from the reference's per-cell node table [n_cells][6] (src/proto_mesh.h:106-218)."""
    nodes = np.asarray(nodes, np.float64).reshape(-1, 6)
    xf = np.concatenate([nodes[:nx, 0], nodes[nx - 1:nx, 1]])
    yf = np.concatenate([nodes[0:nx * ny:nx, 2], nodes[nx * (ny - 1):nx * (ny - 1) + 1, 3]])
    zf = np.concatenate([nodes[0::nx * ny, 4], nodes[nx * ny * (nz - 1):nx * ny * (nz - 1) + 1, 5]])
    return xf.copy(), yf.copy(), zf.copy()


class Context:
    """One device context (bgpu_ctx)."""

    def __init__(self, n_groups, nx, ny, nz, x_faces, y_faces, z_faces, bc, seed, n_user_photons, rank=0, n_ranks=1,
                 device=-1, photon_capacity=0):
        L = lib()
        self._keep = [_f64(x_faces), _f64(y_faces), _f64(z_faces)]
        assert len(self._keep[0]) == nx + 1 and len(self._keep[1]) == ny + 1 and len(self._keep[2]) == nz + 1
        d = MeshDesc()
        d.abi_version, d.n_groups, d.nx, d.ny, d.nz = ABI_VERSION, n_groups, nx, ny, nz
        d.x_faces, d.y_faces, d.z_faces = (_ptr(a) for a in self._keep)
        for i, b in enumerate(bc):
            d.bc[i] = BC[b] if isinstance(b, str) else int(b)
        d.seed, d.n_user_photons, d.rank, d.n_ranks, d.device = seed, n_user_photons, rank, n_ranks, device
        d.photon_capacity = photon_capacity
        h = C.c_void_p()
        if L.bgpu_create(C.byref(h), C.byref(d)):
            raise GpuError(L.bgpu_last_error(None).decode())
        self._h = h
        self.n_cells = nx * ny * nz
        self.n_groups = n_groups

    def _ck(self, rc):
        if rc:
            raise GpuError(lib().bgpu_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().bgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_cell_data(self, f, op_a, op_s):
        f, a, s = _f64(f), _f64(op_a), _f64(op_s)
        assert f.size == a.size == s.size == self.n_cells
        self._ck(lib().bgpu_set_cell_data(self._h, _ptr(f), _ptr(a), _ptr(s)))

    def set_cell_groups(self, f, abs_groups, sct_groups):
        f, a, s = _f64(f), _f64(abs_groups), _f64(sct_groups)
        assert f.size == self.n_cells and a.size == s.size == self.n_cells * self.n_groups
        self._ck(lib().bgpu_set_cell_groups(self._h, _ptr(f), _ptr(a), _ptr(s)))

    def source(self, cycle, dt, E_emission, E_source, E_census, total_E):
        em, so = _f64(E_emission), _f64(E_source)
        ce = _f64(E_census) if E_census is not None else None
        n_new, n_tot = C.c_uint64(), C.c_uint64()
        self._ck(lib().bgpu_source(self._h, cycle, dt, _ptr(em), _ptr(so), _ptr(ce), total_E, C.byref(n_new),
                                   C.byref(n_tot)))
        return n_new.value, n_tot.value

    def transport(self, next_dt, algorithm=HISTORY, tally_mode=TALLY_ATOMIC):
        self._ck(lib().bgpu_transport(self._h, next_dt, algorithm, tally_mode))

    def tallies(self):
        a, t = np.zeros(self.n_cells), np.zeros(self.n_cells)
        st = CycleStats()
        self._ck(lib().bgpu_get_tallies(self._h, _ptr(a), _ptr(t), C.byref(st)))
        return a, t, st.as_dict()

    def stats(self):
        st = CycleStats()
        self._ck(lib().bgpu_get_tallies(self._h, None, None, C.byref(st)))
        return st.as_dict()

    def tally_buffer(self, extra=0):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(lib().bgpu_tally_buffer(self._h, extra, C.byref(p), C.byref(n)))
        return p.value, n.value

    def sync(self):
        self._ck(lib().bgpu_sync(self._h))

    @property
    def device(self):
        return lib().bgpu_device(self._h)

    def enable_counters(self, on=True):
        self._ck(lib().bgpu_enable_counters(self._h, 1 if on else 0))

    def set_launch(self, block_threads=0, blocks_per_sm=0, chunk=0):
        self._ck(lib().bgpu_set_launch(self._h, block_threads, blocks_per_sm, chunk))

    def set_divergence(self, scatter_batch=0, aggregate=-1):
        self._ck(lib().bgpu_set_divergence(self._h, scatter_batch, aggregate))

    def set_tally_copies(self, copies=0):
        self._ck(lib().bgpu_set_tally_copies(self._h, copies))

    def census_energy(self) -> float:
        e = C.c_double()
        self._ck(lib().bgpu_census_energy(self._h, C.byref(e)))
        return e.value

    def sort_census_by_cell(self) -> None:
        """bgpu_sort_census_by_cell: stable sort of the device census by cell."""
        self._ck(lib().bgpu_sort_census_by_cell(self._h))

    def comb_census(self, max_census_photons: int, global_census_E: float = 0.0, rng_stream: int = 0) -> dict:
        """comb_photons (reference src/census_functions.h:48-93) on the device census."""
        st = CombStats()
        self._ck(lib().bgpu_comb_census(self._h, max_census_photons, global_census_E, rng_stream, C.byref(st)))
        return {k: getattr(st, k) for k, _ in CombStats._fields_}

    def set_event_tail(self, n_active):
        self._ck(lib().bgpu_set_event_tail(self._h, n_active))

    def set_kernel(self, choice=KERNEL_AUTO):
        """BGPU_HISTORY: auto / always the history kernel / always the event-queue kernel"""
        self._ck(lib().bgpu_set_kernel(self._h, choice))

    def set_event_mode(self, hbm_passes=-1, batch_scatter=0, batch_refill=0):
        """BGPU_EVENT: shared-memory event queues (default) or the HBM-pass form; queue election thresholds"""
        self._ck(lib().bgpu_set_event_mode(self._h, hbm_passes, batch_scatter, batch_refill))

    def set_group_walk(self, closed_form=True):
        self._ck(lib().bgpu_set_group_walk(self._h, 1 if closed_form else 0))

    def list_size(self, which=LIST_WORK):
        return lib().bgpu_list_size(self._h, which)

    def upload(self, which, cell, group, pos, angle, E, E0, life_dx, ctr, stream):
        n = len(cell)
        arrs = dict(cell=np.ascontiguousarray(cell, np.uint32), group=np.ascontiguousarray(group, np.uint32),
                    pos=_f64(pos).reshape(-1), angle=_f64(angle).reshape(-1), E=_f64(E), E0=_f64(E0),
                    life_dx=_f64(life_dx), ctr=np.ascontiguousarray(ctr, np.uint64),
                    stream=np.ascontiguousarray(stream, np.uint64))
        s = PhotonSoA()
        s.n = n
        for k, v in arrs.items():
            setattr(s, k, _ptr(v))
        self._ck(lib().bgpu_upload_photons(self._h, which, C.byref(s)))

    def download(self, which=LIST_WORK, counters=False):
        n = self.list_size(which)
        out = dict(cell=np.zeros(n, np.uint32), group=np.zeros(n, np.uint32), pos=np.zeros(3 * n), angle=np.zeros(3 * n),
                   E=np.zeros(n), E0=np.zeros(n), life_dx=np.zeros(n), ctr=np.zeros(n, np.uint64),
                   stream=np.zeros(n, np.uint64))
        if which == LIST_WORK:
            out["descriptor"] = np.zeros(n, np.uint8)
            if counters:
                out["counters"] = np.zeros(4 * n, np.uint32)
        s = PhotonSoA()
        s.n = n
        for k, v in out.items():
            setattr(s, k, _ptr(v))
        self._ck(lib().bgpu_download_photons(self._h, which, C.byref(s)))
        return out

    def transport_photons_aos(self, photons: np.ndarray, cell_tallies: np.ndarray, algorithm=HISTORY,
                              tally_mode=TALLY_ATOMIC):
        """Drop-in for gpu_transport_photons: `photons` = uint8 array of 120-byte reference Photon records,
        `cell_tallies` = float64 [n_cells][2]; both updated in place."""
        assert photons.dtype == np.uint8 and photons.flags.c_contiguous and photons.size % 120 == 0
        assert cell_tallies.dtype == np.float64 and cell_tallies.size == 2 * self.n_cells
        self._ck(lib().bgpu_transport_photons_aos(self._h, _ptr(photons), photons.size // 120, _ptr(cell_tallies),
                                                  algorithm, tally_mode))


def aos_from_soa(soa: dict, seed: int, source_type=None) -> np.ndarray:
    """Pack photon arrays (the dict layout of Context.download) into the reference's 120-byte AoS `Photon` records
    (src/photon.h:171-182): cell u32, group u32, source_type u32, descriptors u8[4], pos f64[3], angle f64[3], E, E0,
    life_dx, RNG {ctr_lo, seed << 32, stream, 0}.  Returned as a flat uint8 array (what std::vector<Photon>::data() is)."""
    n = len(soa["cell"])
    rec = np.zeros((n, 15), np.uint64)
    rec[:, 0] = soa["cell"].astype(np.uint64) | (soa["group"].astype(np.uint64) << np.uint64(32))
    st = np.full(n, 2, np.uint64) if source_type is None else np.asarray(source_type).astype(np.uint64)
    rec[:, 1] = st | (np.uint64(1) << np.uint64(32))  # descriptor PASS
    rec[:, 2:5] = np.ascontiguousarray(soa["pos"], np.float64).reshape(n, 3).view(np.uint64)
    rec[:, 5:8] = np.ascontiguousarray(soa["angle"], np.float64).reshape(n, 3).view(np.uint64)
    rec[:, 8] = np.ascontiguousarray(soa["E"], np.float64).view(np.uint64)
    rec[:, 9] = np.ascontiguousarray(soa["E0"], np.float64).view(np.uint64)
    rec[:, 10] = np.ascontiguousarray(soa["life_dx"], np.float64).view(np.uint64)
    rec[:, 11] = soa["ctr"]
    rec[:, 12] = np.uint64(seed) << np.uint64(32)
    rec[:, 13] = soa["stream"]
    return rec.view(np.uint8).reshape(-1)


def context_for_deck(deck, nodes, seed=None, n_user_photons=None, rank=0, n_ranks=1, device=-1, **kw) -> Context:
    nx, ny, nz = deck.n_cells_xyz
    xf, yf, zf = faces_from_nodes(nodes, nx, ny, nz)
    return Context(deck.n_groups, nx, ny, nz, xf, yf, zf, deck.bc, deck.seed if seed is None else seed,
                   deck.photons if n_user_photons is None else n_user_photons, rank=rank, n_ranks=n_ranks,
                   device=device, **kw)


def rng_draws(seed, stream, n):
    out = np.zeros(n)
    if lib().bgpu_test_rng_draws(seed, stream, n, _ptr(out)):
        raise GpuError("bgpu_test_rng_draws failed")
    return out


def threefry(ctr, key):
    ck = np.array(list(ctr) + list(key), np.uint64)
    out = np.zeros(2, np.uint64)
    if lib().bgpu_test_threefry(_ptr(ck), _ptr(out)):
        raise GpuError("bgpu_test_threefry failed")
    return int(out[0]), int(out[1])


def fastmath(which, x):
    """csrc/fastmath.cuh on the device: which = "exp" | "log" | "sincos" | "cuda_sincos" (the last: libdevice) | "div"
    (out[i] = x[i] / x[i ^ 1]: neighbouring entries are paired, even length) | "sqrt"."""
    x = np.ascontiguousarray(x, np.float64)
    out, out2 = np.zeros_like(x), np.zeros_like(x)
    code = {"exp": 0, "log": 1, "sincos": 2, "cuda_sincos": 3, "div": 4, "sqrt": 5}[which]
    if lib().bgpu_test_fastmath(code, x.size, _ptr(x), _ptr(out), _ptr(out2)):
        raise GpuError("bgpu_test_fastmath failed")
    return (out, out2) if code in (2, 3) else out
