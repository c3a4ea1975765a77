"""ctypes front end of oracle/liboracle.so (the plain-C restatement, oracle/imc_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "imc_oracle.c")
    hdr = os.path.join(_HERE, "imc_oracle.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return so


class _Problem(C.Structure):
    _fields_ = [
        ("t_start", C.c_double), ("t_stop", C.c_double), ("dt_start", C.c_double), ("t_mult", C.c_double),
        ("dt_max", C.c_double), ("n_photons", C.c_uint64), ("seed", C.c_uint32), ("n_groups", C.c_uint32),
        ("n_xdiv", C.c_int32), ("n_ydiv", C.c_int32), ("n_zdiv", C.c_int32),
        ("x_start", C.c_void_p), ("x_end", C.c_void_p), ("x_cells", C.c_void_p),
        ("y_start", C.c_void_p), ("y_end", C.c_void_p), ("y_cells", C.c_void_p),
        ("z_start", C.c_void_p), ("z_end", C.c_void_p), ("z_cells", C.c_void_p),
        ("div_region", C.c_void_p), ("bc", C.c_int32 * 6), ("T_source", C.c_double),
        ("n_regions", C.c_int32), ("region_id", C.c_void_p), ("region_props", C.c_void_p),
        ("n_ranks", C.c_int32),
    ]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(_Problem)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_finished.argtypes = [C.c_void_p]
        L.orc_cycle.argtypes = [C.c_void_p, C.c_int]
        L.orc_get.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                              C.POINTER(C.c_int)]
        L.orc_last_transport_seconds.restype = C.c_double
        L.orc_last_transport_seconds.argtypes = [C.c_void_p]
        L.orc_rng_next.restype = C.c_double
        L.orc_rng_next.argtypes = [C.c_void_p]
        L.orc_rng_init.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64]
        L.orc_threefry2x64_20.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_uniform_angle.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_distance_to_boundary.restype = C.c_double
        L.orc_distance_to_boundary.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_transport_list.argtypes = [C.c_void_p, C.c_int, C.c_uint64] + [C.c_void_p] * 13
        _LIB = L
    return _LIB


_DT = {0: np.float64, 1: np.uint32, 2: np.uint64, 3: np.uint8}
_BC = {"REFLECT": 0, "VACUUM": 1, "ELEMENT": 2, "SOURCE": 3}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleSim:
    """One replicated-mode simulation (all emulated ranks) of a branson_b200.decks.Deck."""

    def __init__(self, deck, n_ranks: int = 1):
        self.deck = deck
        self.n_ranks = n_ranks
        keep = self._keep = []

        def arr(x, dt):
            a = np.ascontiguousarray(np.array(x, dtype=dt))
            keep.append(a)
            return a

        p = _Problem()
        p.t_start, p.t_stop, p.dt_start, p.t_mult, p.dt_max = (deck.t_start, deck.t_stop, deck.dt_start, deck.t_mult,
                                                              deck.dt_max)
        p.n_photons, p.seed, p.n_groups = deck.photons, deck.seed, deck.n_groups
        p.n_xdiv, p.n_ydiv, p.n_zdiv = len(deck.x_div), len(deck.y_div), len(deck.z_div)
        for ax, divs in (("x", deck.x_div), ("y", deck.y_div), ("z", deck.z_div)):
            setattr(p, f"{ax}_start", _ptr(arr([d[0] for d in divs], np.float64)))
            setattr(p, f"{ax}_end", _ptr(arr([d[1] for d in divs], np.float64)))
            setattr(p, f"{ax}_cells", _ptr(arr([d[2] for d in divs], np.uint32)))
        dr = np.zeros((p.n_zdiv, p.n_ydiv, p.n_xdiv), dtype=np.uint32)
        for (ix, iy, iz), rid in deck.region_map.items():
            dr[iz, iy, ix] = rid
        p.div_region = _ptr(arr(dr, np.uint32))
        for i, b in enumerate(deck.bc):
            p.bc[i] = _BC[b]
        p.T_source = deck.T_source
        p.n_regions = len(deck.regions)
        p.region_id = _ptr(arr([r.ID for r in deck.regions], np.uint32))
        p.region_props = _ptr(arr([[r.density, r.CV, r.opacA, r.opacB, r.opacC, r.opacS, r.initial_T_e,
                                    r.initial_T_r] for r in deck.regions], np.float64))
        p.n_ranks = n_ranks
        self._h = lib().orc_create(C.byref(p))
        self.cycle_index = 0

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def finished(self) -> bool:
        return bool(lib().orc_finished(self._h))

    def cycle(self, keep_photons: bool = True):
        lib().orc_cycle(self._h, 1 if keep_photons else 0)
        self.cycle_index += 1

    def transport_seconds(self) -> float:
        return lib().orc_last_transport_seconds(self._h)

    def get(self, name: str, rank: int = 0) -> np.ndarray:
        data, count, dt = C.c_void_p(), C.c_uint64(), C.c_int()
        if lib().orc_get(self._h, rank, name.encode(), C.byref(data), C.byref(count), C.byref(dt)):
            raise KeyError(name)
        n = count.value
        if n == 0:
            return np.zeros(0, dtype=_DT[dt.value])
        buf = (C.c_char * (n * np.dtype(_DT[dt.value]).itemsize)).from_address(data.value)
        return np.frombuffer(buf, dtype=_DT[dt.value]).copy()

    def transport_list(self, cell, group, pos, angle, E, E0, life_dx, ctr, stream, rank: int = 0):
        """Transport caller-supplied photons on the current mesh state; returns dict of outputs."""
        n = len(cell)
        nc = self.deck.n_cells
        out = dict(cell=np.array(cell, np.uint32), group=np.array(group, np.uint32),
                   pos=np.array(pos, np.float64).reshape(n, 3).copy(), angle=np.array(angle, np.float64).reshape(n, 3).copy(),
                   E=np.array(E, np.float64), life_dx=np.array(life_dx, np.float64), ctr=np.array(ctr, np.uint64),
                   descriptor=np.zeros(n, np.uint8), abs_E=np.zeros(nc), track_E=np.zeros(nc),
                   counters=np.zeros((n, 4), np.uint32))
        E0 = np.ascontiguousarray(E0, np.float64)
        stream = np.ascontiguousarray(stream, np.uint64)
        lib().orc_transport_list(self._h, rank, n, _ptr(out["cell"]), _ptr(out["group"]), _ptr(out["pos"]),
                                 _ptr(out["angle"]), _ptr(out["E"]), _ptr(E0), _ptr(out["life_dx"]), _ptr(out["ctr"]),
                                 _ptr(stream), _ptr(out["descriptor"]), _ptr(out["abs_E"]), _ptr(out["track_E"]),
                                 _ptr(out["counters"]))
        return out


def rng_draws(seed: int, stream: int, n: int) -> np.ndarray:
    st = np.zeros(4, np.uint64)
    lib().orc_rng_init(_ptr(st), seed, stream)
    return np.array([lib().orc_rng_next(_ptr(st)) for _ in range(n)])


def threefry(ctr, key):
    c = np.array(ctr, np.uint64)
    k = np.array(key, np.uint64)
    o = np.zeros(2, np.uint64)
    lib().orc_threefry2x64_20(_ptr(c), _ptr(k), _ptr(o))
    return int(o[0]), int(o[1])


def uniform_angle(seed: int, stream: int):
    st = np.zeros(4, np.uint64)
    lib().orc_rng_init(_ptr(st), seed, stream)
    a = np.zeros(3)
    lib().orc_uniform_angle(_ptr(st), _ptr(a))
    return a


def comb_photons(cell, E, global_census_E: float, max_census_photons: int, seed: int, stream: int):
    """reference census_functions.h:48-93 on a census given as arrays; returns (keep mask, corrected energies of the
    kept photons, RNG draws consumed)."""
    cell = np.ascontiguousarray(cell, np.uint32)
    E = np.ascontiguousarray(E, np.float64)
    n = len(cell)
    st = np.zeros(4, np.uint64)
    lib().orc_rng_init(_ptr(st), seed, stream)
    keep, new_E = np.zeros(n, np.uint8), np.zeros(n)
    L = lib()
    L.orc_comb_photons.restype = C.c_uint64
    L.orc_comb_photons.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_double, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p]
    kept = L.orc_comb_photons(n, _ptr(cell), _ptr(E), float(global_census_E), int(max_census_photons), _ptr(st),
                              _ptr(keep), _ptr(new_E))
    assert kept == int(keep.sum())
    return keep.astype(bool), new_E[keep.astype(bool)], int(st[0])
