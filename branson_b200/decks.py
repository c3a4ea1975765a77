"""Problem decks for the IMC hot path (the five BASELINE.json configs and scaled variants).

The reference reads XML decks (schema: reference src/input.h:105-449).  The decks
here are *generated* -- nothing is copied from the reference's inputs/ directory --
from compact Python descriptions whose numbers follow
  inputs/marshak_wave_replicated.xml, inputs/hot_zone_input.xml,
  inputs/3D_hohlraum_single_node.xml, inputs/3D_hohlraum_multi_node.xml and
  inputs/big_cube.xml
of the reference.  `Deck.to_xml()` writes a file both the reference binary
(oracle/_ref) and our own host-side `Input` parser accept, so every consumer sees
exactly the same problem.  Doubles are written with repr() (shortest round-trip),
so strtod recovers the identical bits.
"""
from __future__ import annotations

import copy
import os
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

# Constants::bc_type order X_NEG..Z_POS == bc_left, bc_right, bc_down, bc_up, bc_bottom, bc_top
# (reference src/input.h:345-415)
BC_TAGS = ("bc_left", "bc_right", "bc_down", "bc_up", "bc_bottom", "bc_top")
BC_CODES = {"REFLECT": 0, "VACUUM": 1, "ELEMENT": 2, "SOURCE": 3}


@dataclass
class Region:
    ID: int
    density: float
    CV: float
    opacA: float
    opacB: float
    opacC: float
    opacS: float
    initial_T_e: float
    initial_T_r: float


@dataclass
class Deck:
    name: str
    t_start: float
    t_stop: float
    dt_start: float
    t_mult: float
    dt_max: float
    photons: int
    seed: int
    x_div: List[Tuple[float, float, int]]
    y_div: List[Tuple[float, float, int]]
    z_div: List[Tuple[float, float, int]]
    region_map: Dict[Tuple[int, int, int], int]  # (x_div, y_div, z_div) -> region ID
    bc: Tuple[str, str, str, str, str, str]      # X_NEG X_POS Y_NEG Y_POS Z_NEG Z_POS
    regions: List[Region]
    n_groups: int = 1
    T_source: float = 0.0
    dd_transport_type: str = "REPLICATED"
    particle_storage: str = "AOS"
    particle_algorithm: str = "HISTORY"
    use_gpu_transporter: str = "FALSE"
    n_omp_threads: int = 1
    batch_size: int | None = None
    print_verbose: bool = False
    extra: Dict[str, str] = field(default_factory=dict)

    # ---- derived ---------------------------------------------------------
    @property
    def n_cells_xyz(self) -> Tuple[int, int, int]:
        return (sum(d[2] for d in self.x_div), sum(d[2] for d in self.y_div), sum(d[2] for d in self.z_div))

    @property
    def n_cells(self) -> int:
        nx, ny, nz = self.n_cells_xyz
        return nx * ny * nz

    def n_cycles(self) -> int:
        """Number of cycles the reference time stepping produces (src/imc_state.h:112-134,292-296)."""
        t, dt, n = self.t_start, self.dt_start, 0
        while not abs(t - self.t_stop) < 1.0e-8:
            t += dt
            nxt = dt * self.t_mult if dt * self.t_mult < self.dt_max else self.dt_max
            if t + nxt > self.t_stop:
                nxt = self.t_stop - t
            dt = nxt
            n += 1
            if n > 10_000_000:
                raise RuntimeError("time stepping does not terminate")
        return n

    def with_(self, **kw) -> "Deck":
        d = copy.deepcopy(self)
        for k, v in kw.items():
            if not hasattr(d, k):
                raise AttributeError(k)
            setattr(d, k, v)
        return d

    # ---- XML -------------------------------------------------------------
    def to_xml(self) -> str:
        r = repr
        o: List[str] = ["<prototype>", "<common>"]

        def tag(name, val, ind="  "):
            o.append(f"{ind}<{name}>{val}</{name}>")

        tag("t_start", r(float(self.t_start)))
        tag("t_stop", r(float(self.t_stop)))
        tag("dt_start", r(float(self.dt_start)))
        tag("t_mult", r(float(self.t_mult)))
        tag("dt_max", r(float(self.dt_max)))
        tag("photons", int(self.photons))
        tag("seed", int(self.seed))
        tag("use_combing", "FALSE")
        tag("use_gpu_transporter", self.use_gpu_transporter)
        tag("dd_transport_type", self.dd_transport_type)
        tag("particle_storage", self.particle_storage)
        tag("particle_algorithm", self.particle_algorithm)
        tag("n_omp_threads", int(self.n_omp_threads))
        if self.batch_size is not None:
            tag("batch_size", int(self.batch_size))
        tag("output_frequency", 1)
        tag("write_silo", "FALSE")
        for k, v in self.extra.items():
            tag(k, v)
        o.append("</common>")
        o.append("<debug_options>")
        tag("print_verbose", "TRUE" if self.print_verbose else "FALSE")
        tag("print_mesh_info", "FALSE")
        o.append("</debug_options>")
        o.append("<spatial>")
        for ax, divs in (("x", self.x_div), ("y", self.y_div), ("z", self.z_div)):
            for (a, b, n) in divs:
                o.append(f"  <{ax}_division><{ax}_start>{r(float(a))}</{ax}_start><{ax}_end>{r(float(b))}</{ax}_end>"
                         f"<n_{ax}_cells>{int(n)}</n_{ax}_cells></{ax}_division>")
        for (ix, iy, iz), rid in sorted(self.region_map.items(), key=lambda kv: (kv[0][2], kv[0][1], kv[0][0])):
            o.append(f"  <region_map><x_div_ID>{ix}</x_div_ID><y_div_ID>{iy}</y_div_ID><z_div_ID>{iz}</z_div_ID>"
                     f"<region_ID>{rid}</region_ID></region_map>")
        o.append("</spatial>")
        o.append("<boundary>")
        for t, v in zip(BC_TAGS, self.bc):
            tag(t, v)
        if "SOURCE" in self.bc:
            tag("T_source", r(float(self.T_source)))
        o.append("</boundary>")
        o.append("<regions>")
        for g in self.regions:
            o.append(f"  <region><ID>{g.ID}</ID><density>{r(float(g.density))}</density><CV>{r(float(g.CV))}</CV>"
                     f"<opacA>{r(float(g.opacA))}</opacA><opacB>{r(float(g.opacB))}</opacB>"
                     f"<opacC>{r(float(g.opacC))}</opacC><opacS>{r(float(g.opacS))}</opacS>"
                     f"<initial_T_e>{r(float(g.initial_T_e))}</initial_T_e>"
                     f"<initial_T_r>{r(float(g.initial_T_r))}</initial_T_r></region>")
        o.append("</regions>")
        o.append("</prototype>")
        return "\n".join(o) + "\n"

    def write(self, path: str) -> str:
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "w") as fh:
            fh.write(self.to_xml())
        return path


# --------------------------------------------------------------------------
# The five BASELINE.json configurations
# --------------------------------------------------------------------------
def marshak_wave(photons: int = 1_000_000, t_stop: float = 1.0, n_x: int = 25) -> Deck:
    """Gray 1-D-like Marshak wave with a T=1 source on -x (reference inputs/marshak_wave_replicated.xml)."""
    return Deck(
        name="marshak_wave_replicated", t_start=0.0, t_stop=t_stop, dt_start=0.01, t_mult=1.0, dt_max=1.0,
        photons=photons, seed=14706,
        x_div=[(0.0, 0.25, n_x)], y_div=[(0.0, 0.1, 1)], z_div=[(0.0, 0.1, 1)],
        region_map={(0, 0, 0): 1},
        bc=("SOURCE", "VACUUM", "REFLECT", "REFLECT", "REFLECT", "REFLECT"), T_source=1.0,
        regions=[Region(1, 1.0, 1.0, 0.0, 100.0, -3.0, 0.0, 0.01, 0.01)],
        n_groups=1)


def hot_zone(photons: int = 1_000_000, t_stop: float = 0.2, scale: int = 1) -> Deck:
    """Two-region gray problem, hot 5x5 corner in a cold 200x200 sheet (reference inputs/hot_zone_input.xml).
    scale>1 divides the cell counts (5/195 -> 5/s, 195/s) for small test problems."""
    a, b = max(1, 5 // scale), max(1, 195 // scale)
    return Deck(
        name="hot_zone", t_start=0.0, t_stop=t_stop, dt_start=0.01, t_mult=1.0, dt_max=1.0,
        photons=photons, seed=14706,
        x_div=[(0.0, 0.05, a), (0.05, 2.0, b)], y_div=[(0.0, 0.05, a), (0.05, 2.0, b)], z_div=[(0.0, 1.0, 1)],
        region_map={(0, 0, 0): 100, (1, 0, 0): 5, (0, 1, 0): 5, (1, 1, 0): 5},
        bc=("REFLECT",) * 6,
        regions=[Region(100, 1.0, 0.1, 50.0, 0.0, 0.0, 0.0, 1.0, 1.0),
                 Region(5, 1.0, 0.1, 50.0, 0.0, 0.0, 0.0, 0.01, 0.01)],
        n_groups=1,
        # the reference deck names an unknown dd type, which falls back to PARTICLE_PASS and, at one
        # rank, to REPLICATED (src/input.h:176-191); we state REPLICATED directly.
        dd_transport_type="REPLICATED")


_HOHL_LAYERS = [
    # z-division -> 3x3 (y rows, x columns) region IDs (reference inputs/3D_hohlraum_single_node.xml:103-440)
    [[1001, 1001, 1001], [1001, 1001, 1001], [1001, 1001, 1001]],
    [[7777, 1000, 7777], [1000, 1000, 7777], [7777, 7777, 7777]],
    [[1000, 1000, 7777], [1000, 1000, 7777], [7777, 7777, 7777]],
    [[7777, 1000, 7777], [1000, 1000, 7777], [7777, 7777, 7777]],
    [[1000, 1000, 7777], [1000, 1000, 7777], [7777, 7777, 7777]],
    [[7777, 7777, 7777], [7777, 7777, 7777], [7777, 7777, 7777]],
]


def _hohlraum(name: str, photons: int, t_stop: float, dt: float, T_hot: float, T_cold: float, scale: int) -> Deck:
    def n(c):
        return max(1, c // scale)
    xy = [(0.0, 0.4, n(40)), (0.4, 0.6, n(20)), (0.60, .65, n(5))]
    z = [(0.0, 0.1, n(10)), (0.1, 0.15, n(5)), (0.15, 0.55, n(40)), (0.55, 0.95, n(40)), (0.95, 1.35, n(40)),
         (1.35, 1.4, n(5))]
    rmap = {}
    for iz, layer in enumerate(_HOHL_LAYERS):
        for iy, row in enumerate(layer):
            for ix, rid in enumerate(row):
                rmap[(ix, iy, iz)] = rid
    return Deck(
        name=name, t_start=0.0, t_stop=t_stop, dt_start=dt, t_mult=1.0, dt_max=1.0, photons=photons, seed=14706,
        x_div=list(xy), y_div=list(xy), z_div=z, region_map=rmap,
        bc=("REFLECT", "VACUUM", "REFLECT", "VACUUM", "VACUUM", "VACUUM"),
        regions=[Region(1001, 1.0, 0.3, 0.001, 0.0, 0.0, 0.0, T_hot, T_hot),
                 Region(1000, 1.0, 0.3, 0.001, 0.0, 0.0, 0.0, T_cold, T_cold),
                 Region(7777, 1.0, 0.3, 10000.0, 0.0, 0.0, 0.0, T_cold, T_cold)],
        n_groups=30, use_gpu_transporter="TRUE")


def hohlraum_single(photons: int = 10_000_000, t_stop: float = 0.05, scale: int = 1) -> Deck:
    """3-D hohlraum, 65x65x140 cells, 30 groups, 5 cycles (reference inputs/3D_hohlraum_single_node.xml)."""
    return _hohlraum("3D_hohlraum_single_node", photons, t_stop, 0.01, 1.0, 0.001, scale)


def hohlraum_multi(photons: int = 250_000_000, t_stop: float = 0.020, scale: int = 1) -> Deck:
    """Multi-node hohlraum deck (reference inputs/3D_hohlraum_multi_node.xml) with dd_transport_type forced to
    REPLICATED: PARTICLE_PASS exits in this reference snapshot (src/particle_pass_transport.h:141-142)."""
    return _hohlraum("3D_hohlraum_multi_node", photons, t_stop, 0.001, 0.1, 0.1, scale)


def big_cube(n: int = 200, photons: int = 1_000_000_000, t_stop: float = 0.01) -> Deck:
    """All-reflecting cube, sigma_a=100, T=1, dx=0.005 (reference inputs/big_cube.xml is 800^3 = 98 GB in the
    reference layout; BASELINE.json scales it, default here 200^3)."""
    side = 0.005 * n
    return Deck(
        name=f"big_cube_{n}", t_start=0.0, t_stop=t_stop, dt_start=0.001, t_mult=1.0, dt_max=1.0,
        photons=photons, seed=14706,
        x_div=[(0.0, side, n)], y_div=[(0.0, side, n)], z_div=[(0.0, side, n)],
        region_map={(0, 0, 0): 6}, bc=("REFLECT",) * 6,
        regions=[Region(6, 1.0, 1.0, 100.0, 0.0, 0.0, 0.0, 1.0, 1.0)], n_groups=1)


def simple_three_region(photons: int = 20_000, n_groups: int = 1) -> Deck:
    """Small mixed problem for tests: scattering (opacS>0), temperature-dependent opacity, a SOURCE face on +y,
    vacuum and reflecting faces, three regions, uneven divisions."""
    return Deck(
        name="three_region", t_start=0.0, t_stop=0.03, dt_start=0.01, t_mult=1.5, dt_max=0.02,
        photons=photons, seed=777,
        x_div=[(0.0, 0.3, 3), (0.3, 1.0, 5)], y_div=[(0.0, 0.5, 4)], z_div=[(0.0, 0.2, 2), (0.2, 0.7, 3)],
        region_map={(0, 0, 0): 10, (1, 0, 0): 11, (0, 0, 1): 11, (1, 0, 1): 12},
        bc=("REFLECT", "VACUUM", "REFLECT", "SOURCE", "VACUUM", "REFLECT"), T_source=0.8,
        regions=[Region(10, 1.0, 0.2, 5.0, 1.0, -1.0, 2.0, 0.5, 0.4),
                 Region(11, 2.0, 0.1, 0.5, 0.0, 0.0, 0.1, 0.1, 0.1),
                 Region(12, 0.5, 0.3, 100.0, 0.0, 0.0, 0.0, 0.05, 0.05)],
        n_groups=n_groups)


BASELINE_DECKS = {
    "marshak_wave_replicated": marshak_wave,
    "hot_zone": hot_zone,
    "3D_hohlraum_single_node": hohlraum_single,
    "3D_hohlraum_multi_node": hohlraum_multi,
    "big_cube": big_cube,
}
