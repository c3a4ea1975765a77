"""GPU parity of the whole replicated cycle: the C++ host driver (libbranson_host.so: Input -> Mesh -> IMC_State ->
replicated driver) calling the CUDA hot path through the C ABI, against the oracle, cycle after cycle.  Unlike
tests/test_gpu_parity.py nothing is fed from the oracle here: temperatures evolve from the device tallies.

Tolerances (BASELINE.json north_star): integers bit-exact; T_e, T_r, abs_E, track_E, census / exit energies 1e-9
relative; radiation-energy balance 1e-12 of the cycle's energy.
"""
import numpy as np
import pytest

from branson_b200 import decks, driver, gpu
from oracle import port

pytestmark = pytest.mark.gpu


def _rel(got, want, what, rtol=1e-9):
    got, want = np.asarray(got, float), np.asarray(want, float)
    s = np.max(np.abs(want))
    err = np.max(np.abs(got - want)) if got.size else 0.0
    assert err <= rtol * s, f"{what}: max abs err {err:.3e} vs scale {s:.3e}"
    return err / s if s else 0.0


CASES = {
    "three_region_g30": (lambda: decks.simple_three_region(photons=20000, n_groups=30), gpu.TALLY_DETERMINISTIC),
    "marshak": (lambda: decks.marshak_wave(photons=20000, t_stop=0.08), gpu.TALLY_ATOMIC),
    "hot_zone_s10": (lambda: decks.hot_zone(photons=30000, t_stop=0.05, scale=10), gpu.TALLY_ATOMIC),
    "hohlraum_s5_g30": (lambda: decks.hohlraum_single(photons=60000, t_stop=0.04, scale=5), gpu.TALLY_ATOMIC),
    "hohlraum_s5_g30_det": (lambda: decks.hohlraum_single(photons=60000, t_stop=0.03, scale=5), gpu.TALLY_DETERMINISTIC),
    "big_cube_16": (lambda: decks.big_cube(n=16, photons=30000, t_stop=0.004), gpu.TALLY_ATOMIC),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_driver_cycles_match_oracle(name, tmp_path):
    mk, tally_mode = CASES[name]
    deck = mk()
    d = driver.Driver(deck.write(str(tmp_path / "deck.xml")), n_groups=deck.n_groups, device=0, tally_mode=tally_mode,
                      validate=True)
    view = d.gpu_context()
    sim = port.OracleSim(deck)
    cyc = 0
    while not sim.finished():
        cyc += 1
        assert not d.finished()
        sim.cycle(keep_photons=True)
        rep = d.cycle()
        assert rep["step"] == cyc and rep["dt"] == sim.get("dt")[0] and rep["next_dt"] == sim.get("next_dt")[0]
        _rel([rep["global_source_energy"]], sim.get("global_source_energy"), "global_source_energy", 1e-12)
        g = rep["gpu"]
        assert g["n_transported"] == int(sim.get("n_photons")[0]) and g["n_new"] == int(sim.get("n_new")[0])
        assert g["n_census"] == int(sim.get("n_census")[0]) == rep["census_size"]
        post = view.download(gpu.LIST_WORK, counters=True)
        for k in ("cell", "group", "ctr", "descriptor", "counters"):
            assert np.array_equal(post[k], sim.get("post/" + k)), f"cycle {cyc}: post/{k}"
        _rel(d.array("abs_E"), sim.get("abs_E"), "abs_E")
        _rel(d.array("track_E"), sim.get("track_E"), "track_E")
        _rel(d.array("T_e"), sim.get("T_e"), "T_e")
        _rel(d.array("T_r"), sim.get("T_r"), "T_r")
        e = abs(sim.get("global_source_energy")[0]) + abs(sim.get("pre_census_E")[0])
        for k in ("exit_E", "post_census_E", "pre_census_E", "emission_E", "source_E", "absorbed_E", "pre_mat_E",
                  "post_mat_E"):
            assert abs(rep[k] - sim.get(k)[0]) <= 1e-9 * max(e, abs(sim.get(k)[0])), k
        total = rep["pre_census_E"] + rep["emission_E"] + rep["source_E"]
        assert abs(rep["rad_conservation"]) <= 1e-12 * total, (rep["rad_conservation"], total)
        assert abs(rep["rad_balance_exact"]) <= 1e-12 * total, (rep["rad_balance_exact"], total)
        assert abs(rep["mat_conservation"]) <= 1e-12 * max(abs(rep["pre_mat_E"]), abs(rep["post_mat_E"]))
    assert d.finished() and cyc == deck.n_cycles()


def test_command_line_binary_runs_a_deck(tmp_path):
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(gpu.LIB_PATH), "bin", "branson")
    deck = decks.hohlraum_single(photons=30000, t_stop=0.02, scale=5)
    out = subprocess.run([exe, deck.write(str(tmp_path / "d.xml")), "30"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Photons Per Second (FOM)" in out.stdout and out.stdout.count("Radiation conservation") == 2
    sim = port.OracleSim(deck)
    sim.cycle(keep_photons=False)
    assert f"Total Photons transported: {int(sim.get('n_photons')[0])}" in out.stdout


@pytest.mark.parametrize("mesh_on_device", [False, True])
def test_driver_with_census_comb_bounds_the_census_and_conserves_energy(tmp_path, mesh_on_device):
    """Driver option comb_max_census (population control, bgpu_comb_census): on the all-reflecting cube nothing ever
    leaves and the census settles near 40 % of the photon budget; with the comb it stays near the (three times smaller)
    target, every cycle's radiation balance still closes to 1e-12 (the comb conserves each cell's census energy),
    and the temperatures stay within Monte Carlo noise of the uncombed run."""
    deck = decks.big_cube(n=16, photons=60000, t_stop=0.008)
    xml = deck.write(str(tmp_path / "cube.xml"))
    runs = {}
    for target in (0, 8000):
        d = driver.Driver(xml, n_groups=1, device=0, mesh_on_device=mesh_on_device, comb_max_census=target)
        reps = []
        while not d.finished():
            reps.append(d.cycle())
        runs[target] = (reps, d.array("T_e"))
        d.close()
    plain, combed = runs[0][0], runs[8000][0]
    assert len(plain) == len(combed) == 8
    assert all(r["comb_n_before"] == 0 for r in plain)
    assert plain[-1]["census_size"] > 2 * 8000
    assert sum(1 for r in combed if r["comb_n_before"] > 8000) >= 4
    for r in combed:
        total = r["pre_census_E"] + r["emission_E"] + r["source_E"]
        assert abs(r["rad_balance_exact"]) <= 1e-12 * total
        if r["comb_n_before"]:
            assert r["comb_n_after"] < r["comb_n_before"]
            assert 0.8 * 8000 <= r["comb_n_after"] <= 1.2 * 8000 + 16 ** 3   # ~target (+ at most one per cell)
    # the census energy a cycle starts from is the one the previous cycle ended with, combed or not
    for a, b in zip(combed[:-1], combed[1:]):
        assert abs(b["pre_census_E"] - a["post_census_E"]) <= 1e-12 * a["post_census_E"]
    T0, T1 = runs[0][1], runs[8000][1]
    assert abs(T1.mean() - T0.mean()) <= 2e-3 * T0.mean()
