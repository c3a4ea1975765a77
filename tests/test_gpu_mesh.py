"""Mesh physics on the device (csrc/mesh_dev.cuh, bgpu_mesh_*) against the host Mesh (csrc/host/mesh.h, which is
bit-identical to the reference's host code -- tests/test_host_logic.py): the same run stepped both ways.

Only pow() and the order of the running sums differ (last bits), so the per-cell doubles agree to 1e-13 while the
integer outcomes -- photon counts per cell, per-photon cells / groups / RNG counters / event counters -- are equal."""
import numpy as np
import pytest

from branson_b200 import decks, driver, gpu

pytestmark = pytest.mark.gpu

CASES = [
    ("marshak", lambda: decks.marshak_wave(photons=20000, t_stop=0.06), 1),      # T-dependent opacity, SOURCE face
    ("three_region", lambda: decks.simple_three_region(photons=20000, n_groups=30), 30),
    ("hot_zone", lambda: decks.hot_zone(photons=20000, t_stop=0.04, scale=10), 1),
    ("hohlraum", lambda: decks.hohlraum_single(photons=30000, t_stop=0.03, scale=5), 30),
]


def _step(deck, n_groups, tmp_path, on_device):
    d = driver.Driver(deck.write(str(tmp_path / f"{deck.name}_{int(on_device)}.xml")), n_groups=n_groups, device=0,
                      validate=True, mesh_on_device=on_device)
    view = d.gpu_context()
    out = []
    while not d.finished():
        rep = d.cycle()
        post = view.download(gpu.LIST_WORK, counters=True)
        out.append((rep, {k: d.array(k) for k in ("T_e", "T_r", "f", "op_a", "E_emission", "E_source", "abs_E",
                                                   "track_E")}, post))
    d.close()
    return out


@pytest.mark.parametrize("name,make,n_groups", CASES, ids=[c[0] for c in CASES])
def test_device_mesh_matches_host_mesh(name, make, n_groups, tmp_path):
    deck = make()
    host = _step(deck, n_groups, tmp_path, False)
    dev = _step(deck, n_groups, tmp_path, True)
    assert len(host) == len(dev) >= 3
    for (rh, ah, ph), (rd, ad, pd) in zip(host, dev):
        # integers: the same photons do the same things
        for k in ("n_new", "n_transported", "n_census", "n_killed", "n_exit", "n_events", "n_scatters", "n_crossings",
                  "n_reflections"):
            assert rh["gpu"][k] == rd["gpu"][k], (rh["step"], k)
        for k in ("cell", "group", "ctr", "descriptor", "counters"):
            assert np.array_equal(ph[k], pd[k]), (rh["step"], k)
        # per-cell doubles
        for k, a in ah.items():
            scale = np.max(np.abs(a)) or 1.0
            assert np.max(np.abs(a - ad[k])) <= 1e-13 * scale, (rh["step"], k)
        # the cycle's sums and balances
        for k in ("emission_E", "source_E", "absorbed_E", "pre_mat_E", "post_mat_E", "pre_census_E", "post_census_E",
                  "exit_E", "global_source_energy"):
            assert abs(rh[k] - rd[k]) <= 1e-12 * max(abs(rh[k]), 1e-300), (rh["step"], k, rh[k], rd[k])
        total = rd["pre_census_E"] + rd["emission_E"] + rd["source_E"]
        assert abs(rd["rad_balance_exact"]) <= 1e-12 * total
        assert abs(rd["rad_conservation"]) <= 1e-12 * total


def test_device_mesh_against_oracle(tmp_path):
    """the device-mesh driver end to end against the oracle (the same check tests/test_gpu_driver.py runs on the host mesh)"""
    from oracle import port
    deck = decks.marshak_wave(photons=20000, t_stop=0.05)
    sim = port.OracleSim(deck)
    d = driver.Driver(deck.write(str(tmp_path / "m.xml")), n_groups=1, device=0, validate=True, mesh_on_device=True)
    view = d.gpu_context()
    while not sim.finished():
        sim.cycle(keep_photons=True)
        d.cycle()
        post = view.download(gpu.LIST_WORK, counters=True)
        for k in ("cell", "group", "ctr", "descriptor", "counters"):
            assert np.array_equal(post[k], sim.get("post/" + k)), k
        for k in ("T_e", "T_r"):
            want = sim.get(k)
            assert np.max(np.abs(d.array(k) - want)) <= 1e-9 * np.max(np.abs(want)), k
    assert d.finished()
    d.close()
