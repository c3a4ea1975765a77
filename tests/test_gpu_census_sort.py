"""Census ordered by cell between cycles (bgpu_sort_census_by_cell).  The reference keeps the
census in the order post_process_photons appended it (src/post_process_functions.h:33-59); per-photon results cannot
depend on the order, so the checks are: the device sort IS the stable sort by cell of the list
it was given, bit for bit; and a run that sorts after every cycle transports exactly the same photons to exactly the
same fates as one that does not, with tallies equal to summation-order rounding."""
import numpy as np
import pytest

from branson_b200 import decks, driver, gpu
from oracle import port

pytestmark = pytest.mark.gpu

FIELDS = ("cell", "group", "ctr", "stream", "E", "E0", "life_dx")


def test_sort_is_the_stable_sort_by_cell_of_the_census():
    deck = decks.hot_zone(photons=200_000, t_stop=0.02, scale=4)
    sim = port.OracleSim(deck)
    ctx = None
    for cyc in (1, 2):
        sim.cycle(keep_photons=False)
        if ctx is None:
            ctx = gpu.context_for_deck(deck, sim.get("mesh/nodes"), device=0)
        ctx.set_cell_data(sim.get("f"), sim.get("op_a"), sim.get("op_s"))
        ctx.source(cyc, sim.get("dt")[0], sim.get("E_emission"), sim.get("E_source"),
                   sim.get("E_census") if cyc == 1 else None, sim.get("global_source_energy")[0])
        ctx.transport(sim.get("next_dt")[0])
        ctx.tallies()
    pre = ctx.download(gpu.LIST_CENSUS)
    n = len(pre["cell"])
    assert n > 10_000 and np.any(np.diff(pre["cell"].astype(np.int64)) < 0)  # not sorted to begin with
    ctx.sort_census_by_cell()
    post = ctx.download(gpu.LIST_CENSUS)
    order = np.argsort(pre["cell"], kind="stable")
    for k in FIELDS:
        assert np.array_equal(post[k].view(np.uint64) if post[k].dtype == np.float64 else post[k],
                              pre[k][order].view(np.uint64) if pre[k].dtype == np.float64 else pre[k][order]), k
    assert np.array_equal(post["pos"].reshape(-1, 3), pre["pos"].reshape(-1, 3)[order])
    assert np.array_equal(post["angle"].reshape(-1, 3), pre["angle"].reshape(-1, 3)[order])
    # idempotent, and the sorted census feeds the next cycle like any other
    ctx.sort_census_by_cell()
    again = ctx.download(gpu.LIST_CENSUS)
    assert np.array_equal(again["stream"], post["stream"])
    sim.cycle(keep_photons=False)
    ctx.set_cell_data(sim.get("f"), sim.get("op_a"), sim.get("op_s"))
    n_new, n_tot = ctx.source(3, sim.get("dt")[0], sim.get("E_emission"), sim.get("E_source"), None,
                              sim.get("global_source_energy")[0])
    assert n_tot == n_new + n
    ctx.transport(sim.get("next_dt")[0])
    a, t, st = ctx.tallies()
    # same photons as the oracle's third cycle (which kept the reference's order): same integer bookkeeping, same tallies
    assert st["n_transported"] == int(sim.get("n_photons")[0])
    assert st["n_census"] == int(sim.get("n_census")[0])
    want = sim.get("rank_abs_E")
    assert np.max(np.abs(a - want)) <= 1e-9 * want.max()
    ctx.close()


@pytest.mark.parametrize("mesh_on_device", [False, True])
def test_driver_option_sort_census_changes_no_photon(tmp_path, mesh_on_device):
    deck = decks.big_cube(n=16, photons=60000, t_stop=0.005)
    xml = deck.write(str(tmp_path / "cube.xml"))
    runs = {}
    for flag in (False, True):
        d = driver.Driver(xml, n_groups=1, device=0, mesh_on_device=mesh_on_device, sort_census=flag)
        reps = []
        while not d.finished():
            reps.append(d.cycle())
        cen = d.gpu_context().download(gpu.LIST_CENSUS)
        runs[flag] = (reps, d.array("T_e"), cen)
        d.close()
    plain, srt = runs[False], runs[True]
    assert len(plain[0]) == len(srt[0]) == 5
    for a, b in zip(plain[0], srt[0]):
        for k in ("n_transported", "n_census", "n_killed", "n_exit", "n_events", "n_scatters", "n_crossings", "n_reflections"):
            assert a["gpu"][k] == b["gpu"][k], k
        total = b["pre_census_E"] + b["emission_E"] + b["source_E"]
        assert abs(b["rad_balance_exact"]) <= 1e-12 * total
    assert np.max(np.abs(plain[1] - srt[1])) <= 1e-9 * plain[1].max()
    # the same census photons, in cell order: identify photons by their stream number (unique per photon and cycle)
    c0, c1 = plain[2], srt[2]
    assert np.all(np.diff(c1["cell"].astype(np.int64)) >= 0)
    o0, o1 = np.argsort(c0["stream"], kind="stable"), np.argsort(c1["stream"], kind="stable")
    for k in ("stream", "cell", "group", "ctr"):
        assert np.array_equal(c0[k][o0], c1[k][o1]), k
    assert np.max(np.abs(c0["E"][o0] - c1["E"][o1])) <= 1e-9 * c0["E"].max()
