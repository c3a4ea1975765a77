"""Device comb (bgpu_comb_census, csrc/comb.cuh) against the reference's comb_photons (src/census_functions.h:48-93):
the golden fixtures were produced by the UNMODIFIED reference function on a real census (oracle/ref_harness.cc,
oracle/gen_golden.py), the live comparisons use the plain-C restatement pinned to them (tests/test_oracle_golden.py).
Bit-exact: which photons survive, their order, their corrected energies, everything else they carry."""
import os

import numpy as np
import pytest

from branson_b200 import decks, gpu
from oracle import port
from oracle.gen_golden import comb_cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _ctx_for(deck, nodes):
    return gpu.context_for_deck(deck, nodes, device=0)


def _upload_census(ctx, cell, group, pos, angle, E, E0, life_dx, ctr, stream):
    ctx.upload(gpu.LIST_CENSUS, cell, group, pos, angle, E, E0, life_dx, ctr, stream)


@pytest.mark.parametrize("name", sorted(comb_cases()))
def test_comb_matches_reference_fixture(name):
    deck, cycles, max_census, stream = comb_cases()[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sim = port.OracleSim(deck)  # only for the mesh geometry
    sim.cycle(keep_photons=False)
    ctx = _ctx_for(deck, sim.get("mesh/nodes"))
    n = len(g["comb/pre/cell"])
    _upload_census(ctx, g["comb/pre/cell"], g["comb/pre/group"], g["comb/pre/pos"], g["comb/pre/angle"], g["comb/pre/E"],
                   g["comb/pre/E0"], g["comb/pre/life_dx"], g["comb/pre/ctr"], g["comb/pre/stream"])
    st = ctx.comb_census(max_census, g["comb/local_census_E"][0], stream)
    assert st["n_before"] == n and st["rng_draws"] == int(g["comb/rng_draws"][0])
    assert st["n_after"] == len(g["comb/post/cell"])
    post = ctx.download(gpu.LIST_CENSUS)
    assert np.array_equal(post["stream"], g["comb/post/stream"])
    assert np.array_equal(post["cell"], g["comb/post/cell"])
    assert np.array_equal(post["group"], g["comb/post/group"])
    assert np.array_equal(post["ctr"], g["comb/post/ctr"])
    assert np.array_equal(post["E"].view(np.uint64), g["comb/post/E"].view(np.uint64))
    assert np.array_equal(post["pos"].view(np.uint64), g["comb/post/pos"].view(np.uint64))
    assert np.array_equal(post["angle"].view(np.uint64), g["comb/post/angle"].view(np.uint64))
    assert np.array_equal(post["life_dx"].view(np.uint64), g["comb/post/life_dx"].view(np.uint64))
    assert abs(st["E_after"] - st["E_before"]) <= 1e-13 * st["E_before"]
    ctx.close()


@pytest.mark.parametrize("max_census", [1, 500, 10 ** 7])
def test_comb_of_a_device_census_matches_oracle(max_census):
    """The census the device itself leaves after two cycles of a hot_zone run (3e5 photons), combed on the device and
    by the oracle: extreme targets too (1: nearly everything is killed and each cell keeps its last photon; 1e7: the
    target energy is below every photon's, so every draw passes)."""
    deck = decks.hot_zone(photons=300_000, t_stop=0.02, scale=4)
    sim = port.OracleSim(deck)
    ctx = None
    for cyc in (1, 2):
        sim.cycle(keep_photons=False)
        if ctx is None:
            ctx = _ctx_for(deck, sim.get("mesh/nodes"))
        ctx.set_cell_data(sim.get("f"), sim.get("op_a"), sim.get("op_s"))
        ctx.source(cyc, sim.get("dt")[0], sim.get("E_emission"), sim.get("E_source"),
                   sim.get("E_census") if cyc == 1 else None, sim.get("global_source_energy")[0])
        ctx.transport(sim.get("next_dt")[0])
        ctx.tallies()
    pre = ctx.download(gpu.LIST_CENSUS)
    n = len(pre["cell"])
    assert n > 20_000
    local_E = float(np.add.reduce(pre["E"]))  # any value works as long as both sides get the same one
    stream = 9 * 10 ** 12 + 7
    keep, new_E, draws = port.comb_photons(pre["cell"], pre["E"], local_E, max_census, deck.seed, stream)
    st = ctx.comb_census(max_census, local_E, stream)
    post = ctx.download(gpu.LIST_CENSUS)
    assert st["n_after"] == int(keep.sum()) and draws == n
    assert np.array_equal(post["stream"], pre["stream"][keep])
    assert np.array_equal(post["E"].view(np.uint64), new_E.view(np.uint64))
    for k in ("cell", "group", "ctr"):
        assert np.array_equal(post[k], pre[k][keep])
    assert np.array_equal(post["pos"].reshape(-1, 3), pre["pos"].reshape(-1, 3)[keep])
    # every cell's census energy is conserved (reference :86-92)
    before = np.bincount(pre["cell"], weights=pre["E"], minlength=deck.n_cells)
    after = np.bincount(post["cell"], weights=post["E"], minlength=deck.n_cells)
    assert np.max(np.abs(before - after)) <= 1e-12 * before.max()
    if max_census == 1:
        assert st["n_after"] == int((before > 0).sum())  # one survivor per populated cell
    if max_census == 10 ** 7:
        assert st["n_after"] == n
    # the combed census feeds the next cycle like any other
    sim.cycle(keep_photons=False)
    ctx.set_cell_data(sim.get("f"), sim.get("op_a"), sim.get("op_s"))
    n_new, n_tot = ctx.source(3, sim.get("dt")[0], sim.get("E_emission"), sim.get("E_source"), None,
                              sim.get("global_source_energy")[0])
    assert n_tot == n_new + st["n_after"]
    ctx.transport(sim.get("next_dt")[0])
    a, t, s2 = ctx.tallies()
    assert s2["n_transported"] == n_tot and np.isfinite(a).all()
    ctx.close()


def test_comb_of_an_empty_census_is_a_no_op():
    deck = decks.big_cube(n=4, photons=100, t_stop=0.001)
    sim = port.OracleSim(deck)
    sim.cycle(keep_photons=False)
    ctx = _ctx_for(deck, sim.get("mesh/nodes"))
    st = ctx.comb_census(10)
    assert st["n_before"] == 0 and st["n_after"] == 0 and st["E_after"] == 0.0
    with pytest.raises(gpu.GpuError):
        ctx.comb_census(0)
    ctx.close()
