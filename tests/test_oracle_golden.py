"""Pin the plain-C oracle (oracle/imc_oracle.c) against vectors produced by the UNMODIFIED reference.

* Known-answer tests: Threefry2x64-20 vectors equal Random123's published KATs; RNG / angle values were printed by
  the reference build; distance and move answers are the reference's own unit-test
  expectations (reference src/test/test_cell.cc:80-133, src/test/test_photon.cc:70-126).
* tests/golden/*.npz: full per-cycle dumps of the reference (oracle/gen_golden.py).  The oracle must reproduce every
  array BIT FOR BIT (same machine family: glibc libm, no FMA contraction).
"""
import glob
import os

import numpy as np
import pytest

from oracle import port
from oracle.gen_golden import PHOTON_LIMIT, comb_cases, golden_cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_threefry_known_answers():
    assert port.threefry([0, 0], [0, 0]) == (0xc2b6e3a8c2c69865, 0x6f81ed42f350084d)
    m = 2 ** 64 - 1
    assert port.threefry([m, m], [m, m]) == (0xe02cb7c4d95d277a, 0xd06633d0893b8b68)
    assert port.threefry([0x243f6a8885a308d3, 0x13198a2e03707344], [0xa4093822299f31d0, 0x082efa98ec4e6c89]) == (
        0x263c7d30bb0f0af1, 0x56be8361d3311526)


def test_rng_known_answers():
    np.testing.assert_array_equal(port.rng_draws(777, 1, 4), [0.50060536157985436, 0.29511474933665161,
                                                              0.92517102853492827, 0.55080394150283085])
    np.testing.assert_array_equal(port.rng_draws(14706, 0, 4), [0.10704214220457897, 0.37725928982588541,
                                                                0.96811350044478395, 0.7575540980655614])
    np.testing.assert_array_equal(port.uniform_angle(341, 5), [-0.014991543958673631, -0.91278287879468278,
                                                               0.40816990309064516])


def test_rng_properties_like_reference_unit_test():
    # reference src/test/test_counter_rng.cc:27-69 (range, mean) and :72-174 (determinism / stream separation)
    x = port.rng_draws(1234, 7, 20000)
    assert x.min() > 0.0 and x.max() < 1.0
    assert abs(x.mean() - 0.5) < 1e-2
    np.testing.assert_array_equal(x[:100], port.rng_draws(1234, 7, 100))
    assert not np.array_equal(x[:100], port.rng_draws(1234, 8, 100))
    assert not np.array_equal(x[:100], port.rng_draws(1235, 7, 100))


def test_distance_to_boundary_reference_unit_test():
    # reference src/test/test_cell.cc:80-133: unit cube, from the centre
    import ctypes as C
    nodes = np.array([0.0, 1.0, 0.0, 1.0, 0.0, 1.0])
    pos = np.array([0.5, 0.5, 0.5])
    L = port.lib()
    for axis in range(3):
        for sgn in (+1, -1):
            ang = np.full(3, 0.001 * 0.5)  # small positive off-axis components
            ang[axis] = sgn * 0.999
            s = C.c_uint32(99)
            d = L.orc_distance_to_boundary(port._ptr(nodes), port._ptr(pos), port._ptr(ang), C.byref(s))
            assert abs(d - 0.5 / 0.999) < 1e-8
            assert s.value == 2 * axis + (1 if sgn > 0 else 0)


@pytest.mark.parametrize("name", sorted(golden_cases().keys()))
def test_oracle_reproduces_reference_bitwise(name):
    deck, n_ranks = golden_cases()[name]
    path = os.path.join(GOLDEN, name + ".npz")
    assert os.path.exists(path), "golden fixture missing: run oracle/gen_golden.py"
    gold = np.load(path)
    sim = port.OracleSim(deck, n_ranks=n_ranks)
    n_cycles = int(gold["r0/cycles_done"][0])
    assert n_cycles == deck.n_cycles()
    checked = 0
    for cyc in range(1, n_cycles + 1):
        assert not sim.finished()
        sim.cycle(keep_photons=True)
        for r in range(n_ranks):
            pfx = f"r{r}/c{cyc}/"
            for k in gold.files:
                if not k.startswith(pfx):
                    continue
                short = k[len(pfx):]
                want = gold[k]
                got = sim.get(short, r)
                if short.startswith(("pre/", "post/")):
                    per = 3 if short.endswith(("pos", "angle")) else 1
                    got = got[:PHOTON_LIMIT * per]
                assert got.dtype == want.dtype, k
                assert got.shape == want.shape, k
                assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), f"{k} differs from the reference"
                checked += 1
    assert sim.finished()
    assert checked > 40 * n_cycles * n_ranks
    # mesh tables (reference src/proto_mesh.h:106-218)
    for k in ("mesh/nodes", "mesh/region", "mesh/e_next", "mesh/bc"):
        assert np.array_equal(sim.get(k), gold["r0/" + k]), k


def test_golden_fixture_inventory():
    have = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))}
    assert have == set(golden_cases().keys()) | set(comb_cases().keys())


@pytest.mark.parametrize("name", sorted(comb_cases()))
def test_comb_photons_matches_reference_fixture(name):
    """orc_comb_photons against comb_photons of the unmodified reference (src/census_functions.h:48-93), run by
    oracle/ref_harness.cc on a real census: same survivors in the same order, corrected energies bit for bit, same number
    of RNG draws; and the comb conserves the census energy of every cell."""
    deck, cycles, max_census, stream = comb_cases()[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cell, E = g["comb/pre/cell"], g["comb/pre/E"]
    assert int(g["comb/max_census_photons"][0]) == max_census and int(g["comb/rng_stream"][0]) == stream
    keep, new_E, draws = port.comb_photons(cell, E, g["comb/local_census_E"][0], max_census, deck.seed, stream)
    assert draws == int(g["comb/rng_draws"][0]) == len(cell)
    assert np.array_equal(g["comb/pre/stream"][keep], g["comb/post/stream"])
    assert np.array_equal(cell[keep], g["comb/post/cell"])
    assert np.array_equal(new_E.view(np.uint64), g["comb/post/E"].view(np.uint64))
    before = np.bincount(cell, weights=E, minlength=int(g["n_cells"][0]))
    after = np.bincount(cell[keep], weights=new_E, minlength=int(g["n_cells"][0]))
    assert np.max(np.abs(before - after)) <= 1e-13 * before.max()
