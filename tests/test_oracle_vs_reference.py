"""Live cross-check: the plain-C oracle against the UNMODIFIED reference binary, bit for bit.

Runs only where oracle/_ref/ref_harness_g* exists (built from /root/reference by `make -C oracle ref`; the binaries
travel to the GPU box).  Covers deck shapes and rank counts beyond the committed fixtures.
"""
import numpy as np
import pytest

from branson_b200 import decks
from oracle import port, refio

CASES = {
    "three_region_g30_r3": (lambda: decks.simple_three_region(photons=4000, n_groups=30), 3),
    "marshak": (lambda: decks.marshak_wave(photons=6000, t_stop=0.03), 1),
    "hot_zone_s5": (lambda: decks.hot_zone(photons=8000, t_stop=0.02, scale=5), 1),
    "hohlraum_s5_r2": (lambda: decks.hohlraum_single(photons=30000, t_stop=0.02, scale=5), 2),
    "big_cube_8": (lambda: decks.big_cube(n=8, photons=6000, t_stop=0.002), 1),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_equals_reference(name, tmp_path):
    mk, n_ranks = CASES[name]
    deck = mk()
    if not refio.have_reference(deck.n_groups):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    dumps, _ = refio.run_reference(deck, n_ranks=n_ranks, workdir=str(tmp_path))
    sim = port.OracleSim(deck, n_ranks=n_ranks)
    cyc = 0
    while not sim.finished():
        cyc += 1
        sim.cycle(keep_photons=True)
        for r in range(n_ranks):
            for k, want in dumps[r].items():
                if not k.startswith(f"c{cyc}/") or k.endswith("transport_seconds"):
                    continue
                got = sim.get(k.split("/", 1)[1], r)
                assert got.shape == want.shape, k
                assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), f"rank {r} {k}"
    assert cyc == int(dumps[0]["cycles_done"][0])
