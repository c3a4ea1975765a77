"""The drop-in, proven on the reference's own types and call sites.

`oracle/dropin.patch` is the INTEGRATION.md level-1 patch a Branson maintainer would apply: `GPU_Setup` creates a
`bgpu_ctx` from the mesh cells (reference src/gpu_setup.h:19-40), `gpu_transport_photons` hands the real
`std::vector<Photon>` / `std::vector<Cell_Tally>` storage to `bgpu_transport_photons_aos` (reference
src/history_based_transport.h:348-413), and the GPU branch of `replicated_transport` post-processes like the CPU branch
(reference src/replicated_transport.h:75-86).  `make -C oracle ref_dropin` compiles the reference translation unit with
that patch against `libbranson_gpu.so` (binaries under oracle/_ref/, built where /root/reference exists).

A deck with `use_gpu_transporter TRUE` run through the PATCHED reference must reproduce the UNPATCHED reference's CPU
HISTORY path: every photon's final cell, group, descriptor and RNG counter bit for bit, every cycle (the census the
patched run carries from cycle to cycle went through the device), per-photon doubles and tallies to 1e-9.
"""
import os
import re
import subprocess

import numpy as np
import pytest

from branson_b200 import decks
from oracle import refio

pytestmark = pytest.mark.gpu

CASES = [
    ("hohlraum", lambda: decks.hohlraum_single(photons=40000, t_stop=0.03, scale=5)),          # G = 30, reflect + vacuum
    ("marshak", lambda: decks.marshak_wave(photons=20000, t_stop=0.04)),                        # G = 1, SOURCE face
    ("three_region", lambda: decks.simple_three_region(photons=20000, n_groups=30)),
]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_patched_reference_equals_its_cpu_history_path(name, make, tmp_path):
    deck = make().with_(use_gpu_transporter="TRUE", dd_transport_type="REPLICATED", n_omp_threads=4)
    if not (refio.have_reference(deck.n_groups) and os.path.exists(refio.harness_path(deck.n_groups, dropin=True))):
        pytest.skip("oracle/_ref drop-in binaries are not built (make -C oracle ref ref_dropin)")
    (tmp_path / "cpu").mkdir()
    (tmp_path / "gpu").mkdir()
    cpu, _ = refio.run_reference(deck, workdir=str(tmp_path / "cpu"))
    dev, out = refio.run_reference(deck, workdir=str(tmp_path / "gpu"), dropin=True)
    cpu, dev = cpu[0], dev[0]
    assert "gpu transport time" in out  # the reference's own GPU branch ran (src/replicated_transport.h:78-79)
    n_cycles = int(cpu["cycles_done"][0])
    assert n_cycles == int(dev["cycles_done"][0]) >= 3
    for c in range(1, n_cycles + 1):
        p = f"c{c}/"
        assert int(cpu[p + "n_photons"][0]) == int(dev[p + "n_photons"][0])
        # the photons entering transport are the same (the census of the previous cycle came back from the device)
        for k in ("cell", "group", "ctr", "stream"):
            assert np.array_equal(cpu[p + "pre/" + k], dev[p + "pre/" + k]), (c, "pre", k)
        for k in ("cell", "group", "ctr", "descriptor"):
            assert np.array_equal(cpu[p + "post/" + k], dev[p + "post/" + k]), (c, "post", k)
        for k in ("pos", "angle", "E", "life_dx"):
            a, b = cpu[p + "post/" + k], dev[p + "post/" + k]
            assert np.max(np.abs(a - b)) <= 1e-9 * max(np.max(np.abs(a)), 1e-300), (c, k)
        assert int(cpu[p + "n_census"][0]) == int(dev[p + "n_census"][0])
        for k in ("abs_E", "track_E", "T_e", "T_r"):
            a, b = cpu[p + k], dev[p + k]
            assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(a)), (c, k)
        for k in ("exit_E", "post_census_E", "absorbed_E"):
            a, b = float(cpu[p + k][0]), float(dev[p + k][0])
            assert abs(a - b) <= 1e-9 * max(abs(a), 1e-300), (c, k)


def test_patched_stock_binary_runs_the_deck(tmp_path):
    """the reference's own main() with the patch: `BRANSON deck.xml` end to end, GPU branch taken every cycle"""
    deck = decks.hohlraum_single(photons=40000, t_stop=0.03, scale=5).with_(use_gpu_transporter="TRUE", n_omp_threads=4)
    exe_cpu, exe_gpu = refio.stock_binary_path(30), refio.stock_binary_path(30, dropin=True)
    if not (os.path.exists(exe_cpu) and os.path.exists(exe_gpu)):
        pytest.skip("oracle/_ref binaries are not built")
    xml = deck.write(str(tmp_path / "deck.xml"))
    env = dict(os.environ, BRANSON_SHIM_NRANKS="1")
    outs = []
    for exe in (exe_cpu, exe_gpu):
        r = subprocess.run([exe, xml], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs.append(r.stdout)
    assert outs[1].count("gpu transport time") == 3 and "gpu transport time" not in outs[0]
    for pat in (r"Total Photons transported: (\d+)", r"Post census Size: (\d+)"):
        assert re.findall(pat, outs[0]) == re.findall(pat, outs[1]), pat
    for pat in (r"Absorption E: ([0-9.eE+-]+)", r"Exit E: ([0-9.eE+-]+)", r"Post census E: ([0-9.eE+-]+)"):
        a = [float(x) for x in re.findall(pat, outs[0])]
        b = [float(x) for x in re.findall(pat, outs[1])]
        assert len(a) == len(b) == 3 and np.allclose(a, b, rtol=1e-5), (pat, a, b)
    for x in re.findall(r"Radiation conservation: ([0-9.eE+-]+)", outs[1]):
        assert abs(float(x)) < 1e-10
