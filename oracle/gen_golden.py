"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/ref_harness_g*).

Run in the build container (needs /root/reference for `make -C oracle ref`):
    python oracle/gen_golden.py
Each fixture stores, per cycle and per emulated rank, the reference's per-cell arrays and scalars in
full and the first PHOTON_LIMIT photons' pre/post transport records, as raw IEEE doubles / integers.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from branson_b200 import decks  # noqa: E402
from oracle import refio  # noqa: E402

PHOTON_LIMIT = 1000


def golden_cases():
    """name -> (deck, n_ranks).  Also imported by the tests so both sides build identical decks."""
    return {
        "three_region_g1_r1": (decks.simple_three_region(photons=3000, n_groups=1), 1),
        "three_region_g30_r2": (decks.simple_three_region(photons=3000, n_groups=30), 2),
        "marshak_r1": (decks.marshak_wave(photons=3000, t_stop=0.04), 1),
        "hot_zone_s10_r1": (decks.hot_zone(photons=5000, t_stop=0.03, scale=10), 1),
        "hohlraum_s5_g30_r1": (decks.hohlraum_single(photons=20000, t_stop=0.02, scale=5), 1),
        "hohlraum_multi_s10_g30_r4": (decks.hohlraum_multi(photons=8000, t_stop=0.003, scale=10), 4),
    }


def comb_cases():
    """name -> (deck, cycles, max_census_photons, rng stream): comb_photons (reference src/census_functions.h:48-93) run
    by the unmodified reference on the census left after `cycles` cycles."""
    return {
        "comb_hot_zone_s10": (decks.hot_zone(photons=30000, t_stop=0.03, scale=10), 3, 1000, 9 * 10 ** 12),
        "comb_three_region_g30": (decks.simple_three_region(photons=20000, n_groups=30), 2, 60, 9 * 10 ** 12 + 1),
    }


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (deck, cycles, max_census, stream) in comb_cases().items():
        dumps, _ = refio.run_reference(deck, max_cycles=cycles, photon_limit=0, comb_max=max_census, comb_stream=stream)
        flat = {k: v for k, v in dumps[0].items() if k.startswith("comb/") or k in ("seed", "n_cells")}
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **flat)
        print(f"{name}: census {len(flat['comb/pre/cell'])} -> {len(flat['comb/post/cell'])}, "
              f"{os.path.getsize(path) / 1e3:.0f} kB")
    if "--comb-only" in sys.argv:
        return
    for name, (deck, n_ranks) in golden_cases().items():
        dumps, _ = refio.run_reference(deck, n_ranks=n_ranks, photon_limit=PHOTON_LIMIT)
        flat = {}
        for r, d in enumerate(dumps):
            for k, v in d.items():
                if k.endswith("transport_seconds"):
                    continue
                flat[f"r{r}/{k}"] = v
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **flat)
        print(f"{name}: {len(flat)} arrays, {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
