#!/usr/bin/env python
"""Per-cycle throughput of the transport kernel on every BASELINE.json deck (one GPU), with the algorithmic bytes of
DESIGN.md section 5 and the HBM fraction they imply.  Not the contract bench (bench.py is); this is the table that says
which regime each deck is in (streaming: HBM / atomics; scattering: alu pipe).

  python tools/bench_configs.py [--out profiles/configs_rNN.json] [--only big_cube,hot_zone] [--algorithm history]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (algorithmic_bytes, measured_peak)
from branson_b200 import decks, driver, gpu  # noqa: E402


def cases(scale_photons: float):
    P = lambda n: max(1000, int(n * scale_photons))  # noqa: E731
    return {
        # name: (deck, cycles)
        "marshak_wave": (decks.marshak_wave(photons=P(1_000_000), t_stop=0.10), 10),
        "hot_zone": (decks.hot_zone(photons=P(1_000_000), t_stop=0.08), 8),
        "hot_zone_1e7": (decks.hot_zone(photons=P(10_000_000), t_stop=0.05), 5),
        "hohlraum_single": (decks.hohlraum_single(photons=P(10_000_000), t_stop=0.05), 5),
        "hohlraum_multi_1gpu_share": (decks.hohlraum_multi(photons=P(31_250_000), t_stop=0.004), 4),
        "big_cube_200": (decks.big_cube(n=200, photons=P(125_000_000), t_stop=0.003), 3),
    }


def run_case(name, deck, cycles, algorithm):
    tmp = tempfile.mkdtemp(prefix="bcfg_")
    xml = deck.write(os.path.join(tmp, f"{name}.xml"))
    d = driver.Driver(xml, n_groups=deck.n_groups, device=0, algorithm=algorithm, mesh_on_device=True)
    peak, peak_src = bench.measured_peak()
    rows = []
    for c in range(cycles):
        r = d.cycle()
        g = r["gpu"]
        ab = bench.algorithmic_bytes(g)
        ms = g["ms_transport"]
        rows.append({
            "cycle": c + 1, "n_transported": g["n_transported"], "n_census": g["n_census"],
            "ms_transport": ms, "ms_source": g["ms_source"], "ms_census": g["ms_census"],
            "histories_per_s": g["n_transported"] / (ms * 1e-3) if ms > 0 else None,
            "events_per_history": g["n_events"] / max(1, g["n_transported"]),
            "scatters_per_history": g["n_scatters"] / max(1, g["n_transported"]),
            "crossings_per_history": g["n_crossings"] / max(1, g["n_transported"]),
            "bytes_per_history": ab / max(1, g["n_transported"]),
            "achieved_GBs": ab / (ms * 1e-3) / 1e9 if ms > 0 else None,
            "hbm_frac": ab / (ms * 1e-3) / 1e9 / peak if ms > 0 else None,
            "Gdraws_per_s": (g["n_events"] + 4 * g["n_scatters"]) / (ms * 1e-3) / 1e9 if ms > 0 else None,
            "rad_balance_rel": abs(r["rad_balance_exact"]) / max(1e-300, r["pre_census_E"] + r["emission_E"] + r["source_E"]),
        })
    d.close()
    return {"deck": name, "n_groups": deck.n_groups, "photons": deck.photons,
            "peak_GBs": peak, "peak_source": peak_src, "cycles": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None)
    ap.add_argument("--scale-photons", type=float, default=1.0)
    ap.add_argument("--algorithm", default="history", choices=["history", "event"])
    a = ap.parse_args()
    algo = gpu.EVENT if a.algorithm == "event" else gpu.HISTORY
    out = []
    for name, (deck, cycles) in cases(a.scale_photons).items():
        if a.only and name not in a.only.split(","):
            continue
        res = run_case(name, deck, cycles, algo)
        out.append(res)
        print(f"== {name}  (G={deck.n_groups}, photons={deck.photons:.3g})")
        print("  cyc   transported     ms    Mhist/s  ev/h  sc/h  cr/h   B/hist   GB/s  hbm%  Gdraw/s")
        for r in res["cycles"]:
            print(f"  {r['cycle']:3d} {r['n_transported']:13d} {r['ms_transport']:7.2f} {r['histories_per_s'] / 1e6:9.1f}"
                  f" {r['events_per_history']:5.1f} {r['scatters_per_history']:5.1f} {r['crossings_per_history']:5.1f}"
                  f" {r['bytes_per_history']:8.0f} {r['achieved_GBs']:6.0f} {100 * r['hbm_frac']:5.1f}"
                  f" {r['Gdraws_per_s']:7.1f}", flush=True)
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
