"""Throughput probe (development tool, not the bench): full-size hohlraum mesh on one GPU, host-side per-cycle
quantities taken from a low-statistics oracle run (same mesh, fewer photons), device runs `--photons` per cycle."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from branson_b200 import decks, gpu  # noqa: E402
from oracle import port  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--deck", default="hohlraum")
ap.add_argument("--photons", type=int, default=10_000_000)
ap.add_argument("--oracle-photons", type=int, default=100_000)
ap.add_argument("--cycles", type=int, default=4)
ap.add_argument("--scale", type=int, default=1)
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--blocks-per-sm", type=int, default=0)
ap.add_argument("--algorithm", type=int, default=0)
a = ap.parse_args()

if a.deck == "hohlraum":
    small = decks.hohlraum_single(photons=a.oracle_photons, t_stop=0.01 * a.cycles, scale=a.scale)
elif a.deck == "hot_zone":
    small = decks.hot_zone(photons=a.oracle_photons, t_stop=0.01 * a.cycles, scale=a.scale)
elif a.deck == "big_cube":
    small = decks.big_cube(n=200 // a.scale, photons=a.oracle_photons, t_stop=0.001 * a.cycles)
else:
    small = decks.marshak_wave(photons=a.oracle_photons, t_stop=0.01 * a.cycles)
sim = port.OracleSim(small)
ctx = None
for cyc in range(1, a.cycles + 1):
    t0 = time.time()
    sim.cycle(keep_photons=False)
    t_or = time.time() - t0
    if ctx is None:
        ctx = gpu.context_for_deck(small, sim.get("mesh/nodes"), n_user_photons=a.photons, device=0)
        ctx.set_launch(blocks_per_sm=a.blocks_per_sm, chunk=a.chunk)
    ctx.set_cell_data(sim.get("f"), sim.get("op_a"), sim.get("op_s"))
    t0 = time.time()
    n_new, n_tot = ctx.source(cyc, sim.get("dt")[0], sim.get("E_emission"), sim.get("E_source"),
                              sim.get("E_census") if cyc == 1 else None, sim.get("global_source_energy")[0])
    ctx.transport(sim.get("next_dt")[0], a.algorithm, gpu.TALLY_ATOMIC)
    wall = time.time() - t0
    ab, tr, st = ctx.tallies()
    ev = st["n_events"] / max(1, n_tot)
    print(f"cycle {cyc}: photons {n_tot} (new {n_new}) src {st['ms_source']:.2f} ms transport {st['ms_transport']:.2f} ms "
          f"census {st['ms_census']:.2f} ms wall {wall*1e3:.1f} ms -> {n_tot / (st['ms_transport'] * 1e-3):.3e} hist/s "
          f"(kernel) | events/hist {ev:.1f} scat {st['n_scatters']/max(1,n_tot):.1f} cross {st['n_crossings']/max(1,n_tot):.1f} "
          f"dep {st['n_deposits']/max(1,n_tot):.1f} lookups {st['n_group_lookups']/max(1,n_tot):.1f} census {st['n_census']} "
          f"| oracle {t_or:.1f}s abs_sum {ab.sum():.6e} vs oracle {sim.get('abs_E').sum():.6e}", flush=True)
