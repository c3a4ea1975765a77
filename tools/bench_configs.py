#!/usr/bin/env python
"""Per-cycle throughput of the transport kernel on every BASELINE.json deck (one GPU), with the algorithmic bytes of
DESIGN.md section 5 and the HBM fraction they imply.  Not the contract bench (bench.py is); this is the table that says
which regime each deck is in (streaming: HBM / atomics; scattering: alu pipe).

  python tools/bench_configs.py [--out profiles/configs_rNN.json] [--only big_cube,hot_zone] [--algorithm history]

Under torchrun (WORLD_SIZE > 1) the decks are the replicated multi-GPU configurations of BASELINE.json at their named
sizes -- rank r of the launch plays replicated-mode rank r, photons partitioned, one NCCL tally all-reduce per cycle:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      tools/bench_configs.py --multi [--only hohlraum_multi,big_cube_200]
Per cycle: histories of all ranks / the slowest rank's device-timed transport, and the same over the whole cycle's
wall time (source + transport + census + all-reduce + mesh update, barrier to barrier).
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


def multi_cases(world: int):
    return {
        # configs[3]: 2.5e8 photons over the ranks, dt 0.001 (inputs/3D_hohlraum_multi_node.xml, forced REPLICATED)
        "hohlraum_multi": (decks.hohlraum_multi(photons=250_000_000, t_stop=0.005), 5),
        # configs[4]: big_cube scaled to 200^3 cells, 1.25e8 photons per GPU (1e9 at 8 GPUs)
        "big_cube_200": (decks.big_cube(n=200, photons=125_000_000 * world, t_stop=0.003), 3),
    }


def run_case_multi(name, deck, cycles, algorithm):
    import time

    import torch
    import torch.distributed as dist
    rank, world, local = dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", "0"))
    tmp = tempfile.mkdtemp(prefix="bcfg_")
    xml = deck.write(os.path.join(tmp, f"{name}_{rank}.xml"))
    d = driver.Driver(xml, n_groups=deck.n_groups, rank=rank, n_ranks=world, device=local, algorithm=algorithm,
                      mesh_on_device=True)
    driver.init_nccl(d, dist)  # the cycle's collective is native (csrc/comm_native.cuh)
    rows = []
    for c in range(cycles):
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        t0 = time.perf_counter()
        r = d.cycle()
        d.array("T_e")
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        wall = time.perf_counter() - t0
        g = r["gpu"]
        v = torch.tensor([g["ms_transport"], wall * 1e3], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        n = torch.tensor([g["n_transported"], g["n_census"], g["n_events"], g["n_scatters"], g["n_crossings"]],
                         dtype=torch.float64, device=f"cuda:{local}")
        lo = n.clone()
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        ms, wall_ms = (float(x) for x in v.tolist())
        tot = [float(x) for x in n.tolist()]
        total = r["pre_census_E"] + r["emission_E"] + r["source_E"]
        rows.append({"cycle": c + 1, "n_transported": int(tot[0]), "n_transported_min_rank": int(lo[0].item()),
                     "n_census": int(tot[1]), "ms_transport_max": ms, "ms_cycle_wall_max": wall_ms,
                     "histories_per_s": tot[0] / (ms * 1e-3), "histories_per_s_whole_cycle": tot[0] / (wall_ms * 1e-3),
                     "events_per_history": tot[2] / max(1.0, tot[0]), "scatters_per_history": tot[3] / max(1.0, tot[0]),
                     "crossings_per_history": tot[4] / max(1.0, tot[0]),
                     "rad_balance_rel": abs(r["rad_balance_exact"]) / max(1e-300, total)})
    d.close()
    return {"deck": name, "n_groups": deck.n_groups, "photons": deck.photons, "n_gpus": world, "cycles": rows}


def main_multi(a, algo):
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    out = []
    for name, (deck, cycles) in multi_cases(world).items():
        if a.only and name not in a.only.split(","):
            continue
        res = run_case_multi(name, deck, cycles, algo)
        out.append(res)
        if rank == 0:
            print(f"== {name}  (G={deck.n_groups}, photons={deck.photons:.3g} over {world} GPUs)")
            print("  cyc   transported  min/rank      ms_max  Mhist/s   cycle_ms  Mhist/s(cycle)  ev/h  sc/h  cr/h  balance")
            for r in res["cycles"]:
                print(f"  {r['cycle']:3d} {r['n_transported']:13d} {r['n_transported_min_rank']:9d} {r['ms_transport_max']:10.2f}"
                      f" {r['histories_per_s'] / 1e6:8.1f} {r['ms_cycle_wall_max']:10.2f} {r['histories_per_s_whole_cycle'] / 1e6:12.1f}"
                      f" {r['events_per_history']:7.1f} {r['scatters_per_history']:5.1f} {r['crossings_per_history']:5.1f}"
                      f" {r['rad_balance_rel']:8.1e}", flush=True)
    if a.out and rank == 0:
        with open(a.out, "w") as fh:
            json.dump(out, fh, indent=1)
    dist.destroy_process_group()


def run_case(name, deck, cycles, algorithm, sort_census=False):
    tmp = tempfile.mkdtemp(prefix="bcfg_")
    xml = deck.write(os.path.join(tmp, f"{name}.xml"))
    d = driver.Driver(xml, n_groups=deck.n_groups, device=0, algorithm=algorithm, mesh_on_device=True,
                      sort_census=sort_census)
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
            "kernel": {0: "history", 1: "queues", 2: "passes"}.get(g.get("transport_kernel", 0), "?"),
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
    ap.add_argument("--sort-census", action="store_true", help="driver option sort_census (census ordered by cell)")
    ap.add_argument("--multi", action="store_true", help="the multi-GPU decks, one rank per GPU (run under torchrun)")
    a = ap.parse_args()
    algo = gpu.EVENT if a.algorithm == "event" else gpu.HISTORY
    if a.multi:
        return main_multi(a, algo)
    out = []
    for name, (deck, cycles) in cases(a.scale_photons).items():
        if a.only and name not in a.only.split(","):
            continue
        res = run_case(name, deck, cycles, algo, a.sort_census)
        out.append(res)
        print(f"== {name}  (G={deck.n_groups}, photons={deck.photons:.3g})")
        print("  cyc   transported     ms    Mhist/s  ev/h  sc/h  cr/h   B/hist   GB/s  hbm%  Gdraw/s  kernel")
        for r in res["cycles"]:
            print(f"  {r['cycle']:3d} {r['n_transported']:13d} {r['ms_transport']:7.2f} {r['histories_per_s'] / 1e6:9.1f}"
                  f" {r['events_per_history']:5.1f} {r['scatters_per_history']:5.1f} {r['crossings_per_history']:5.1f}"
                  f" {r['bytes_per_history']:8.0f} {r['achieved_GBs']:6.0f} {100 * r['hbm_frac']:5.1f}"
                  f" {r['Gdraws_per_s']:7.1f}  {r['kernel']}", flush=True)
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
