#!/usr/bin/env python
"""bench.py -- photon histories/sec of the replicated IMC cycle on the 30-group 3-D hohlraum (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU implementation on the host cores

A "step" is one IMC cycle = one pass of the hot path (source -> transport -> census/tally) over that cycle's photons.
Workload at N = 1: BASELINE.json configs[2], `3D_hohlraum_single_node` (65x65x140 cells, 30 groups, 1e7 user photons
per cycle); at N > 1 the same deck with 1e7 x N user photons, rank r playing replicated-mode rank r ("weak" scaling:
per-GPU photons fixed) and one in-place NCCL all-reduce of the packed tally buffer per cycle.  Synthetic problem: the
deck is generated (branson_b200/decks.py) from the reference deck's numbers; all randomness is Threefry(seed, stream).

JSON line (rank 0):
  value     whole-job histories/s over the K timed cycles, device-resident (CUDA events on the ctx stream around
            source + transport + census of every cycle; max over ranks)
  e2e       the same K cycles timed by wall clock through the public cycle API with HOST buffers: host physics
            (calculate_photon_energy / update_temperature), H2D of f/op_a/op_s and the source energies, the device
            work, the tally all-reduce and the D2H of the tallies
  roofline  the history-transport kernel: algorithmic bytes (counted exactly by the kernel)
            / its CUDA-event time, against the measured HBM copy bandwidth
  cpu_baseline  the UNMODIFIED reference (oracle/_ref/branson_ref_g30, OpenMP on all host cores) on a bounded
            sample of the same workload (N = 1, rank 0 only)
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_GROUPS = 30
PHOTONS_PER_GPU = 10_000_000
DT = 0.01
METRIC = "photon histories/sec (3D hohlraum, 30 groups, replicated)"
UNIT = "histories/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent
# Threefry2x64-20 alone on all 148 SMs (tools/ubench/threefry_variants.cu, best formulation, measured on this pool's
# B200: profiles/threefry_ubench_r01.txt): the alu-pipe ceiling of the draws the reference's algorithm prescribes
THREEFRY_PEAK_GDRAWS = 181.5


def make_deck(n_gpus: int, cycles: int, photons_per_gpu: int = PHOTONS_PER_GPU, threads: int = 1):
    from branson_b200 import decks
    d = decks.hohlraum_single(photons=photons_per_gpu * n_gpus, t_stop=DT * cycles)
    return d.with_(n_omp_threads=threads, dd_transport_type="REPLICATED")


REF_SAMPLE_PHOTONS = 2_000_000  # user photons per cycle of the reference arm's / cpu_baseline's bounded sample


def config_dict(n_gpus: int, photons_per_gpu: int, warmup: int, steps: int, ref_photons: int = REF_SAMPLE_PHOTONS):
    """The same dict in both arms (ours and --impl reference): it names the workload AND says what the CPU arm runs."""
    return {"workload": "3D_hohlraum_single_node (BASELINE configs[2]): 65x65x140 cells, N_GROUPS=30, REPLICATED, "
                        f"{photons_per_gpu:.0e} user photons per cycle per GPU, dt=0.01, HISTORY algorithm, atomic tallies; "
                        f"a step is one IMC cycle; the deck's 5 cycles are run on at the same dt: cycles 1-{warmup} warm "
                        f"up, cycles {warmup + 1}-{warmup + steps} are timed (scattering-dominated steady state; cycle 1 "
                        "is the streaming transient)",
            "photons_per_cycle": photons_per_gpu * n_gpus, "n_cells": 591500, "n_groups": N_GROUPS,
            "cycles_timed": [warmup + 1, warmup + steps],
            "parallelism": f"replicated x{n_gpus} (photons partitioned, one packed tally all-reduce per cycle)",
            "l2": "inputs larger than L2 (>= 1 GB of photon state per GPU and cycle; no explicit flush)",
            "cpu_arm_sample_photons_per_cycle": ref_photons,
            "cpu_arm_note": f"the CPU arms (--impl reference, cpu_baseline) run the same deck and the same cycle window at "
                            f"{ref_photons} user photons per cycle -- a bounded sample: the full {photons_per_gpu:.0e} "
                            "would take ~15 s per cycle on 16 cores"}


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref; the oracle port if the reference binary is absent)
# ----------------------------------------------------------------------------------------------------------------
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_cpu(cycles: int, photons: int, threads: int, timeout: float = 1500.0):
    """Runs the unmodified reference binary on a reduced-photon copy of the bench deck.  Returns per-cycle
    (photons transported, transport seconds, source seconds) parsed from its own report, and the kind."""
    from oracle import refio
    deck = make_deck(1, cycles, photons_per_gpu=photons, threads=threads)
    exe = refio.stock_binary_path(N_GROUPS)
    if os.path.exists(exe):
        tmp = tempfile.mkdtemp(prefix="branson_cpu_")
        xml = deck.write(os.path.join(tmp, "deck.xml"))
        env = dict(os.environ, BRANSON_SHIM_NRANKS="1", OMP_NUM_THREADS=str(threads))
        t0 = time.time()
        res = subprocess.run([exe, xml], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                             timeout=timeout)
        wall = time.time() - t0
        if res.returncode != 0:
            raise RuntimeError("reference binary failed:\n" + res.stdout[-2000:])
        n = [int(x) for x in re.findall(r"Total Photons transported: (\d+)", res.stdout)]
        t = [float(x) for x in re.findall(r"Transport time max/min: ([0-9.eE+-]+)/", res.stdout)]
        s = [float(x) for x in re.findall(r"source time: ([0-9.eE+-]+)", res.stdout)]
        assert len(n) == len(t) == cycles, res.stdout[-2000:]
        return {"kind": "reference", "cores": threads, "photons": n, "transport_s": t, "source_s": s, "wall_s": wall}
    # fallback: the plain-C restatement (scalar, one core)
    from oracle import port
    sim = port.OracleSim(deck)
    n, t = [], []
    t0 = time.time()
    for _ in range(cycles):
        sim.cycle(keep_photons=False)
        n.append(int(sim.get("n_photons")[0]))
        t.append(sim.transport_seconds())
    return {"kind": "port", "cores": 1, "photons": n, "transport_s": t, "source_s": [0.0] * cycles,
            "wall_s": time.time() - t0}


def cpu_sample_text(r, photons, warmup, steps, hist, t_tr):
    return (f"the unmodified reference binary (OpenMP, {r['cores']} threads) on the same deck at {photons} user photons "
            f"per cycle ({hist // steps} histories per cycle incl. the one-per-cell minimum), cycles 1-{warmup + steps} run, "
            f"cycles {warmup + 1}-{warmup + steps} quoted: {hist} histories in {t_tr:.2f} s of the reference's own "
            f"'Transport time' (serial sourcing {sum(r['source_s'][warmup:]):.2f} s extra; whole run {r['wall_s']:.1f} s)")


def cpu_baseline_block(sample_photons: int, warmup: int, steps: int):
    """One protocol for every CPU number of a record: same deck, same cycle window as the timed GPU cycles."""
    threads = host_cores()
    r = run_reference_cpu(warmup + steps, sample_photons, threads)
    hist = sum(r["photons"][warmup:])
    secs = sum(r["transport_s"][warmup:])
    return {"value": hist / secs, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
            "sample": cpu_sample_text(r, sample_photons, warmup, steps, hist, secs)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_cores()
    sample = args.ref_photons
    cycles = args.warmup + args.steps
    r = run_reference_cpu(cycles, sample, threads)
    hist = sum(r["photons"][args.warmup:])
    t_tr = sum(r["transport_s"][args.warmup:])
    t_all = t_tr + sum(r["source_s"][args.warmup:])
    line = {"impl": "reference", "metric": METRIC, "value": hist / t_tr, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tr / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.gpus, args.photons, args.warmup, args.steps, sample),
            "cpu_baseline": {"value": hist / t_tr, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                             "sample": cpu_sample_text(r, sample, args.warmup, args.steps, hist, t_tr)},
            "e2e": {"value": hist / t_all, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(g: dict) -> float:
    """DESIGN.md section 5 / BASELINE.md section 4: bytes one launch of the transport kernel must move."""
    n_hist, n_visit = g["n_transported"], g["n_transported"] + g["n_crossings"]
    return (n_hist * (96 + 9) + n_visit * 8 + g["n_group_lookups"] * 16 + g["n_deposits"] * 32 + g["n_census"] * 96)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBps"):
                if k in j:
                    return float(j[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def ncu_traffic():
    """dram bytes per launch of the transport kernel from the committed ncu capture of this workload, else None."""
    p = os.path.join(ROOT, "profiles", "transport_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def aos_dropin_block(d, n_groups: int, device: int, repeats: int = 2):
    """The literal drop-in for the reference's gpu_transport_photons(rank_cell_offset, std::vector<Photon>&, const Cell*,
    std::vector<Cell_Tally>&) (src/history_based_transport.h:348-413), timed through the C ABI with HOST buffers: one call
    of bgpu_transport_photons_aos on the next cycle's freshly sourced photons as 120-byte AoS records in host memory (both
    PCIe directions of the photon list and of the tally array inside the timed region)."""
    import numpy as np

    from branson_b200 import gpu
    f, op_a, op_s = d.array("f"), d.array("op_a"), d.array("op_s")
    total_E = d.calculate_photon_energy()  # the next cycle's emission / source energies from the current temperatures
    E_em, E_src = d.array("E_emission"), d.array("E_source")
    nx, ny, nz = (int(d.param(k)) for k in ("nx", "ny", "nz"))
    seed, n_user = int(d.param("seed")), int(d.param("n_user_photons"))
    bc = ("REFLECT", "VACUUM", "REFLECT", "VACUUM", "VACUUM", "VACUUM")  # decks.hohlraum_single
    ctx = gpu.Context(n_groups, nx, ny, nz, d.array("x_faces"), d.array("y_faces"), d.array("z_faces"), bc, seed,
                      n_user, device=device)
    try:
        ctx.set_cell_data(f, op_a, op_s)
        n_new, n_tot = ctx.source(int(d.param("step")), d.param("dt"), E_em, E_src, None, total_E)
        soa = ctx.download(gpu.LIST_WORK)
        aos0 = gpu.aos_from_soa(soa, seed)
        del soa
        best, hist = None, n_tot
        for _ in range(repeats):
            aos = aos0.copy()
            tal = np.zeros((nx * ny * nz, 2))
            t0 = time.perf_counter()
            ctx.transport_photons_aos(aos, tal)
            dt_s = time.perf_counter() - t0
            best = dt_s if best is None else min(best, dt_s)
        rec = aos.view(np.uint64).reshape(-1, 15)
        desc = ((rec[:, 1] >> np.uint64(32)) & np.uint64(0xff)).astype(np.uint8)
        assert not (desc == gpu.PASS).any(), "drop-in left photons unprocessed"
        E_left = rec[:, 8].copy().view(np.float64)
        e_in = rec[:, 9].copy().view(np.float64).sum()
        e_out = tal[:, 0].sum() + E_left[desc != gpu.KILLED].sum()
        return {"value": hist / best, "unit": UNIT, "photons": int(hist), "ms": 1e3 * best,
                "h2d_bytes_per_step": int(120 * hist + 16 * nx * ny * nz),
                "d2h_bytes_per_step": int(120 * hist + 16 * nx * ny * nz),
                "rel_energy_balance": float(abs(e_in - e_out) / e_in),
                "path": "bgpu_transport_photons_aos: the reference's std::vector<Photon> (120-byte AoS) and "
                        "std::vector<Cell_Tally> in pageable host memory, updated in place (best of %d calls)" % repeats}
    finally:
        ctx.close()


def parity_check(world: int, rank: int, local: int, dist):
    """Two cycles of hohlraum_multi (mesh / 5) through exactly the path timed below -- `world` ranks, device mesh,
    k_mesh_redistribute, the native packed all-reduce, k_mesh_update_temperature -- against the `world`-rank oracle, on
    every rank: per-photon integers bit for bit, T_e / T_r / tallies 1e-9, conservation 1e-12.  (The oracle is the
    checker here, outside every timed region; reference src/mesh.h:291-315, src/replicated_driver.h:91-104.)"""
    import numpy as np
    import torch

    from branson_b200 import decks, driver, gpu
    from oracle import port
    deck = decks.hohlraum_multi(photons=60000, t_stop=0.002, scale=5).with_(dd_transport_type="REPLICATED")
    tmp = tempfile.mkdtemp(prefix="branson_parity_")
    d = driver.Driver(deck.write(os.path.join(tmp, f"deck_{rank}.xml")), n_groups=deck.n_groups, rank=rank,
                      n_ranks=world, device=local, validate=True, mesh_on_device=True)
    if world > 1:
        driver.init_nccl(d, dist)
    view = d.gpu_context()
    sim = port.OracleSim(deck, n_ranks=world)
    ok, what, cycles = True, "", 0
    try:
        while not sim.finished():
            cycles += 1
            sim.cycle(keep_photons=True)
            rep = d.cycle()
            post = view.download(gpu.LIST_WORK, counters=True)
            for k in ("cell", "group", "ctr", "descriptor", "counters"):
                if not np.array_equal(post[k], sim.get("post/" + k, rank)):
                    ok, what = False, f"cycle {cycles} rank {rank}: post/{k}"
            for k in ("abs_E", "track_E", "T_e", "T_r"):
                want = sim.get(k)
                if not np.max(np.abs(d.array(k) - want)) <= 1e-9 * np.max(np.abs(want)):
                    ok, what = False, f"cycle {cycles} rank {rank}: {k}"
            total = rep["pre_census_E"] + rep["emission_E"] + rep["source_E"]
            if not abs(rep["rad_conservation"]) <= 1e-12 * total:
                ok, what = False, f"cycle {cycles} rank {rank}: radiation conservation {rep['rad_conservation']:.3e}"
            if rep["trans_particles"] != sum(int(sim.get("n_photons", r)[0]) for r in range(world)):
                ok, what = False, f"cycle {cycles} rank {rank}: global photon count"
    finally:
        kind = gpu.comm_info(view._h)["kind"]
        d.close()
    if world > 1:
        flag = torch.tensor([1.0 if ok else 0.0], device=f"cuda:{local}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        all_ok = bool(flag.item() > 0.5)
    else:
        all_ok = ok
    if not ok:
        print(f"[bench] PARITY FAILURE: {what}", file=sys.stderr, flush=True)
    return {"n_ranks": world, "ok": all_ok, "cycles": cycles, "deck": "hohlraum_multi(scale=5), 60000 photons",
            "against": f"{world}-rank oracle (oracle/imc_oracle.c, pinned to the unmodified reference)",
            "collective": {gpu.COMM_NONE: "none (1 rank)", gpu.COMM_NCCL: "native ncclAllReduce, one per cycle",
                           gpu.COMM_LOCAL: "in-process"}[kind],
            "checked": "per-photon cell/group/rng counter/descriptor/event counters bit-exact on every rank; "
                       "abs_E, track_E, T_e, T_r 1e-9; radiation conservation 1e-12"}


def extra_configs(world: int):
    from branson_b200 import decks
    return {
        # configs[3]: 3D_hohlraum_multi_node.xml forced REPLICATED, 2.5e8 photons over the ranks, dt 0.001, 5 cycles
        "hohlraum_multi": (decks.hohlraum_multi(photons=250_000_000, t_stop=0.005), 5),
        # configs[4]: big_cube.xml scaled to 200^3 cells, 1.25e8 photons per GPU (1e9 at 8 GPUs), 5 cycles
        "big_cube_200": (decks.big_cube(n=200, photons=125_000_000 * world, t_stop=0.005), 5),
    }


def run_extra_config(name, deck, cycles, world, rank, local, dist):
    """`cycles` cycles of one more BASELINE deck on all ranks: histories of all ranks / the slowest rank's device-timed
    transport, and / the whole cycle's wall time (barrier to barrier).  The first two cycles (streaming transient,
    photon lists still growing) are reported but not part of the quoted figures."""
    import torch

    from branson_b200 import driver, gpu
    tmp = tempfile.mkdtemp(prefix="branson_cfg_")
    d = driver.Driver(deck.with_(dd_transport_type="REPLICATED").write(os.path.join(tmp, f"{name}_{rank}.xml")),
                      n_groups=deck.n_groups, rank=rank, n_ranks=world, device=local, mesh_on_device=True)
    driver.init_nccl(d, dist)
    rows = []
    for c in range(cycles):
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        t0 = time.perf_counter()
        r = d.cycle()
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        wall = time.perf_counter() - t0
        g = r["gpu"]
        v = torch.tensor([g["ms_transport"], wall * 1e3, g["ms_source"], g["ms_census"]], dtype=torch.float64,
                         device=f"cuda:{local}")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        total = r["pre_census_E"] + r["emission_E"] + r["source_E"]
        rows.append({"cycle": c + 1, "histories": int(r["trans_particles"]), "ms_transport_max": v[0].item(),
                     "algorithmic_GBs_rank0": algorithmic_bytes(g) / max(1e-9, g["ms_transport"] * 1e-3) / 1e9,
                     "kernel_rank0": {0: "history", 1: "queues", 2: "passes"}.get(g.get("transport_kernel", 0), "?"),
                     "ms_cycle_wall_max": v[1].item(), "ms_source_max": v[2].item(), "ms_census_max": v[3].item(),
                     "events_per_history_rank0": g["n_events"] / max(1, g["n_transported"]),
                     "rad_balance_rel": abs(r["rad_balance_exact"]) / max(1e-300, total)})
    info = gpu.comm_info(d.gpu_context()._h)
    d.close()
    # quoted: from the third cycle on -- the first is the streaming transient, the second still grows the photon lists
    # (cudaMalloc inside the cycle)
    q = rows[2:] if len(rows) > 3 else rows[-1:]
    hist = sum(x["histories"] for x in q)
    return {"deck": name, "n_gpus": world, "photons_per_cycle": deck.photons, "n_groups": deck.n_groups,
            "n_cells": deck.n_cells, "cycles_quoted": [q[0]["cycle"], q[-1]["cycle"]],
            "transport_histories_per_s": hist / (sum(x["ms_transport_max"] for x in q) * 1e-3),
            "whole_cycle_histories_per_s": hist / (sum(x["ms_cycle_wall_max"] for x in q) * 1e-3),
            "ms_transport_per_cycle": sum(x["ms_transport_max"] for x in q) / len(q),
            "ms_whole_cycle": sum(x["ms_cycle_wall_max"] for x in q) / len(q),
            "collectives_per_cycle": info["calls"] / cycles, "max_rad_balance_rel": max(x["rad_balance_rel"] for x in rows),
            "roofline_rank0": {"bound": "hbm", "achieved": sum(x["algorithmic_GBs_rank0"] for x in q) / len(q),
                               "peak": measured_peak()[0], "unit": "GB/s",
                               "frac": sum(x["algorithmic_GBs_rank0"] for x in q) / len(q) / measured_peak()[0],
                               "kernel": q[-1]["kernel_rank0"]},
            "unit": UNIT, "cycles": rows}


def our_arm(args):
    # host physics (calculate_photon_energy / update_temperature) is OpenMP: give each rank its share of the cores
    # (torchrun would otherwise pin every process to OMP_NUM_THREADS=1)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    os.environ["OMP_NUM_THREADS"] = str(max(1, host_cores() // max(1, local_world)))
    import torch

    from branson_b200 import driver, gpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # the driver launches N > 1 through torchrun; a bare call re-executes itself that way
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__] + sys.argv[1:]
            return subprocess.call(cmd)
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        # torch.distributed is launch plumbing here: it carries the NCCL unique id of the drivers' own communicators
        # (driver.init_nccl) and reduces this script's timing numbers; the cycle's collective is native C++
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    # ---- parity gate of exactly the path that is timed below (N-rank device mesh + native all-reduce) ----
    parity = None
    if not args.no_parity_check:
        parity = parity_check(world, rank, local, dist)

    cycles = args.warmup + args.steps
    # one cycle more than is run, so that the state after the last timed cycle still has a time step ahead of it (the
    # drop-in measurement below sources that next cycle's photons)
    deck = make_deck(world, cycles + 1, photons_per_gpu=args.photons)
    tmp = tempfile.mkdtemp(prefix="branson_bench_")
    xml = deck.write(os.path.join(tmp, f"deck_rank{rank}.xml"))
    on_device = args.mesh == "device"
    d = driver.Driver(xml, n_groups=N_GROUPS, rank=rank, n_ranks=world, device=local,
                      algorithm=gpu.EVENT if args.algorithm == "event" else gpu.HISTORY, mesh_on_device=on_device)
    if world > 1:
        driver.init_nccl(d, dist)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def step(drv, dev_mesh):
        r = drv.cycle()
        if dev_mesh:
            drv.array("T_e")  # the cycle's result arrays live on the device in this mode: read one back every step
        return r

    for _ in range(args.warmup):
        step(d, on_device)
    launches_before = d.gpu_context().stats()["n_launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync_all()
    t0 = time.perf_counter()
    reps = [step(d, on_device) for _ in range(args.steps)]
    sync_all()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    # the same cycles through the array-exchanging path (host Mesh: f / op_a / op_s / E arrays up, tallies down)
    wall_host = None
    if on_device and not args.no_host_mesh_e2e:
        d.close()
        d2 = driver.Driver(xml, n_groups=N_GROUPS, rank=rank, n_ranks=world, device=local,
                           algorithm=gpu.EVENT if args.algorithm == "event" else gpu.HISTORY, mesh_on_device=False)
        if world > 1:
            driver.init_nccl(d2, dist)
        for _ in range(args.warmup):
            d2.cycle()
        sync_all()
        t0 = time.perf_counter()
        reps_host = [d2.cycle() for _ in range(args.steps)]
        sync_all()
        wall_host = time.perf_counter() - t0
        hist_host = sum(r["gpu"]["n_transported"] for r in reps_host)
        d = d2

    g = [r["gpu"] for r in reps]
    dev_ms = sum(x["ms_source"] + x["ms_transport"] + x["ms_census"] for x in g)
    tr_ms = sum(x["ms_transport"] for x in g)
    hist = sum(x["n_transported"] for x in g)
    abytes = sum(algorithmic_bytes(x) for x in g)
    launches = g[-1]["n_launches"] - launches_before  # kernels launched through the ctx inside the timed region
    vals = torch.tensor([dev_ms, wall, tr_ms, float(hist), float(abytes), wall_host or 0.0,
                         float(hist_host) if wall_host else 0.0], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        mx = vals.clone()
        torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX)
        sm = vals.clone()
        torch.distributed.all_reduce(sm, op=torch.distributed.ReduceOp.SUM)
        dev_ms_max, wall_max, tr_ms_max, wall_host_max = mx[0].item(), mx[1].item(), mx[2].item(), mx[5].item()
        hist_all, hist_host_all = sm[3].item(), sm[6].item()
    else:
        dev_ms_max, wall_max, tr_ms_max, hist_all = dev_ms, wall, tr_ms, float(hist)
        wall_host_max, hist_host_all = wall_host or 0.0, float(hist_host) if wall_host else 0.0

    if rank == 0:
        n_cells = int(d.param("n_cells"))
        peak, peak_src = measured_peak()
        achieved = abytes / (tr_ms * 1e-3) / 1e9  # rank 0's kernel: bytes per launch / its average duration
        h2d_host = 5 * n_cells * 8  # f, op_a, op_s, E_emission, E_source
        d2h_host = 2 * n_cells * 8 + 256
        # device mesh: (dt, step, total_E) in, the running sums and cycle statistics out, plus the T_e read-back
        h2d, d2h = (64, 8 * 16 + 256 + 8 * n_cells) if on_device else (h2d_host, d2h_host)
        line = {"metric": METRIC, "value": hist_all / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(world, args.photons, args.warmup, args.steps, args.ref_photons),
                "parity_checked": parity,
                "e2e": {"value": hist_all / wall_max, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * wall_max / args.steps,
                        "path": ("Driver.cycle() with the mesh physics on the device (bgpu_mesh_*): cell state resident "
                                 "in HBM, scalars in, sums + T_e array out every cycle" if on_device else
                                 "Driver.cycle() with the host Mesh: f/op_a/op_s/E arrays uploaded, tallies downloaded "
                                 "every cycle"),
                        "host_phase_ms_per_step": {k[2:]: 1e3 * sum(r[k] for r in reps) / args.steps for k in
                                                   ("t_calc_energy", "t_cell_upload", "t_source", "t_transport",
                                                    "t_allreduce", "t_tally_download", "t_update_T")}},
                "e2e_host_mesh": ({"value": hist_host_all / wall_host_max, "unit": UNIT,
                                   "h2d_bytes_per_step": h2d_host, "d2h_bytes_per_step": d2h_host,
                                   "ms_per_step": 1e3 * wall_host_max / args.steps,
                                   "path": "the same cycles with the host Mesh (bit-identical host physics; arrays cross "
                                           "PCIe both ways every cycle)"} if wall_host else None),
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": "k_transport_history", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                             "traffic": ncu_traffic(),
                             "algorithmic_bytes_per_launch": abytes / args.steps,
                             "bytes_per_history": abytes / hist, "kernel_ms_per_launch": tr_ms / args.steps,
                             "kernel_histories_per_s": hist / (tr_ms * 1e-3),
                             "events_per_history": sum(x["n_events"] for x in g) / hist,
                             "scatters_per_history": sum(x["n_scatters"] for x in g) / hist,
                             "note": "scattering-dominated cycles are alu-pipe bound (Threefry), not HBM bound: see "
                                     "compute_bound below, DESIGN.md section 5 and profiles/"},
                # what actually limits the kernel on this workload (ncu: alu + FP64 pipes 100 % of their shared issue
                # rate, DRAM 2 %): Threefry.  The reference's algorithm prescribes 1 draw per event + 4 per effective
                # scatter; on this deck (sigma_s = 0, faux-multigroup cells) the kernel evaluates 1 + 2 of them -- the
                # physical-vs-effective test draw is void when sigma_s = 0 (u > 0 always) and only a history's last
                # group draw is observable (evaluated once, when the history ends) -- measured against a kernel that
                # does nothing but Threefry
                "compute_bound": {"bound": "alu pipe (Threefry2x64-20 evaluations)",
                                  "achieved": sum(x["n_events"] + 2 * x["n_scatters"] for x in g) / (tr_ms * 1e-3) / 1e9,
                                  "peak": THREEFRY_PEAK_GDRAWS, "unit": "Gdraws/s",
                                  "frac": sum(x["n_events"] + 2 * x["n_scatters"] for x in g) / (tr_ms * 1e-3) / 1e9
                                  / THREEFRY_PEAK_GDRAWS,
                                  "prescribed_draws": sum(x["n_events"] + 4 * x["n_scatters"] for x in g) / (tr_ms * 1e-3) / 1e9,
                                  "peak_source": "tools/ubench/threefry_variants.cu on this pool's B200 "
                                                 "(profiles/threefry_ubench_r01.txt)"},
                "clocks": clocks,
                "conservation": {"max_rel_radiation_balance": max(
                    abs(r["rad_balance_exact"]) / (r["pre_census_E"] + r["emission_E"] + r["source_E"]) for r in reps)}}
        if world == 1 and not args.no_aos_dropin and (wall_host or not on_device):  # needs the host Mesh's current state
            try:
                line["e2e_aos_dropin"] = aos_dropin_block(d, N_GROUPS, local)
            except Exception as e:
                line["e2e_aos_dropin"] = {"value": None, "unit": UNIT, "path": f"failed: {type(e).__name__}: {e}"}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_block(args.ref_photons, args.warmup, args.steps)
            except Exception as e:  # the bench line must still be printed
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "reference",
                                        "sample": f"failed: {e}"}
    d.close()
    # BASELINE configs[3] and [4] at their named sizes, N > 1 only: extra keys of the same JSON line
    extra = None
    if world > 1 and not args.no_extra_configs:
        extra = {}
        for name, (deck_x, cyc_x) in extra_configs(world).items():
            try:
                res = run_extra_config(name, deck_x, cyc_x, world, rank, local, dist)
            except Exception as e:  # the bench line must still be printed
                res = {"failed": f"{type(e).__name__}: {e}"}
            extra[name] = res
    if rank == 0:
        if extra is not None:
            line["configs"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier(device_ids=[local])
        torch.distributed.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--photons", type=int, default=PHOTONS_PER_GPU, help="user photons per cycle per GPU")
    ap.add_argument("--algorithm", default="history", choices=["history", "event"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aos-dropin", action="store_true", help="skip the gpu_transport_photons drop-in measurement")
    ap.add_argument("--mesh", default="device", choices=["device", "host"],
                    help="where calculate_photon_energy / update_temperature run (e2e path)")
    ap.add_argument("--no-host-mesh-e2e", action="store_true", help="skip the second, host-mesh e2e measurement")
    ap.add_argument("--ref-photons", type=int, default=REF_SAMPLE_PHOTONS,
                    help="user photons per cycle of the CPU arms' bounded sample (--impl reference and cpu_baseline)")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the N-rank oracle check before timing")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="N > 1: skip BASELINE configs[3] (hohlraum_multi) and configs[4] (big_cube) after the main run")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return our_arm(args)


if __name__ == "__main__":
    sys.exit(main())
