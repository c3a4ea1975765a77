"""Multi-rank replicated mode: rank r of the C++ driver = replicated-mode rank r of the reference.

Two ways of running N ranks, both through the NATIVE communicator of the device contexts (csrc/comm_native.cuh; no
Python on the collective path):

* in one process, one thread per rank, all ranks on cuda:0 (`driver.init_local`: the in-process rank-ordered device
  sum) -- runs on a one-GPU box and covers exactly the path `bench.py --gpus N` times: the device mesh with
  k_mesh_redistribute (the low-energy-cell rule of reference src/mesh.h:291-315 with every rank's totals formed locally),
  ONE packed in-place all-reduce of {tallies, every rank's scalars} per cycle, and k_mesh_update_temperature consuming
  it on the stream (reference src/replicated_driver.h:56-59,91-104, src/imc_state.h:207-252);
* one process per GPU under torchrun (NCCL over NVLink; needs >= 2 CUDA devices, skipped otherwise).

Every rank's per-photon integers must equal the N-rank oracle's bit for bit every cycle; temperatures and tallies 1e-9;
the conservation sums (formed from the all-reduced tail) 1e-12.
"""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from branson_b200 import decks, driver, gpu

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_cycle(rank, d, view, sim, rep, cyc, n_ranks):
    post = view.download(gpu.LIST_WORK, counters=True)
    for k in ("cell", "group", "ctr", "descriptor", "counters"):
        assert np.array_equal(post[k], sim.get("post/" + k, rank)), (cyc, rank, k)
    for k in ("abs_E", "track_E", "T_e", "T_r"):
        want = sim.get(k)
        got = d.array(k)
        assert np.max(np.abs(got - want)) <= 1e-9 * np.max(np.abs(want)), (cyc, rank, k)
    total = rep["pre_census_E"] + rep["emission_E"] + rep["source_E"]
    assert abs(rep["rad_balance_exact"]) <= 1e-12 * total, rep["rad_balance_exact"]
    assert abs(rep["rad_conservation"]) <= 1e-12 * total, rep["rad_conservation"]
    assert rep["trans_particles"] == sum(int(sim.get("n_photons", r)[0]) for r in range(n_ranks))
    assert rep["census_size"] == sum(int(sim.get("n_census", r)[0]) for r in range(n_ranks))
    # the global sums every rank formed from the all-reduced tail against the oracle's rank-ordered sums
    for k in ("emission_E", "source_E", "pre_census_E", "post_census_E", "exit_E"):
        want = 0.0
        for r in range(n_ranks):
            want += float(sim.get(k, r)[0])
        assert abs(rep[k] - want) <= 1e-12 * max(abs(want), 1e-300), (cyc, rank, k, rep[k], want)
    assert abs(rep["global_source_energy"] - sim.get("global_source_energy")[0]) <= 1e-12 * rep["global_source_energy"]


CASES = [
    ("hohlraum_multi", lambda: decks.hohlraum_multi(photons=60000, t_stop=0.004, scale=5), 2),
    ("hohlraum_multi", lambda: decks.hohlraum_multi(photons=60000, t_stop=0.004, scale=5), 4),
    ("three_region", lambda: decks.simple_three_region(photons=30000, n_groups=30), 3),
]


@pytest.mark.parametrize("mesh", ["device", "host"])
@pytest.mark.parametrize("name,make,n_ranks", CASES, ids=[f"{c[0]}-r{c[2]}" for c in CASES])
def test_in_process_ranks_match_n_rank_oracle(name, make, n_ranks, mesh, tmp_path):
    from oracle import port
    deck = make().with_(dd_transport_type="REPLICATED")
    sim = port.OracleSim(deck, n_ranks=n_ranks)
    drivers = []
    for r in range(n_ranks):
        drivers.append(driver.Driver(deck.write(str(tmp_path / f"deck_{r}.xml")), n_groups=deck.n_groups, rank=r,
                                     n_ranks=n_ranks, device=0, validate=True, mesh_on_device=(mesh == "device")))
    driver.init_local(drivers)
    views = [d.gpu_context() for d in drivers]
    assert all(gpu.comm_info(v._h)["kind"] == gpu.COMM_LOCAL for v in views)
    cyc = 0
    while not sim.finished():
        cyc += 1
        sim.cycle(keep_photons=True)
        reps = driver.run_ranks(drivers, lambda r, d: d.cycle())
        for r, d in enumerate(drivers):
            _check_cycle(r, d, views[r], sim, reps[r], cyc, n_ranks)
        # every rank reports the same global sums, bit for bit (they all hold the same all-reduced tail)
        for k in ("emission_E", "source_E", "absorbed_E", "pre_census_E", "post_census_E", "exit_E", "pre_mat_E",
                  "post_mat_E", "rad_conservation", "trans_particles", "census_size", "global_source_energy"):
            assert all(rep[k] == reps[0][k] for rep in reps), k
    assert cyc >= 3 and all(d.finished() for d in drivers)
    if mesh == "device":
        # ONE collective per cycle: {abs_E, track_E}[n_cells] + tail[n_ranks][BGPU_RANK_SCALARS]
        for v in views:
            info = gpu.comm_info(v._h)
            assert info["calls"] == cyc, info
            assert info["bytes"] == cyc * 8 * (2 * deck.n_cells + n_ranks * 12), info
    for d in drivers:
        d.close()


def test_cli_two_ranks_in_one_process(tmp_path):
    """bin/branson --ranks 2: the reference's `mpirun -n 2 BRANSON deck.xml` in one process, one host thread per rank
    (rank r on GPU r % n_devices: in-process collectives on a one-GPU box, ncclCommInitAll + NCCL on a box with two); its
    printed photon counts equal the 2-rank oracle's"""
    from oracle import port
    deck = decks.hohlraum_multi(photons=40000, t_stop=0.003, scale=5).with_(dd_transport_type="REPLICATED")
    xml = deck.write(str(tmp_path / "deck.xml"))
    exe = os.path.join(ROOT, "branson_b200", "bin", "branson")
    out = subprocess.run([exe, xml, str(deck.n_groups), "--ranks", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    import re
    got = [int(x) for x in re.findall(r"Total Photons transported: (\d+)", out.stdout)]
    sim = port.OracleSim(deck, n_ranks=2)
    want = []
    while not sim.finished():
        sim.cycle(keep_photons=False)
        want.append(sum(int(sim.get("n_photons", r)[0]) for r in range(2)))
    assert got == want, (got, want)
    assert "collectives:" in out.stdout
    for x in re.findall(r"Radiation conservation: ([0-9.eE+-]+)", out.stdout):
        assert abs(float(x)) < 1e-10


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, os.environ["BRANSON_ROOT"])
    import torch
    import torch.distributed as dist
    from branson_b200 import decks, driver, gpu
    from oracle import port

    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")   # bootstrap only: carries the NCCL unique id; the collectives are native
    rank, world = dist.get_rank(), dist.get_world_size()
    deck = decks.hohlraum_multi(photons=60000, t_stop=0.004, scale=5).with_(dd_transport_type="REPLICATED")
    path = os.path.join(os.environ["BRANSON_TMP"], f"deck_{rank}.xml")
    deck.write(path)
    on_device = os.environ.get("BRANSON_MESH_ON_DEVICE") == "1"
    d = driver.Driver(path, n_groups=deck.n_groups, rank=rank, n_ranks=world, device=local, validate=True,
                      mesh_on_device=on_device)
    driver.init_nccl(d, dist)
    view = d.gpu_context()
    assert gpu.comm_info(view._h)["kind"] == gpu.COMM_NCCL
    sim = port.OracleSim(deck, n_ranks=world)
    cyc = 0
    while not sim.finished():
        cyc += 1
        sim.cycle(keep_photons=True)
        rep = d.cycle()
        post = view.download(gpu.LIST_WORK, counters=True)
        for k in ("cell", "group", "ctr", "descriptor", "counters"):
            assert np.array_equal(post[k], sim.get("post/" + k, rank)), (cyc, k)
        for k in ("abs_E", "track_E", "T_e", "T_r"):
            want = sim.get(k)
            got = d.array(k)
            assert np.max(np.abs(got - want)) <= 1e-9 * np.max(np.abs(want)), (cyc, k)
        total = rep["pre_census_E"] + rep["emission_E"] + rep["source_E"]
        assert abs(rep["rad_balance_exact"]) <= 1e-12 * total, rep["rad_balance_exact"]
        assert abs(rep["rad_conservation"]) <= 1e-12 * total, rep["rad_conservation"]
        assert rep["trans_particles"] == sum(int(sim.get("n_photons", r)[0]) for r in range(world))
        assert rep["census_size"] == sum(int(sim.get("n_census", r)[0]) for r in range(world))
    if on_device:
        assert gpu.comm_info(view._h)["calls"] == cyc
    d.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok after {cyc} cycles")
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mesh", ["host", "device"])
def test_two_gpus_match_two_rank_oracle(tmp_path, mesh):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices (the in-process test above covers the same path on one)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, BRANSON_ROOT=ROOT, BRANSON_TMP=str(tmp_path),
               BRANSON_MESH_ON_DEVICE="1" if mesh == "device" else "0")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)], env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout
