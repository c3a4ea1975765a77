"""Multi-GPU replicated mode: one process per GPU (NCCL), rank r = replicated-mode rank r of the reference.

Needs >= 2 CUDA devices (skipped otherwise; the single-GPU suite covers the same rank partitioning with n contexts on
one device in tests/test_gpu_parity.py, and the host partitioning runs under gloo in tests/test_multirank_gloo.py).
Each rank steps the C++ driver with a torch.distributed communicator; the tally all-reduce runs in place on the device
buffer over NCCL.  Rank 0 compares temperatures, tallies and conservation with the n-rank oracle every cycle; every rank
compares its own per-photon integers.
"""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

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
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    deck = decks.hohlraum_multi(photons=60000, t_stop=0.004, scale=5)
    path = os.path.join(os.environ["BRANSON_TMP"], f"deck_{rank}.xml")
    deck.write(path)
    comm = driver.TorchComm(f"cuda:{local}")
    on_device = os.environ.get("BRANSON_MESH_ON_DEVICE") == "1"
    d = driver.Driver(path, n_groups=deck.n_groups, rank=rank, n_ranks=world, device=local, validate=True, comm=comm,
                      mesh_on_device=on_device)
    view = d.gpu_context()
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
    assert comm.device_allreduce_bytes == cyc * 16 * deck.n_cells
    d.close()
    dist.barrier(device_ids=[local])
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
        pytest.skip("needs 2 CUDA devices")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    # mesh = device: calculate_photon_energy (with the rank redistribution) and update_temperature run on each GPU
    env = dict(os.environ, BRANSON_ROOT=ROOT, BRANSON_TMP=str(tmp_path),
               BRANSON_MESH_ON_DEVICE="1" if mesh == "device" else "0")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)], env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout
