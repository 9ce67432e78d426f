"""Two-rank run on real GPUs (skipped unless >= 2 GPUs are visible): accelerations computed by target-sharded ranks +
NCCL all-gather must equal the single-GPU result (Barnes-Hut: bit for bit; naive: to 1e-13, its source-range
segmentation depends on the slice size)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import importlib, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, %r)
nb = importlib.import_module("n-body-simulation_b200")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
def comm_id():
    ids = [nb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    return ids[0]
n = 20001
m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=3)
ref = nb.Context(device=rank, theta=0.5)
ref.set_bodies(m, x, y, z, vx, vy, vz)
ctx = nb.Context(device=rank, theta=0.5, world_size=world, rank=rank)
ctx.comm_init(comm_id(), world, rank)
ctx.set_bodies(m, x, y, z, vx, vy, vz)
for c in (ref, ctx):
    c.naive_accel()
a, b = ref.accelerations(), ctx.accelerations()
# the source range may be split into a different number of segments for a slice: rounding-level differences only
assert all(np.allclose(u, v, rtol=1e-13, atol=0) for u, v in zip(a, b)), "naive sharded != single"
for c in (ref, ctx):
    c.bh_build(); c.bh_accel(); c.leapfrog_part1(0.01); c.bh_build(); c.bh_accel(); c.leapfrog_part2(0.01)
a, b = ref.accelerations() + ref.positions() + ref.velocities(), ctx.accelerations() + ctx.positions() + ctx.velocities()
assert all(np.array_equal(u, v) for u, v in zip(a, b)), "barnes-hut sharded != single"
e1, e2 = ref.energy(), ctx.energy()
assert np.all(np.abs(e1 - e2) <= 1e-12 * np.abs(e1)), (e1, e2)
dist.barrier()
if rank == 0:
    print("MULTI_OK")
dist.destroy_process_group()
'''


def test_two_ranks_match_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "MULTI_OK" in r.stdout


def test_host_executable_two_ranks(tmp_path):
    """The C++ driver under a torchrun-style launcher: one process per GPU, file rendezvous for the NCCL id, rank 0
    writes the output; lastState.csv must equal the single-process run byte for byte."""
    import glob
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "n-body-simulation_b200", "N_Body_Simulation")
    fixture = os.path.join(ROOT, "tests", "golden", "solar_178.csv")
    common = ["--file=" + fixture, "--dt=1h", "--t_end=3d", "--vs=1d", "--algorithm=BarnesHut", "--theta=0.5", "--energy=true"]
    r1 = subprocess.run([exe] + common + ["--vs_dir=" + str(tmp_path / "one")], capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0, r1.stderr
    r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                         "--master-addr", "127.0.0.1", "--master-port", "29544", "--no-python", exe] + common +
                        ["--vs_dir=" + str(tmp_path / "two")], capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stdout[-1500:] + r2.stderr[-3000:]
    d1 = glob.glob(str(tmp_path / "one" / "*"))
    d2 = glob.glob(str(tmp_path / "two" / "*"))
    assert len(d1) == 1 and len(d2) == 1          # only rank 0 writes
    for name in ("lastState.csv", "simulation_step3.vtp"):
        assert open(os.path.join(d1[0], name)).read() == open(os.path.join(d2[0], name)).read(), name
