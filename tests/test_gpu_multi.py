"""Several ranks on real GPUs (each case skipped unless that many GPUs are visible): target-sharded ranks must reproduce
the single-GPU result bit for bit -- naive and Barnes-Hut (separate calls and the fused nb_advance, with the walk's peer
stores over IPC-mapped memory and, with NB_DISABLE_P2P=1, with the NCCL all-gather) -- and the sharded energy to 1e-12,
for body counts that the ranks do not divide."""
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
def same(u, v):
    return all(np.array_equal(a, b) for a, b in zip(u, v))
def state(c):
    return c.positions() + c.velocities() + c.accelerations()
want_p2p = os.environ.get("NB_DISABLE_P2P") is None
# third case: the persistent walk forced (walk_variant 50) so that the cost-weighted slices are in play on every world
# size; the accelerations must not depend on where the sorted order is cut
for n, gen, wv in ((20001, "plummer", 0), (700001, "uniform_sphere", 0), (600011, "plummer", 50)):
    m, x, y, z, vx, vy, vz = getattr(nb.generators, gen)(n, seed=3)
    ref = nb.Context(device=rank, theta=0.5)
    ref.set_bodies(m, x, y, z, vx, vy, vz)
    ctx = nb.Context(device=rank, theta=0.5, world_size=world, rank=rank, walk_variant=wv)
    ctx.comm_init(comm_id(), world, rank)
    ctx.set_bodies(m, x, y, z, vx, vy, vz)
    assert ctx.p2p_enabled() == want_p2p, "peer mapping: %%s (expected %%s)" %% (ctx.p2p_enabled(), want_p2p)
    if n < 100000:
        for c in (ref, ctx):
            c.naive_accel()
        a, b = ref.accelerations(), ctx.accelerations()
        # the source segmentation of the all-pairs kernel depends on N only: the same bits on one GPU and on several
        assert same(a, b), "naive sharded != single"
        e1, e2 = ref.energy(), ctx.energy()
        assert np.all(np.abs(e1 - e2) <= 1e-12 * np.abs(e1)), (e1, e2)
    # separate calls
    for c in (ref, ctx):
        c.bh_build(); c.bh_accel(); c.leapfrog_part1(0.01); c.bh_build(); c.bh_accel(); c.leapfrog_part2(0.01)
    assert same(state(ref), state(ctx)), "barnes-hut sharded != single (separate calls, n=%%d)" %% n
    # batches of steps: fused walk + integrator, results stored into every rank's arrays by the walk itself
    for k in (1, 2, 5):
        ref.advance("BarnesHut", 0.01, k); ctx.advance("BarnesHut", 0.01, k)
        assert same(state(ref), state(ctx)), "barnes-hut sharded != single (nb_advance %%d, n=%%d)" %% (k, n)
    # and separate calls again on the state the batches left behind
    for c in (ref, ctx):
        c.leapfrog_part1(0.01); c.bh_build(); c.bh_accel(); c.leapfrog_part2(0.01)
    assert same(state(ref), state(ctx)), "barnes-hut sharded != single (after the batches, n=%%d)" %% n
    if n < 100000:
        e1, e2 = ref.energy(), ctx.energy()
        assert np.all(np.abs(e1 - e2) <= 1e-12 * np.abs(e1)), (e1, e2)
    dist.barrier()
    ctx.close(); ref.close()
dist.barrier()
if rank == 0:
    print("MULTI_OK p2p=%%s world=%%d" %% (want_p2p, world))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("p2p", [True, False])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks_match_single_gpu(tmp_path, world, p2p):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ)
    if not p2p:
        env["NB_DISABLE_P2P"] = "1"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(29533 + world + (0 if p2p else 10)), str(script)],
                       capture_output=True, text=True, timeout=900, env=env)
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
