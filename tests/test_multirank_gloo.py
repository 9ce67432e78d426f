"""N > 1 host-side logic on CPU: world_size-2 `gloo` process group.  Each rank evaluates its contiguous target slice
(the partition nb_slice_bounds defines, the one nb_naive_accel / nb_bh_accel use on the GPU) with the CPU oracle,
the slices are re-assembled with the same in-place padded all-gather layout the library uses with NCCL, and the
NCCL-unique-id style broadcast + max-over-ranks timing reduction of bench.py are exercised."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    nb = importlib.import_module("n-body-simulation_b200")
    import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    # 1. the 128-byte communicator id travels from rank 0 to everybody (bench.py does this for ncclUniqueId)
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert ids[0] == bytes(range(128))
    # 2. identical bodies on every rank without communication (counter-based generator)
    m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=5)
    digest = torch.tensor([float(x.sum()), float(vz.sum())], dtype=torch.float64)
    gathered = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, digest)
    assert all(torch.equal(g, digest) for g in gathered)
    # 3. slice -> oracle rows -> padded in-place all-gather (chunk = ceil(n / world))
    b, e = nb.slice_bounds(n, world, rank)
    chunk = -(-n // world)
    ax, ay, az = O.naive_accel(m, x, y, z, rows=(b, e), nthreads=1)
    full = []
    for a in (ax, ay, az):
        buf = torch.zeros(world * chunk, dtype=torch.float64)
        buf[b:e] = torch.from_numpy(a[b:e])
        parts = list(buf.view(world, chunk).unbind(0))
        dist.all_gather(parts, buf[rank * chunk:(rank + 1) * chunk].clone())
        full.append(torch.cat(parts)[:n].numpy())
    ref = O.naive_accel(m, x, y, z, nthreads=1)
    for f, r in zip(full, ref):
        assert np.array_equal(f, r)
    # 4. energy: sqrt-balanced target ranges + all-reduce equals the single-rank sum up to rounding
    jb = int(np.floor(n * np.sqrt(rank / world)))
    je = n if rank + 1 == world else int(np.floor(n * np.sqrt((rank + 1) / world)))
    _, ek, ep = O.energy(m, x, y, z, vx, vy, vz, per_body=True)
    part = torch.tensor([ek[jb:je].sum(), ep[jb:je].sum()], dtype=torch.float64)
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    tot = O.energy(m, x, y, z, vx, vy, vz)
    assert abs(part[0].item() - tot[0]) <= 1e-12 * abs(tot[0]) and abs(-part[1].item() - tot[1]) <= 1e-12 * abs(tot[1])
    # 5. timing reduction: the job time is the max over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    dist.barrier()
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 1001])
def test_world2_slices_allgather(tmp_path, nb, oracle, n):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
