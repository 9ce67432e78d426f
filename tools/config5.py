"""BASELINE configs[4]: Barnes-Hut theta = 0.2, N = 2^26 bodies, --energy, on the GPUs torchrun gives us (8 x B200 in the
record under profiles/; any count works, one GPU without --energy to obtain the checksum the sharded run must reproduce).

What runs is what the reference's time loop does around one visualised step (BarnesHutAlgorithm.cpp:149-276): initial
force evaluation, leapfrog part 1, tree build, traversal, leapfrog part 2, then nBodyAlgorithm::computeEnergy
(nBodyAlgorithm.cpp:11-86) sharded over the ranks (sqrt-balanced target ranges + ncclAllReduce).  Rank 0 prints one
JSON object: step time, phase times, energy time and values, device memory in use, checksums of the state.

    torchrun --nproc-per-node 8 tools/config5.py [--n 67108864] [--theta 0.2] [--energy 1] [--gen uniform_sphere]
"""
import argparse, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 26)
ap.add_argument("--theta", type=float, default=0.2)
ap.add_argument("--energy", type=int, default=1)
ap.add_argument("--gen", default="uniform_sphere")
ap.add_argument("--out", default=None)
args = ap.parse_args()
import torch
nb = importlib.import_module("n-body-simulation_b200")
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
t0 = time.perf_counter()
if args.gen == "uniform_sphere":   # in chunks: 7 arrays of N doubles and nothing else (8 ranks share the host's memory)
    arrs = [np.empty(args.n) for _ in range(7)]
    step_ids = 1 << 22
    for first in range(0, args.n, step_ids):
        k = min(step_ids, args.n - first)
        part = nb.generators.uniform_sphere(k, seed=1, velocity_scale=0.3, first_id=first, n_total=args.n)
        for dst, src in zip(arrs, part):
            dst[first:first + k] = src
    m, x, y, z, vx, vy, vz = arrs
else:
    m, x, y, z, vx, vy, vz = getattr(nb.generators, args.gen)(args.n, seed=1)
t_gen = time.perf_counter() - t0
total_mass = float(m.sum())
ctx = nb.Context(device=local, theta=args.theta, wg_size_barnes_hut=128, world_size=world, rank=rank)
if world > 1:
    ids = [nb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], world, rank)
t0 = time.perf_counter()
ctx.set_bodies(m, x, y, z, vx, vy, vz)
t_upload = time.perf_counter() - t0
del m, x, y, z, vx, vy, vz
if args.gen == "uniform_sphere":
    del arrs
ctx.enable_timers(True)
dt = 1e-3


def timed(fn):
    ctx.synchronize()
    if world > 1:
        dist.barrier()
    t = time.perf_counter()
    fn()
    ctx.synchronize()
    return time.perf_counter() - t


def step():
    ctx.leapfrog_part1(dt); ctx.bh_build(); ctx.bh_accel(); ctx.leapfrog_part2(dt)


t_init = timed(lambda: (ctx.bh_build(), ctx.bh_accel()))
ph_init = ctx.timers()
t_step = timed(step)
ph_step = ctx.timers()
info = ctx.bh_tree_info()
free_b, total_b = torch.cuda.mem_get_info(dev)
energy = None
t_energy = None
if args.energy:
    e = [None]
    t_energy = timed(lambda: e.__setitem__(0, ctx.energy()))
    energy = [float(v) for v in e[0]]
def abs_sum(get):
    parts = get()
    return float(sum(np.abs(c).sum() for c in parts))


sum_a, sum_x, sum_v = abs_sum(ctx.accelerations), abs_sum(ctx.positions), abs_sum(ctx.velocities)
out = {
    "config": "Barnes-Hut theta=%g, %s N=%d, %d GPU(s), energy=%d (BASELINE configs[4])" % (args.theta, args.gen, args.n, world, args.energy),
    "p2p": ctx.p2p_enabled() if world > 1 else None,
    "seconds": {"generate_bodies_host": t_gen, "upload": t_upload, "initial_forces": t_init, "visualised_step": t_step, "energy": t_energy},
    "phases_ms_step": {k: round(val, 3) for k, val in ph_step.items() if val},
    "phases_ms_initial": {k: round(val, 3) for k, val in ph_init.items() if val},
    "tree": {"internal_nodes": int(info.num_internal), "canonical_nodes": int(info.num_nodes_canonical), "max_depth": int(info.max_depth)},
    "device_memory_in_use_gb": (total_b - free_b) / 2 ** 30,
    "energy": energy,
    "checksum": {"sum_abs_a": sum_a, "sum_abs_x": sum_x, "sum_abs_v": sum_v},
    "host_memory": open("/proc/meminfo").read().split("\n")[0:3],
}
if energy is not None and args.gen == "uniform_sphere":
    G = ctx.cfg.G
    M = total_mass
    analytic = -0.6 * G * M * M / 1.0          # homogeneous sphere of radius 1 AU (the positions moved by one tiny step)
    out["energy_check"] = {"potential_analytic_continuum": analytic, "relative_deviation": (energy[1] - analytic) / abs(analytic),
                           "pairs": args.n * (args.n - 1) / 2.0, "pairs_per_s": args.n * (args.n - 1) / 2.0 / t_energy}
ctx.close()
if rank == 0:
    s = json.dumps(out)
    print(s, flush=True)
    if args.out:
        open(args.out, "w").write(s + "\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
