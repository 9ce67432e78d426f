#!/usr/bin/env python
"""Benchmark of the gravity hot path on B200 (contract: see the task's bench.py section and SURVEY.md 8d).

Headline workload (BASELINE.json configs[1]): naive all-pairs fp64, synthetic Plummer sphere, N = 2^20 bodies.
A "step" is one leapfrog step of the whole system: kick-drift, one all-pairs force evaluation (N^2 body
interactions, self term included as in the reference), closing kick.  metric = body-interactions/s over the whole job.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n BODIES] [--no-bh]

  * our arm: the product library through its C ABI (ctypes), inputs resident in HBM for `value`; `e2e` repeats the
    measurement through nb_op_naive_accelerations with pinned HOST buffers (H2D of m,x,y,z and D2H of ax,ay,az inside
    the timed region).  `roofline` is the FP64 pipe: achieved = 21 flop x N^2 / kernel time (CUDA events on the
    launching stream), peak = a DFMA-chain microbenchmark run in the same process (MEASURED_PEAKS.json has no fp64
    figure).  `bh` adds the secondary metric (Barnes-Hut steps/s, uniform sphere, theta = 0.5) with its HBM roofline.
  * --impl reference: the reference's own NaiveAlgorithm::computeAccelerations_opt_N (unmodified source compiled with
    g++/OpenMP into oracle/_ref) on all host threads, over a bounded sample of the workload.
Multi-GPU: launched by torchrun with one rank per GPU; targets are sharded by contiguous ranges, accelerations are
all-gathered by the library's NCCL communicator (strong scaling: total work fixed).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_INTERACTION = 21.0          # SURVEY 8d (NaiveAlgorithm.cpp:332-342)
BH_BYTES_PER_VISIT = 40.0            # com xyz + mass (32 B) + skip/meta (8 B)
BH_BYTES_PER_BODY = 48.0             # position in, acceleration out
NAIVE_TILE = 256                     # --block_size used for the benchmark (shared-memory tile length)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures summarised under profiles/
# (naive: N = 2^20, profiles/naive_accel_r01b.txt; Barnes-Hut walk: N = 2^24, profiles/bh_traverse_r01c.txt)
NCU_TRAFFIC_BYTES = {("naive", 1 << 20): 123.167744e6 + 55.533824e6, ("bh", 1 << 24): 2.557424e9 + 400.486656e6}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1 << 20, help="bodies of the naive workload (default 2^20)")
    ap.add_argument("--bh-n", type=int, default=1 << 24, help="bodies of the Barnes-Hut workload (default 2^24)")
    ap.add_argument("--no-bh", action="store_true", help="skip the secondary Barnes-Hut measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(power)) if power else None)
        return out


def flush_l2(torch, dev):
    """Write a buffer larger than the 126 MB L2 between timed iterations."""
    if not hasattr(flush_l2, "buf"):
        flush_l2.buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_l2.buf.fill_(1)
    torch.cuda.synchronize(dev)


# ---------------------------------------------------------------------------------------------------------------------
def host_cores():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would hide them)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class CpuReference:
    """The reference's own CPU implementation of the naive path, timed on the host cores.

    kind "reference": oracle/_ref/libnbody_ref.so, the reference's unmodified NaiveAlgorithm::computeAccelerations_opt_N
    compiled with g++/OpenMP (oracle/Makefile).  kind "port": the oracle restatement, only if oracle/_ref is missing.
    The reference evaluates all N^2 pairs of the bodies it is given, so the bounded sample is the first n_s bodies of
    the workload's N = 2^20 Plummer set (the CPU rate in interactions/s does not depend on N)."""

    def __init__(self, nb, n):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        self.m, self.x, self.y, self.z, *_ = nb.generators.plummer(n, seed=1)
        self.n = n
        self.cores = host_cores()
        try:
            import refimpl as R
            if not R.available():
                raise RuntimeError("oracle/_ref not built")
            R.lib()
            R.set_threads(self.cores)
            self.R, self.kind = R, "reference"
        except Exception as e:  # noqa: BLE001 - fall back to the restatement, and say so
            import oracle as O
            self.O, self.kind, self.why = O, "port", repr(e)
        self.stage = 0

    def evaluate(self, ns, stage=None):
        m, x, y, z = (a[:ns] for a in (self.m, self.x, self.y, self.z))
        t0 = time.perf_counter()
        if self.kind == "reference":
            self.R.naive_accel(m, x, y, z, opt_stage=self.stage if stage is None else stage, block_size=64)
        else:
            self.O.naive_accel(m, x, y, z, nthreads=self.cores)
        return time.perf_counter() - t0

    def calibrate(self, seconds):
        """Pick the faster of the reference's opt stages 0 and 2 and a sample size worth ~`seconds` of CPU work."""
        probe = min(self.n, 16384)
        t = {0: self.evaluate(probe, 0)}
        if self.kind == "reference":
            t[2] = self.evaluate(probe, 2)
        self.stage = min(t, key=t.get)
        rate = probe * float(probe) / max(t[self.stage], 1e-6)
        ns = probe
        while ns * 2 <= self.n and (2.0 * ns) ** 2 / rate <= seconds:
            ns *= 2
        return ns

    def describe(self, ns, dt=None):
        what = ("reference NaiveAlgorithm::computeAccelerations_opt_%d (unmodified source, g++ -O3 -fopenmp via oracle/_ref)"
                % self.stage) if self.kind == "reference" else "oracle restatement of the reference loop (oracle/_ref missing)"
        s = "first %d bodies of the N=%d Plummer set, all %.3g pairs, %s" % (ns, self.n, ns * float(ns), what)
        return s + (", %.1f s" % dt if dt is not None else "")


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    nb = importlib.import_module("n-body-simulation_b200")
    n = args.n
    cpu = CpuReference(nb, n)
    ns = cpu.calibrate(seconds=6.0)
    for _ in range(min(args.warmup, 1)):
        cpu.evaluate(ns)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.evaluate(ns)
    dt = time.perf_counter() - t0
    value = ns * float(ns) * args.steps / dt
    line = {
        "impl": "reference", "metric": "body-interactions/s", "value": value, "unit": "interactions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "naive all-pairs fp64, Plummer sphere N=%d (BASELINE configs[1])" % n,
                   "note": "CPU arm: each step is one all-pairs evaluation over a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": cpu.cores, "kind": cpu.kind,
                         "sample": cpu.describe(ns)},
        "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline(nb, n, seconds=12.0):
    cpu = CpuReference(nb, n)
    ns = cpu.calibrate(seconds)
    dt = cpu.evaluate(ns)
    return {"value": ns * float(ns) / dt, "unit": "interactions/s", "cores": cpu.cores, "kind": cpu.kind,
            "sample": cpu.describe(ns, dt)}


# ---------------------------------------------------------------------------------------------------------------------
def fresh_comm_id(nb, dist, rank, world):
    """A new NCCL unique id per communicator (an id cannot be reused), created by rank 0 and broadcast by the host."""
    if world == 1:
        return None
    ids = [nb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    return ids[0]


def measure_bh(nb, torch, dist, args, rank, local_rank, world, dev):
    """Secondary metric: full Barnes-Hut steps/s (kick-drift, AABB, keys+sort, build, COM, traversal, kick)."""
    n = args.bh_n
    theta = 0.5
    m, x, y, z, vx, vy, vz = nb.generators.uniform_sphere(n, seed=1, velocity_scale=0.3)
    ctx = nb.Context(device=local_rank, theta=theta, wg_size_barnes_hut=128, world_size=world, rank=rank)
    if world > 1:
        ctx.comm_init(fresh_comm_id(nb, dist, rank, world), world, rank)
    ctx.set_bodies(m, x, y, z, vx, vy, vz)
    dt = 1e-3  # days: bodies move, the tree changes every step
    ctx.bh_build(); ctx.bh_accel(); ctx.synchronize()

    def step():
        ctx.leapfrog_part1(dt)
        ctx.bh_build()
        ctx.bh_accel()
        ctx.leapfrog_part2(dt)

    for _ in range(max(args.warmup, 1)):
        step()
    ctx.synchronize()
    ctx.enable_timers(True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    l0 = ctx.launch_count()
    ctx.event_record(0)
    for _ in range(args.steps):
        step()
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    ctx.synchronize()
    launches = ctx.launch_count() - l0
    timers = ctx.timers()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # visits for the roofline: one extra traversal with counters on (outside the timed region)
    ctx.bh_enable_stats(True)
    ctx.bh_build(); ctx.bh_accel()
    visits, accepts = ctx.bh_stats()
    info = ctx.bh_tree_info()
    if world > 1:
        t = torch.tensor([float(visits), float(accepts)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        visits, accepts = float(t[0].item()), float(t[1].item())
    t_trav = timers["Acceleration Kernel Time"] * 1e-3
    b0, b1 = nb.slice_bounds(n, world, rank)
    # per-rank traversal: its share of visits (approx. visits/world) + its bodies
    alg_bytes = BH_BYTES_PER_VISIT * visits / world + BH_BYTES_PER_BODY * (b1 - b0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / t_trav / 1e9 if t_trav > 0 else 0.0
    ctx.close()
    return {
        "metric": "Barnes-Hut steps/s", "value": args.steps / (ms * 1e-3), "unit": "steps/s", "ms_per_step": ms / args.steps,
        "config": {"workload": "Barnes-Hut theta=0.5, uniform sphere N=%d, full step (BASELINE configs[3])" % n,
                   "inputs": "larger than L2", "wg_size_barnes_hut": 128},
        "phases_ms": {k: round(v, 4) for k, v in timers.items() if v},
        "visits_per_body": visits / n, "accepts_per_body": accepts / n, "max_depth": int(info.max_depth),
        "internal_nodes_per_body": info.num_internal / n,
        "roofline": {"bound": "hbm", "kernel": "bh_traverse_iw_kernel", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                     "frac": achieved / hbm, "traffic": NCU_TRAFFIC_BYTES.get(("bh", n)) if world == 1 else None,
                     "peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
                     "note": "algorithmic bytes = 40 B x non-empty visits + 48 B x bodies; warp-uniform node loads are "
                             "served from L1/L2, so achieved can exceed the HBM peak"},
        "gpu_launches": int(launches),
    }


def run_ours(args, rank, local_rank, world):
    import torch
    nb = importlib.import_module("n-body-simulation_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the product has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n = args.n
    m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=1)
    ctx = nb.Context(device=local_rank, block_size=NAIVE_TILE, world_size=world, rank=rank)
    if world > 1:
        ctx.comm_init(fresh_comm_id(nb, dist, rank, world), world, rank)
    ctx.set_bodies(m, x, y, z, vx, vy, vz)
    dt = 1.0 / 24.0

    def step():
        ctx.leapfrog_part1(dt)
        ctx.naive_accel()
        ctx.leapfrog_part2(dt)

    ctx.naive_accel()
    for _ in range(args.warmup):
        step()
    ctx.synchronize()
    fp64_peak = ctx.measure_fp64_peak()

    ctx.enable_timers(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    l0 = ctx.launch_count()
    kernel_ms = []
    total_ms = 0.0
    for _ in range(args.steps):
        flush_l2(torch, dev)  # between timed iterations (inputs are 50 MB < L2)
        ctx.event_record(0)
        step()
        ctx.event_record(1)
        total_ms += ctx.event_elapsed_ms(0, 1)
        kernel_ms.append(ctx.timers()["Acceleration Kernel Time"])
    ctx.synchronize()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else {}
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = float(n) * n * args.steps / (total_ms * 1e-3)

    # ---- e2e: host buffers through the operator-form C-ABI call, copies inside the timed region --------------------
    pinned = [torch.from_numpy(a.copy()).pin_memory() for a in (m, x, y, z)]
    outs = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
    pin_np = [t.numpy() for t in pinned]
    out_np = [t.numpy() for t in outs]
    ctx.op_naive_accelerations(*pin_np, out=out_np)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.op_naive_accelerations(*pin_np, out=out_np)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = float(n) * n * args.steps / e2e_s
    a_norm = float(np.sqrt(out_np[0] ** 2 + out_np[1] ** 2 + out_np[2] ** 2).sum())  # the step's result, read on the host

    # ---- roofline of the dominant kernel (per rank: its slice of the targets) ------------------------------------------
    b0, b1 = nb.slice_bounds(n, world, rank)
    k_ms = float(np.mean(kernel_ms))
    achieved_tf = FLOP_PER_INTERACTION * (b1 - b0) * float(n) / (k_ms * 1e-3) / 1e12
    roofline = {"bound": "fp64", "kernel": "naive_accel_kernel", "achieved": achieved_tf, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak if fp64_peak else None,
                "traffic": NCU_TRAFFIC_BYTES.get(("naive", n)) if world == 1 else None,
                "peak_source": "DFMA-chain microbenchmark in this run (MEASURED_PEAKS.json has no fp64 figure); "
                               "nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2",
                "algorithmic_flop_per_interaction": FLOP_PER_INTERACTION, "kernel_ms": k_ms,
                "dp_instructions_per_interaction": 15}

    bh = None
    if not args.no_bh:
        ctx.close()
        try:
            bh = measure_bh(nb, torch, dist, args, rank, local_rank, world, dev)
        except Exception as e:  # the headline line must survive a failure of the secondary metric
            bh = {"error": repr(e)}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_baseline(nb, n)
        except Exception as e:
            cpu = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": "body-interactions/s", "value": value, "unit": "interactions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "naive all-pairs fp64, Plummer sphere N=%d, leapfrog step (BASELINE configs[1])" % n,
                       "block_size": NAIVE_TILE, "parallelism": "targets sharded x%d, NCCL all-gather of accelerations" % world,
                       "l2": "flushed between timed iterations"},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "interactions/s", "h2d_bytes_per_step": 4 * 8 * n,
                    "d2h_bytes_per_step": 3 * 8 * n, "sum_abs_a": a_norm},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "bh": bh,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank, local_rank, world = dist_env()
    # stdout must carry exactly ONE JSON line: libraries (NCCL / c10d print "NCCL version ..." on stdout) are diverted to
    # stderr for the whole run and the line is written to the saved descriptor at the end
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
