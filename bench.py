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
    figure; the pipe rate at the sampled clock is the stricter denominator and the one `frac` uses).
    `bh` carries the second half of BASELINE's metric: Barnes-Hut steps/s (uniform sphere N = 2^24, theta = 0.5,
    configs[3]) through nb_advance, with phase times, a binding roofline for the walk (FP64 lane-operations against the
    DP pipe), `e2e` through nb_op_barnes_hut_accelerations with pinned host buffers, a checksum, and an in-run parity
    gate (sampled bodies against the CPU oracle: identical visit counts, accelerations <= 1e-10).  At N = 1 it also
    runs configs[2] (Plummer N = 2^20, theta = 0.5) on the GPU with the reference's own Barnes-Hut step timed beside it
    (`cpu_baseline`), and configs[0] (solar system, 8760 steps) as wall times of the two executables.
  * --impl reference: the reference's own NaiveAlgorithm::computeAccelerations_opt_N (unmodified source compiled with
    g++/OpenMP into oracle/_ref) on all host threads, over a bounded sample of the workload; its `bh` block times the
    reference's buildOctree + BarnesHutAlgorithm::computeAccelerations on configs[2].
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
BH_DP_PER_ACCEPT = 16.0              # fp64 instructions of an accepted node (3 DADD, 3 DFMA, 7 for m*d^-3, 3 DFMA)
BH_DP_PER_OPEN = 6.0                 # an opened node stops after the distance (3 DADD, 3 DFMA)
FP64_LANES_PER_SM = 64.0             # DP pipe: 64 lanes per SM and clock (one DFMA per lane)
NAIVE_TILE = 256                     # --block_size used for the benchmark (shared-memory tile length)
# dram__bytes_read.sum + dram__bytes_write.sum per launch, NOT measured in this run: from the `ncu --set full` captures
# summarised under profiles/ (naive: N = 2^20, profiles/naive_accel_r02.txt -- 444 MB of the writes are the per-segment
# partial sums; walk: N = 2^24, acceleration-only form, profiles/bh_traverse_r02f.txt)
NCU_TRAFFIC_BYTES = {("naive", 1 << 20): 168.282112e6 + 443.659520e6, ("bh", 1 << 24): 2.479455e9 + 398.411776e6}
NCU_TRAFFIC_SOURCE = {"naive": "profiles/naive_accel_r02.txt", "bh": "profiles/bh_traverse_r02f.txt"}
PARITY_TOL = 1e-10


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1 << 20, help="bodies of the naive workload (default 2^20)")
    ap.add_argument("--bh-n", type=int, default=1 << 24, help="bodies of the Barnes-Hut workload (default 2^24)")
    ap.add_argument("--no-bh", action="store_true", help="skip the Barnes-Hut measurements")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity gate of the Barnes-Hut line")
    ap.add_argument("--no-plummer", action="store_true", help="skip the Plummer N=2^24 Barnes-Hut scaling line")
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

    def __init__(self, gpu_index, interval_ms=200):
        self.gpu = gpu_index
        self.interval_ms = interval_ms
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.interval_ms)],
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
    if not args.no_bh:   # second half of the metric: the reference's Barnes-Hut force evaluation on configs[2]
        try:
            mm, xx, yy, zz, *_ = nb.generators.plummer(1 << 20, seed=1)
            b = cpu_bh_baseline(nb, mm, xx, yy, zz, 0.5)
            line["bh_config3"] = {"metric": "Barnes-Hut steps/s", "value": b["value"], "unit": "steps/s",
                                  "config": {"workload": "Barnes-Hut theta=0.5, Plummer sphere N=%d (BASELINE configs[2])" % (1 << 20)},
                                  "cpu_baseline": b}
        except Exception as e:  # noqa: BLE001
            line["bh_config3"] = {"error": repr(e)}
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


def walk_roofline(visits, accepts, bodies_share, t_walk_s, sm_count, sm_mhz, n, world):
    """Binding bound of the Barnes-Hut walk: fp64 lane-operations through the DP pipe.
    Algorithmic work = 16 fp64 operations per accepted node + 6 per opened node, per body that visits it; the pipe retires
    64 lane-operations per SM and clock.  (The HBM accounting of SURVEY 8d, 40 B per visit, does not bind: a warp's 32
    lanes share one broadcast load and the bytes come from L1/L2 -- it is kept as `algorithmic_gbs`.)"""
    opens = visits - accepts
    lane_ops = (BH_DP_PER_ACCEPT * accepts + BH_DP_PER_OPEN * opens) * bodies_share
    peak = sm_count * FP64_LANES_PER_SM * sm_mhz * 1e6 / 1e9            # G lane-ops/s
    achieved = lane_ops / t_walk_s / 1e9 if t_walk_s > 0 else 0.0
    alg_bytes = (BH_BYTES_PER_VISIT * visits + BH_BYTES_PER_BODY) * bodies_share
    return {"bound": "fp64", "kernel": "bh_traverse_iw_kernel", "achieved": achieved, "peak": peak,
            "unit": "G fp64 lane-ops/s", "frac": achieved / peak if peak else None,
            "traffic": NCU_TRAFFIC_BYTES.get(("bh", n)) if world == 1 else None,
            "traffic_source": "ncu capture %s (not measured in this run)" % NCU_TRAFFIC_SOURCE["bh"],
            "peak_source": "%d SMs x 64 fp64 lanes x %.0f MHz (SM clock sampled under load in this run)" % (sm_count, sm_mhz),
            "algorithmic": "16 fp64 ops per accepted node + 6 per opened node, per visiting body",
            "kernel_ms": t_walk_s * 1e3, "algorithmic_gbs": alg_bytes / t_walk_s / 1e9 if t_walk_s > 0 else 0.0,
            "note": "algorithmic_gbs (40 B x visits + 48 B x bodies, SURVEY 8d) exceeds the HBM peak because warp-uniform "
                    "node loads are L1/L2 broadcasts; the DP pipe is the bound that binds"}


def bh_parity_gate(nb, ctx, m, theta, world, samples=256):
    """Sampled parity of the state the timed steps left behind, on rank 0: the CPU oracle (oracle/, the restatement of
    BarnesHutAlgorithm.cpp:319-393 pinned against the reference) builds the canonical tree of the same positions and
    walks the sampled bodies; visit counts must be identical and accelerations within 1e-10.  Every rank takes part in
    the instrumented GPU evaluation; only rank 0 runs the oracle."""
    n = m.shape[0]
    px, py, pz = ctx.positions()
    ctx.bh_enable_stats(True)
    ctx.bh_build(); ctx.bh_accel()
    got = ctx.accelerations()
    info = ctx.bh_tree_info()
    tv, ta, per_body = ctx.bh_stats(per_body=True)
    ctx.bh_enable_stats(False)
    out = {"visits": tv, "accepts": ta, "max_depth": int(info.max_depth), "num_internal": int(info.num_internal)}
    if ctx.cfg.rank != 0:
        return out
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    t0 = time.perf_counter()
    tree = O.Tree(m, px, py, pz, storage_param=5, insertion_order="morton")
    b0, b1 = nb.slice_bounds(n, world, 0)
    rng = np.random.default_rng(11)
    # per-body visit counters exist for the slots this rank walked; sample among its bodies (sorted slots [b0, b1))
    mine = np.nonzero(per_body)[0] if world > 1 else np.arange(n)
    ids = np.unique(mine[rng.integers(0, mine.size, samples)]).astype(np.uint32)
    ax, ay, az, st = tree.accel_sample(theta, ids, stats=True)
    g = np.stack([a[ids] for a in got], 1)
    r = np.stack([ax, ay, az], 1)
    err = float((np.linalg.norm(g - r, axis=1) / np.linalg.norm(r, axis=1)).max())
    visits_equal = bool(np.array_equal(per_body[ids], st[:, 1].astype(np.uint32)))
    nodes_equal = bool(info.num_nodes_canonical == tree.num_nodes and info.max_depth == tree.max_depth)
    out["parity"] = {"samples": int(ids.size), "visit_counts_identical": visits_equal, "tree_nodes_and_depth_identical": nodes_equal,
                     "max_rel_err": err, "tol": PARITY_TOL, "ok": bool(visits_equal and nodes_equal and err <= PARITY_TOL),
                     "oracle": "oracle/nbody_oracle.cpp: canonical tree of the same positions + BarnesHutAlgorithm.cpp:319-393 on the sampled bodies",
                     "oracle_seconds": time.perf_counter() - t0}
    return out


_BODIES = {}


def time_advance(ctx, torch, dist, dev, world, dt, steps, warmup):
    """steps/s of nb_advance (the time loop's batch of steps): CUDA events on the library's stream, max over ranks."""
    ctx.advance("BarnesHut", dt, max(warmup, 1))
    ctx.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    l0 = ctx.launch_count()
    ctx.event_record(0)
    ctx.advance("BarnesHut", dt, steps)
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    ctx.synchronize()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, ctx.launch_count() - l0


def measure_bh(nb, torch, dist, args, rank, local_rank, world, dev, sampler, gen="uniform_sphere", full=True, static_slices=0):
    """Second half of BASELINE's metric: full Barnes-Hut steps/s (kick-drift, AABB, keys + sort, build, COM, traversal,
    kick), uniform sphere N = 2^24, theta = 0.5 (configs[3]), targets sharded over the ranks.  full=False: the light
    form used for the Plummer scaling line (no e2e, no parity gate)."""
    n = args.bh_n
    theta = 0.5
    if (gen, n) not in _BODIES:
        _BODIES.clear()   # one body set at a time (3.5 GB of host memory at 2^26)
        _BODIES[(gen, n)] = (nb.generators.uniform_sphere(n, seed=1, velocity_scale=0.3) if gen == "uniform_sphere"
                             else getattr(nb.generators, gen)(n, seed=1))
    m, x, y, z, vx, vy, vz = _BODIES[(gen, n)]
    ctx = nb.Context(device=local_rank, theta=theta, wg_size_barnes_hut=128, world_size=world, rank=rank,
                     static_slices=static_slices)
    if world > 1:
        ctx.comm_init(fresh_comm_id(nb, dist, rank, world), world, rank)
    ctx.set_bodies(m, x, y, z, vx, vy, vz)
    dt = 1e-3  # days: bodies move, the tree changes every step
    ctx.bh_build(); ctx.bh_accel(); ctx.synchronize()

    # phase times of one step issued call by call (what the reference's time loop does), walk timed without communication
    ctx.enable_timers(True)
    for _ in range(3):
        ctx.leapfrog_part1(dt); ctx.bh_build(); ctx.bh_accel(); ctx.leapfrog_part2(dt)
    timers = ctx.timers()
    ctx.enable_timers(False)
    walk_per_rank = [timers["Acceleration Kernel Time"]]
    if world > 1:   # balance of the slices: every rank's walk time of the same step
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = timers["Acceleration Kernel Time"]
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        walk_per_rank = [float(v) for v in t.tolist()]

    if rank == 0:
        sampler.start()
    ms, launches = time_advance(ctx, torch, dist, dev, world, dt, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else {}
    sm_mhz = (clocks.get("sm_mhz") if clocks else None) or 1965.0

    # the step's result, read on the host: checksum over ALL bodies (identical for every GPU count)
    a = ctx.accelerations()
    p = ctx.positions()
    checksum = {"sum_abs_a": float(np.abs(a[0]).sum() + np.abs(a[1]).sum() + np.abs(a[2]).sum()),
                "sum_abs_x": float(np.abs(p[0]).sum() + np.abs(p[1]).sum() + np.abs(p[2]).sum()),
                "steps_from_t0": 3 + max(args.warmup, 1) + args.steps}
    no_gate = args.no_parity or not full
    gate = {} if no_gate else bh_parity_gate(nb, ctx, m, theta, world)
    if no_gate:
        ctx.bh_enable_stats(True); ctx.bh_build(); ctx.bh_accel()
        tv, ta = ctx.bh_stats(); info = ctx.bh_tree_info(); ctx.bh_enable_stats(False)
        gate = {"visits": tv, "accepts": ta, "max_depth": int(info.max_depth), "num_internal": int(info.num_internal)}
    visits, accepts = float(gate["visits"]), float(gate["accepts"])
    if world > 1:
        t = torch.tensor([visits, accepts], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        visits, accepts = float(t[0].item()), float(t[1].item())
    b0, b1 = nb.slice_bounds(n, world, rank)
    roof = walk_roofline(visits / n, accepts / n, float(b1 - b0), timers["Acceleration Kernel Time"] * 1e-3,
                         ctx_sm_count(ctx), sm_mhz, n, world)

    # e2e: the reference-facing operator with HOST buffers (H2D of m, x, y, z; build; walk; D2H of the accelerations)
    e2e = None
    try:
        if not full:
            raise StopIteration
        pinned = [torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for v in (m, p[0], p[1], p[2])]
        outs = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
        pin_np = [t.numpy() for t in pinned]
        out_np = [t.numpy() for t in outs]
        ctx.op_barnes_hut_accelerations(*pin_np, out=out_np)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.op_barnes_hut_accelerations(*pin_np, out=out_np)
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": args.steps / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": 4 * 8 * n, "d2h_bytes_per_step": 3 * 8 * n,
               "call": "nb_op_barnes_hut_accelerations (BarnesHutAlgorithm.hpp:43-46 + BarnesHutOctree.hpp:102-103): pinned host "
                       "m, x, y, z in, tree build, walk, accelerations out; the integrator's 0.5 ms is not part of the operator",
               "sum_abs_a": float(np.abs(out_np[0]).sum() + np.abs(out_np[1]).sum() + np.abs(out_np[2]).sum())}
    except StopIteration:
        e2e = None
    except Exception as e:  # noqa: BLE001
        e2e = {"error": repr(e)}
    p2p = ctx.p2p_enabled() if world > 1 else None
    ctx.close()
    return {
        "metric": "Barnes-Hut steps/s", "value": args.steps / (ms * 1e-3), "unit": "steps/s", "ms_per_step": ms / args.steps,
        "config": {"workload": ("Barnes-Hut theta=0.5, uniform sphere N=%d, full step (BASELINE configs[3])" % n) if gen == "uniform_sphere"
                   else "Barnes-Hut theta=0.5, %s sphere N=%d, full step (imbalanced workload for the slice balance, SURVEY 8e)" % (gen, n),
                   "inputs": "larger than L2", "wg_size_barnes_hut": 128,
                   "slices": "equal count (nb_slice_bounds)" if (static_slices or world == 1 or not p2p) else
                             "equal cost: cut from the clock ticks each 32-body tile took in the previous walk",
                   "step": "nb_advance: build + walk with the leapfrog half-steps in its epilogue; %s" %
                           ("one GPU" if world == 1 else ("targets sharded x%d, results stored into every rank's arrays by the walk "
                                                          "(IPC peer memory over NVLink) between two barriers" % world if p2p else
                                                          "targets sharded x%d, NCCL all-gather of accelerations" % world))},
        "phases_ms": {k: round(v, 4) for k, v in timers.items() if v},
        "phases_note": "one step issued call by call (part 1, build, walk, part 2); the walk's time excludes communication",
        "visits_per_body": visits / n, "accepts_per_body": accepts / n, "max_depth": gate["max_depth"],
        "internal_nodes_per_body": gate["num_internal"] / n,
        "roofline": roof, "e2e": e2e, "checksum": checksum, "parity": gate.get("parity"),
        "gpu_launches": int(launches), "p2p": p2p, "clocks": clocks,
        "walk_ms_per_rank": [round(v, 3) for v in walk_per_rank],
        "walk_max_over_mean": max(walk_per_rank) / (sum(walk_per_rank) / len(walk_per_rank)),
    }


def ctx_sm_count(ctx):
    import torch
    return torch.cuda.get_device_properties(ctx.cfg.device).multi_processor_count


def measure_config3(nb, torch, args, dev, sm_mhz):
    """BASELINE configs[2]: Barnes-Hut theta = 0.5, Plummer sphere N = 2^20, one B200 -- with the reference's own
    buildOctree + BarnesHutAlgorithm::computeAccelerations (oracle/_ref, all host cores) timed on the same bodies."""
    n, theta, dt = 1 << 20, 0.5, 1e-3
    m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=1)
    ctx = nb.Context(device=dev.index or 0, theta=theta, wg_size_barnes_hut=128)
    ctx.set_bodies(m, x, y, z, vx, vy, vz)
    ctx.bh_build(); ctx.bh_accel(); ctx.synchronize()
    a0 = ctx.accelerations()
    ctx.enable_timers(True)
    for _ in range(2):
        ctx.leapfrog_part1(dt); ctx.bh_build(); ctx.bh_accel(); ctx.leapfrog_part2(dt)
    timers = ctx.timers()
    ctx.enable_timers(False)
    steps = max(args.steps, 10)
    ms, launches = time_advance(ctx, torch, None, dev, 1, dt, steps, args.warmup)
    ctx.bh_enable_stats(True); ctx.bh_build(); ctx.bh_accel()
    tv, ta = ctx.bh_stats(); ctx.bh_enable_stats(False)
    roof = walk_roofline(tv / n, ta / n, float(n), timers["Acceleration Kernel Time"] * 1e-3, ctx_sm_count(ctx), sm_mhz, n, 1)
    roof["traffic"] = None
    # operator form with host buffers on the t = 0 bodies: the same call the CPU baseline makes
    pinned = [torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for v in (m, x, y, z)]
    outs = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
    pin_np = [t.numpy() for t in pinned]; out_np = [t.numpy() for t in outs]
    ctx.op_barnes_hut_accelerations(*pin_np, out=out_np)
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.op_barnes_hut_accelerations(*pin_np, out=out_np)
    e2e_s = (time.perf_counter() - t0) / steps
    ctx.close()
    line = {"metric": "Barnes-Hut steps/s", "value": steps / (ms * 1e-3), "unit": "steps/s", "ms_per_step": ms / steps,
            "config": {"workload": "Barnes-Hut theta=0.5, Plummer sphere N=%d, full step, one GPU (BASELINE configs[2])" % n},
            "phases_ms": {k: round(v, 4) for k, v in timers.items() if v},
            "visits_per_body": tv / n, "accepts_per_body": ta / n, "roofline": roof,
            "e2e": {"value": 1.0 / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": 4 * 8 * n, "d2h_bytes_per_step": 3 * 8 * n,
                    "call": "nb_op_barnes_hut_accelerations, pinned host buffers"},
            "gpu_launches": int(launches)}
    if not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_bh_baseline(nb, m, x, y, z, theta, compare=(a0, out_np))
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": repr(e)}
    return line


def cpu_bh_baseline(nb, m, x, y, z, theta, compare=None):
    """The reference's own Barnes-Hut force evaluation on the host cores: BarnesHutOctree::buildOctree
    (ParallelOctreeTopDownSubtrees.cpp:15-93) + BarnesHutAlgorithm::computeAccelerations (.cpp:280-401), unmodified source
    compiled into oracle/_ref.  One evaluation of the full configs[2] body set (the bounded sample: ~10-20 s)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refimpl as R
    cores = host_cores()
    R.lib(); R.set_threads(cores)
    t0 = time.perf_counter()
    ax, ay, az, nodes = R.bh_accel(m, x, y, z, theta)
    dt = time.perf_counter() - t0
    out = {"value": 1.0 / dt, "unit": "steps/s", "cores": cores, "kind": "reference",
           "sample": "one force evaluation (buildOctree + computeAccelerations) of all %d bodies, %.1f s; a full step adds two "
                     "streaming leapfrog kernels (< 1 %% of it)" % (m.shape[0], dt),
           "nodes": int(nodes)}
    if compare is not None:   # and the GPU result against the reference's own output on the same bodies
        ref = np.stack([ax, ay, az], 1)
        for name, got in zip(("gpu_vs_reference_max_rel_err", "gpu_operator_vs_reference_max_rel_err"), compare):
            g = np.stack(got, 1)
            out[name] = float((np.linalg.norm(g - ref, axis=1) / np.linalg.norm(ref, axis=1)).max())
    return out


def measure_config1(nb):
    """BASELINE configs[0]: solar-system CSV (178 bodies), dt = 1 h, t_end = 365 d (8760 steps), vs = 1 d, naive
    opt_stage 2 -- wall time of the reference's executable (host cores) and of ours (GPU), same command line."""
    fixture = os.path.join(ROOT, "tests", "golden", "solar_178.csv")
    exes = {"reference": os.path.join(ROOT, "oracle", "_ref", "N_Body_Simulation"),
            "ours": os.path.join(ROOT, "n-body-simulation_b200", "N_Body_Simulation")}
    out = {"config": "naive opt_stage 2, solar_178.csv, dt=1h t_end=365d vs=1d (BASELINE configs[0]); wall seconds of the "
                     "whole executable incl. CSV input and ParaView output"}
    last = {}
    for name, exe in exes.items():
        if not os.path.exists(exe):
            out[name + "_wall_s"] = None
            continue
        with tempfile.TemporaryDirectory() as d:
            t0 = time.perf_counter()
            env = dict(os.environ)
            env.setdefault("CUDA_VISIBLE_DEVICES", "0")   # CUDA start-up time grows with the number of visible devices
            r = subprocess.run([exe, "--file=" + fixture, "--dt=1h", "--t_end=365d", "--vs=1d", "--vs_dir=" + d,
                                "--algorithm=naive", "--opt_stage=2"], capture_output=True, text=True, timeout=600, env=env)
            out[name + "_wall_s"] = time.perf_counter() - t0
            out[name + "_rc"] = r.returncode
            for root, _, files in os.walk(d):
                if "lastState.csv" in files:
                    last[name] = open(os.path.join(root, "lastState.csv")).read().split()
    if len(last) == 2 and len(last["ours"]) == len(last["reference"]):
        diff = sum(1 for u, v in zip(last["ours"], last["reference"]) if u != v)
        out["lastState_tokens_different"] = diff
        out["lastState_tokens"] = len(last["ours"])
    return out


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
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_mhz = (clocks.get("sm_mhz") if clocks else None) or 1965.0
    pipe_peak = sm_count * FP64_LANES_PER_SM * 2.0 * sm_mhz * 1e6 / 1e12   # TFLOP/s: one DFMA per lane and clock
    roofline = {"bound": "fp64", "kernel": "naive_accel_kernel", "achieved": achieved_tf, "peak": pipe_peak,
                "unit": "TFLOP/s", "frac": achieved_tf / pipe_peak if pipe_peak else None,
                "traffic": NCU_TRAFFIC_BYTES.get(("naive", n)) if world == 1 else None,
                "traffic_source": "ncu capture %s (not measured in this run)" % NCU_TRAFFIC_SOURCE["naive"],
                "peak_source": "FP64 pipe rate: %d SMs x 64 lanes x 2 flop x %.0f MHz (SM clock sampled under load in this run; "
                               "MEASURED_PEAKS.json has no fp64 figure)" % (sm_count, sm_mhz),
                "peak_microbenchmark": fp64_peak, "frac_of_microbenchmark": achieved_tf / fp64_peak if fp64_peak else None,
                "peak_microbenchmark_source": "DFMA-chain microbenchmark run in this process (reaches ~92 % of the pipe rate)",
                "algorithmic_flop_per_interaction": FLOP_PER_INTERACTION, "kernel_ms": k_ms,
                "dp_instructions_per_interaction": 15,
                "instruction_mix_ceiling": FLOP_PER_INTERACTION / 30.0}

    bh = None
    extra = {}
    if not args.no_bh:
        ctx.close()
        bh_sampler = ClockSampler(local_rank, interval_ms=50)   # started inside, right before the timed steps
        try:
            bh = measure_bh(nb, torch, dist, args, rank, local_rank, world, dev, bh_sampler)
        except Exception as e:  # the headline line must survive a failure of the secondary metric
            import traceback
            bh = {"error": repr(e), "traceback": traceback.format_exc()[-1500:]}
            if rank == 0 and bh_sampler.proc:
                bh_sampler.stop()
        if not args.no_plummer:   # imbalanced workload: slice balance with equal-cost and (several GPUs) equal-count slices
            for name, static in (("bh_plummer", 0),) + ((("bh_plummer_static_slices", 1),) if world > 1 else ()):
                try:
                    r = measure_bh(nb, torch, dist, args, rank, local_rank, world, dev, ClockSampler(local_rank, interval_ms=50),
                                   gen="plummer", full=False, static_slices=static)
                except Exception as e:  # noqa: BLE001
                    r = {"error": repr(e)}
                if rank == 0:
                    extra[name] = r
        if rank == 0 and world == 1:
            for name, fn in (("bh_config3", lambda: measure_config3(nb, torch, args, dev, sm_mhz)),
                             ("config1", lambda: measure_config1(nb))):
                try:
                    extra[name] = fn()
                except Exception as e:  # noqa: BLE001
                    extra[name] = {"error": repr(e)}
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
        line.update(extra)
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
