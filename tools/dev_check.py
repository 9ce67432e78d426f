"""Developer GPU check: parity vs the oracle + quick timings.  Run under gpurun; writes gpurun_out/dev_check.log."""
import importlib, os, sys, time, traceback, faulthandler
faulthandler.dump_traceback_later(int(os.environ.get("HANG_S", "500")), exit=True)
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
nb = importlib.import_module("n-body-simulation_b200")
import oracle as O

def relerr(a, b):
    a = np.stack(a, 1); b = np.stack(b, 1)
    return (np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)).max()

def section(name):
    print("\n=== " + name, flush=True)

def run(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()
    sys.stdout.flush()

def t_naive():
    section("naive parity")
    m, x, y, z, vx, vy, vz = nb.generators.plummer(4096 + 37, seed=2)
    ref = O.naive_accel(m, x, y, z)
    for ipt in (1, 2, 4):
        for precise in (1, 0):
            for bs in (64, 256, 100):
                ctx = nb.Context(ipt=ipt, precise_rsqrt=precise, block_size=bs)
                ctx.set_bodies(m, x, y, z, vx, vy, vz)
                ctx.naive_accel()
                a = ctx.accelerations()
                print("ipt", ipt, "precise", precise, "bs", bs, "max rel err", relerr(a, ref))
                ctx.close()
    m, x, y, z, vx, vy, vz = nb.generators.solar_like(178)
    ref = O.naive_accel(m, x, y, z)
    ctx = nb.Context(); ctx.set_bodies(m, x, y, z, vx, vy, vz); ctx.naive_accel()
    print("solar_like 178 rel err", relerr(ctx.accelerations(), ref))
    out = ctx.op_naive_accelerations(m, x, y, z)
    print("op form rel err", relerr(out, ref))

def t_integrator():
    section("integrator + energy")
    m, x, y, z, vx, vy, vz = nb.generators.plummer(3000, seed=5)
    ctx = nb.Context(); ctx.set_bodies(m, x, y, z, vx, vy, vz); ctx.naive_accel()
    ax, ay, az = ctx.accelerations()
    dt = 1.0 / 24
    X, Y, Z = x.copy(), y.copy(), z.copy()
    vh = O.leapfrog_part1(dt, X, Y, Z, vx, vy, vz, ax, ay, az)
    ctx.leapfrog_part1(dt)
    p = ctx.positions(); v = ctx.velocities()
    print("lf1 pos bitwise", all(np.array_equal(a, b) for a, b in zip(p, (X, Y, Z))), "vhalf bitwise", all(np.array_equal(a, b) for a, b in zip(v, vh)))
    V = [vx.copy(), vy.copy(), vz.copy()]
    O.leapfrog_part2(dt, *V, *vh, ax, ay, az)
    ctx.leapfrog_part2(dt)
    v = ctx.velocities()
    print("lf2 bitwise", all(np.array_equal(a, b) for a, b in zip(v, V)))
    e_gpu = ctx.energy()
    e_ref = O.energy(m, X, Y, Z, *V)
    print("energy gpu", e_gpu, "\n       ref", e_ref, "\n rel", np.abs(e_gpu - e_ref) / np.abs(e_ref))
    an = ctx.acceleration_norms()
    print("anorm bitwise", np.array_equal(an, O.accel_norm(ax, ay, az)))
    # fused part2+part1
    ctx2 = nb.Context(); ctx2.set_bodies(m, x, y, z, vx, vy, vz); ctx2.naive_accel(); ctx2.leapfrog_part1(dt)
    ctx2.leapfrog_part2_part1(dt)
    ctx.leapfrog_part1(dt)
    print("fused bitwise", all(np.array_equal(a, b) for a, b in zip(ctx.positions() + ctx.velocities(), ctx2.positions() + ctx2.velocities())))

def canon_compare(cg, co):
    ok = True
    n = len(co["depth"])
    if len(cg["depth"]) != n:
        print("  node count differs", len(cg["depth"]), n); return False
    for k in ("depth", "path_hi", "path_lo", "kind", "body", "count"):
        if not np.array_equal(cg[k], co[k]):
            bad = np.nonzero(cg[k] != co[k])[0]
            print("  field", k, "differs at", bad[:5], cg[k][bad[:5]], co[k][bad[:5]]); ok = False
    for k in ("edge", "minx", "miny", "minz", "mass", "comx", "comy", "comz"):
        if not np.array_equal(cg[k], co[k]):
            d = np.abs(cg[k] - co[k]); s = np.abs(co[k]) + 1e-300
            print("  field", k, "not bitwise; max rel", (d / s).max(), "count", (d > 0).sum()); ok = False
    return ok

def t_bh_small():
    section("BH golden 3 bodies")
    x = np.array([0, 0, 2.]); y = np.array([1, 0, 0.]); z = np.array([0, 2, 0.]); m = np.array([10, 10, 10.])
    ctx = nb.Context(); ctx.set_bodies(m, x, y, z)
    print("aabb", ctx.bh_aabb())
    ctx.bh_build(); info = ctx.bh_tree_info()
    print("nodes", info.num_nodes_materialised, info.num_internal, info.num_nodes_canonical, "depth", info.max_depth)
    cg = ctx.bh_export_canonical(); t = O.Tree(m, x, y, z); co = t.canonical()
    print("canonical equal", canon_compare(cg, co), "sorted", ctx.bh_sorted_bodies(), t.sorted_bodies)
    print("group_by_subtree", ctx.util_group_by_subtree([1, 1, 1, 1, 0, 4, 4, 5, 5, 7], 9))
    for n in (1, 2, 5, 33, 1000, 20000):
        for gen in ("plummer", "uniform_sphere"):
            m, x, y, z, vx, vy, vz = getattr(nb.generators, gen)(n, seed=7)
            ctx = nb.Context(theta=0.5); ctx.set_bodies(m, x, y, z)
            ctx.bh_build(); info = ctx.bh_tree_info()
            t = O.Tree(m, x, y, z)
            ok = canon_compare(ctx.bh_export_canonical(), t.canonical())
            okab = np.array_equal(ctx.bh_aabb(), t.aabb())
            sb = np.array_equal(ctx.bh_sorted_bodies(), t.sorted_bodies)
            ctx.bh_enable_stats(True)
            ctx.bh_build(); ctx.bh_accel(); a = ctx.accelerations()
            tv, ta, pb = ctx.bh_stats(per_body=True)
            ax, ay, az, st = t.accel(0.5, stats=True)
            err = relerr(a, (ax, ay, az)) if n > 1 else 0.0
            print(gen, n, "canon", ok, "aabb", okab, "sorted", sb, "depth", info.max_depth, t.max_depth, "visits equal", np.array_equal(pb, st[:, 1].astype(np.uint32)), tv, int(st[:, 1].sum()), "acc", ta, int(st[:, 2].sum()), "rel err", err)
            ctx.close()

def t_perf():
    section("perf")
    ctx = nb.Context()
    print("device", ctx.device_name(), "fp64 peak TF", ctx.measure_fp64_peak(), ctx.measure_fp64_peak())
    n = 1 << 17
    m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=1)
    for ipt in (1, 2, 4):
        for precise in (1, 0):
            for bs in (64, 128, 256, 512):
                c = nb.Context(ipt=ipt, precise_rsqrt=precise, block_size=bs)
                c.set_bodies(m, x, y, z, vx, vy, vz); c.enable_timers(True)
                c.naive_accel(); c.synchronize()
                c.naive_accel(); ms = c.timers()["Acceleration Kernel Time"]
                print("naive n=%d ipt=%d precise=%d bs=%d: %.3f ms  %.3e inter/s  %.2f TF(21)" % (n, ipt, precise, bs, ms, n * n / ms * 1e3, 21.0 * n * n / ms * 1e3 / 1e12), flush=True)
                c.close()
    for gen, n, theta in (("plummer", 1 << 20, 0.5), ("uniform_sphere", 1 << 20, 0.5), ("uniform_sphere", 1 << 22, 0.5)):
        m, x, y, z, vx, vy, vz = getattr(nb.generators, gen)(n, seed=1)
        for wg in (64, 128, 256):
            c = nb.Context(theta=theta, wg_size_barnes_hut=wg); c.set_bodies(m, x, y, z, vx, vy, vz); c.enable_timers(True)
            c.bh_build(); c.bh_accel(); c.synchronize()
            c.bh_enable_stats(wg == 64)
            c.bh_build(); c.bh_accel(); t = c.timers(); info = c.bh_tree_info()
            extra = ""
            if wg == 64:
                tv, ta = c.bh_stats(); extra = "visits/body %.1f accepts/body %.1f" % (tv / n, ta / n)
            print(gen, n, "wg", wg, "depth", info.max_depth, "internal/body %.3f" % (info.num_internal / n), {k: round(v, 3) for k, v in t.items() if v}, extra, flush=True)
            c.close()

def t_naive_sweep():
    section("naive sweep N=2^20")
    n = 1 << 20
    m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=1)
    names = {0: "ipt2 u4 b2 (default)", 1: "ipt6 u1 b1", 2: "ipt6 u2 b1", 3: "ipt4 u3 b2", 4: "ipt2 u4 b4", 5: "ipt2 u8 b2", 6: "ipt8 u1 b1",
             7: "ipt4 u2 b2", 8: "ipt4 u2 b3", 9: "ipt4 u4 b1", 10: "ipt3 u4 b2", 11: "ipt3 u2 b2"}
    for var in range(12):
        for bs in (128, 256):
            c = nb.Context(naive_variant=var, block_size=bs)
            c.set_bodies(m, x, y, z, vx, vy, vz); c.enable_timers(True)
            c.naive_accel(); ms = c.timers()["Acceleration Kernel Time"]
            print("variant %2d %-22s bs=%d: %.2f ms  %.4e inter/s  %.2f TF(21)" % (var, names[var], bs, ms, n * n / ms * 1e3, 21.0 * n * n / ms * 1e3 / 1e12), flush=True)
            c.close()


def t_naive_segments():
    section("naive segments")
    for n in (1 << 20, 1 << 17):
        m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=1)
        for seg in (0, 1, 2, 4, 8, 16):
            c = nb.Context(naive_segments=seg, block_size=256)
            c.set_bodies(m, x, y, z, vx, vy, vz); c.enable_timers(True)
            c.naive_accel(); c.naive_accel(); ms = c.timers()["Acceleration Kernel Time"]
            print("n=%d segments=%d: %.2f ms  %.4e inter/s  %.2f TF(21)" % (n, seg, ms, n * n / ms * 1e3, 21.0 * n * n / ms * 1e3 / 1e12), flush=True)
            c.close()


def t_energy_perf():
    section("energy perf")
    for n in (1 << 17, 1 << 19, 1 << 20):
        m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=1)
        c = nb.Context(); c.set_bodies(m, x, y, z, vx, vy, vz); c.enable_timers(True)
        c.energy(); e = c.energy(); ms = c.timers()["Energy"]
        pairs = n * (n - 1) / 2
        print("n=%d: %.2f ms  %.3e pairs/s  (12 DP ops/pair -> %.2f T DP-inst-lanes/s)" % (n, ms, pairs / ms * 1e3, 12 * pairs / ms * 1e3 / 1e12), e, flush=True)
        c.close()


def t_perf_bh():
    section("perf bh")
    for gen, n, theta in (("plummer", 1 << 20, 0.5), ("uniform_sphere", 1 << 22, 0.5), ("uniform_sphere", 1 << 24, 0.5)):
        m, x, y, z, vx, vy, vz = getattr(nb.generators, gen)(n, seed=1)
        for single in (0, 3):
            for wg in (64, 128, 256):
                c = nb.Context(theta=theta, wg_size_barnes_hut=wg, single_phase_walk=single); c.set_bodies(m, x, y, z, vx, vy, vz); c.enable_timers(True)
                c.bh_build(); c.bh_accel(); c.synchronize()
                c.bh_build(); c.bh_accel(); t = c.timers(); info = c.bh_tree_info()
                print(gen, n, "walk" if single == 0 else "group", "wg", wg, "depth", info.max_depth, {k: round(v, 3) for k, v in t.items() if v}, flush=True)
                c.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["naive", "integrator", "bh_small", "perf"]
    for w in which:
        run(globals()["t_" + w])
