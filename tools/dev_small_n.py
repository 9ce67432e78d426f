"""Per-step wall time of the small-N time loop (BASELINE configs[0] size): eager calls vs nb_advance (CUDA graph)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("n-body-simulation_b200")
rows = [l.rstrip("\n").split(",") for l in open(os.path.join(ROOT, "tests", "golden", "solar_178.csv"))][1:]
cols = [np.array(c, dtype=np.float64) for c in list(zip(*rows))[3:10]]
dt, K = 1.0 / 24, 2400
for alg in ("naive", "BarnesHut"):
    c = nb.Context(device=0)
    c.set_bodies(*cols)
    f = (lambda: c.naive_accel()) if alg == "naive" else (lambda: (c.bh_build(), c.bh_accel()))
    f(); c.advance(alg, dt, 24); c.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        c.leapfrog_part1(dt); f(); c.leapfrog_part2(dt)
    c.synchronize(); t_eager = (time.perf_counter() - t0) / K
    t0 = time.perf_counter()
    for _ in range(K // 24):
        c.advance(alg, dt, 24)
    c.synchronize(); t_graph = (time.perf_counter() - t0) / K
    print("%-9s N=178: eager %.1f us/step, nb_advance(24) %.1f us/step" % (alg, t_eager * 1e6, t_graph * 1e6), flush=True)
    c.close()
