import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("n-body-simulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
single = int(sys.argv[2]) if len(sys.argv) > 2 else 0
wv = int(sys.argv[3]) if len(sys.argv) > 3 else 0
wg = int(sys.argv[4]) if len(sys.argv) > 4 else 128
m, x, y, z, vx, vy, vz = nb.generators.uniform_sphere(n, seed=1)
c = nb.Context(theta=0.5, wg_size_barnes_hut=wg, single_phase_walk=single, walk_variant=wv)
c.set_bodies(m, x, y, z, vx, vy, vz); c.enable_timers(True)
for _ in range(3):
    c.bh_build(); c.bh_accel()
c.synchronize()
print(n, 'bh_variant', single, 'walk_variant', wv, 'wg', wg, {k: round(v, 3) for k, v in c.timers().items() if v})
