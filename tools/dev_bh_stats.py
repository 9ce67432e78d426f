import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("n-body-simulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
m, x, y, z, vx, vy, vz = nb.generators.uniform_sphere(n, seed=1)
c = nb.Context(theta=0.5)
c.set_bodies(m, x, y, z, vx, vy, vz); c.bh_enable_stats(True)
c.bh_build(); c.bh_accel(); print(n, c.bh_stats(), n // 32, "warps")
