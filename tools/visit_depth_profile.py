"""Where do the walk's node visits go, by tree depth?  CPU analysis with the oracle's tree (no GPU): every sampled body
walks the canonical tree with the reference's criterion and the visits are binned by the depth of the visited node.
Answers whether staging the top levels of the tree in shared memory (north_star) can matter: a level is worth staging
only if a large share of the node loads go to it AND those loads miss L1 today.
usage: python tools/visit_depth_profile.py [N] [theta] [generator] [samples]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O
nb = importlib.import_module("n-body-simulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
theta = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
gen = sys.argv[3] if len(sys.argv) > 3 else "uniform_sphere"
samples = int(sys.argv[4]) if len(sys.argv) > 4 else 512
m, x, y, z, *_ = getattr(nb.generators, gen)(n, seed=1)
t = O.Tree(m, x, y, z, storage_param=6, insertion_order="morton")
S = t.S
oct_ = t._arr("octants", np.uint32, 8 * S).reshape(8, S)
mass = t._arr("sum_masses", np.float64, t.num_nodes)
cx = t._arr("com_x", np.float64, t.num_nodes) ; cy = t._arr("com_y", np.float64, t.num_nodes); cz = t._arr("com_z", np.float64, t.num_nodes)
edge = t._arr("edge", np.float64, t.num_nodes)
bon = t.body_of_node
rng = np.random.default_rng(3)
ids = rng.integers(0, n, samples)
visits = np.zeros(64); accepts = np.zeros(64)
for i in ids:
    stack = [(0, 0)]
    while stack:
        node, d = stack.pop()
        if mass[node] != 0 and bon[node] != i:
            visits[d] += 1
            dx = cx[node] / mass[node] - x[i]; dy = cy[node] / mass[node] - y[i]; dz = cz[node] / mass[node] - z[i]
            r = 1.0 / np.sqrt(dx * dx + dy * dy + dz * dz)
            if edge[node] * r < theta or bon[node] != n:
                accepts[d] += 1
            else:
                for o in (5, 7, 4, 6, 1, 3, 0, 2):
                    stack.append((int(oct_[o, node]), d + 1))
tot = visits.sum()
print("N=%d %s theta=%g: %.0f non-empty visits per body (%d sampled bodies), tree depth %d" % (n, gen, theta, tot / samples, samples, t.max_depth))
print("depth  nodes_at_depth<=d(max)  share_of_visits  cumulative")
cum = 0.0
for d in range(0, t.max_depth + 1):
    cum += visits[d] / tot
    print("%5d  %20d  %14.2f%%  %9.2f%%" % (d, (8 ** (d + 1) - 1) // 7, 100 * visits[d] / tot, 100 * cum))
