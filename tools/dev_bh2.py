import importlib, os, sys, faulthandler
import numpy as np
faulthandler.dump_traceback_later(int(os.environ.get("HANG_S", "40")), exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
nb = importlib.import_module("n-body-simulation_b200")
import oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
def p(*a): print(*a, flush=True)
m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=7)
ctx = nb.Context(theta=0.5); p("ctx")
ctx.set_bodies(m, x, y, z); p("set")
ctx.bh_build(); p("build enq")
ctx.synchronize(); p("build sync")
info = ctx.bh_tree_info(); p("info", info.num_nodes_materialised, info.num_internal, info.max_depth)
t = O.Tree(m, x, y, z); p("oracle tree", t.num_nodes)
c = ctx.bh_export_canonical(); p("export", c["kind"])
p("sorted", ctx.bh_sorted_bodies())
ctx.bh_enable_stats(True)
ctx.bh_build(); ctx.bh_accel(); p("accel enq")
ctx.synchronize(); p("accel sync")
a = ctx.accelerations(); p("acc", a)
p(ctx.bh_stats(per_body=True))
p(t.accel(0.5, stats=True))
