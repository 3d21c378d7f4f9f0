import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onbody_b200.api import GpuSession, driver_inputs
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
x, r, s = driver_inputs("grav3d", N, True)
g = GpuSession("grav3d", N, N)
for rep in range(3):
    g.set_sources(x, r, s); g.set_targets(x, r)
    g.make_tree(0); a = g.phase_ms("tree"); g.make_tree(1); b = g.phase_ms("tree")
    g.set_sources(x, r, s); g.set_targets(x, r)
    g.make_trees(); c = g.phase_ms("tree")
    print("sequential %.2f + %.2f = %.2f ms   concurrent %.2f ms" % (a, b, a + b, c), flush=True)
