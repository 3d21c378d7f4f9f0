import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onbody_b200.api import GpuSession, driver_inputs
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
x, r, s = driver_inputs("grav3d", N, True)
g = GpuSession("grav3d", N, N)
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    g.set_sources(x, r, s)
    g.make_tree(0)
    print("tree %.2f ms" % g.phase_ms("tree"), g.build_stats(), flush=True)
