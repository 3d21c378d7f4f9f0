"""one pointwise treecode evaluation (nbody_treecode2: k_pointwise, the fused warp traversal): the launch captured by
profiles/r2_ncu_full_pointwise.txt     python tools/prof_pointwise.py [N] [theta]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onbody_b200.api import GpuSession
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2000000
theta = float(sys.argv[2]) if len(sys.argv) > 2 else 1.11111
g = GpuSession("grav3d", N, N); g.init_driver()
g.make_tree(0); g.upward(0)
if os.environ.get("ONB_PW_UNSORTED") is None:
    # the drivers build (and, when the dual tree is scheduled, refine) the target tree before any treecode runs
    # (ongrav3d.cpp:675-724), so a warp's 32 consecutive targets are spatial neighbours; ONB_PW_UNSORTED=1 skips it
    g.make_tree(1); g.refine(1)
for it in range(2):
    g.zero_vels(); g.treecode2(theta)
    st = g.stats()
    print("treecode2 N=%d theta=%g: eval %.3f ms, pairs %d -> %.1f Gpairs/s (%.1f TFLOP/s at 19 flop/pair); leaf visits %d box visits %d" % (
        N, theta, g.phase_ms("eval"), g.last_pairs(), g.last_pairs() / g.phase_ms("eval") * 1e-6, g.last_pairs() * 19 / g.phase_ms("eval") * 1e-9, st["sltp"], st["sbtp"]), flush=True)
