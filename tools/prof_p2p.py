"""profiling helper: one boxwise (treecode3) evaluation = ONE k_p2p_lists launch over every target leaf"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onbody_b200.api import GpuSession, load_library
phys = sys.argv[1] if len(sys.argv) > 1 else "grav3d"
N = int(float(sys.argv[2])) if len(sys.argv) > 2 else 4000000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
tpts = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
FL = {"grav3d": (19, 12), "vort3d": (28, 17), "vortgrad3d": (64, 37), "vort2d": (13, 7), "vort2dtr": (15, 8)}[phys]
g = GpuSession(phys, N, N)
g.init_driver(); g.make_tree(0); g.upward(0); g.make_tree(1); g.refine(1); g.upward(1)
peak = g.measure_fp32_peak()
for tpt in tpts:
    load_library().onb_set_p2p_tpt(tpt)
    for rep in range(reps):
        g.zero_vels(); g.treecode3(1.4)
        pairs = g.last_pairs(); ms = g.phase_ms("p2p")
    print("%s N=%d tpt=%d boxwise pairs %d p2p %.3f ms -> %.1f Gpairs/s = %.1f TFLOP/s (%.1f%% of %.1f), fp32 issue %.1f%%" % (
        phys, N, tpt, pairs, ms, pairs / ms * 1e-6, pairs * FL[0] / ms * 1e-9, pairs * FL[0] / ms * 1e-9 / peak * 100, peak,
        pairs * FL[1] * 2 / ms * 1e-9 / peak * 100), flush=True)
