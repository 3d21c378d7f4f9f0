"""small end-to-end pass for compute-sanitizer (memcheck): every kernel family once"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onbody_b200.api import GpuSession
for phys, n in (("grav3d", 40000), ("vortgrad3d", 3000), ("vort2dtr", 5000)):
    g = GpuSession(phys, n, n)
    g.init_driver(); g.make_tree(0); g.upward(0); g.make_tree(1); g.refine(1); g.upward(1)
    g.zero_vels(); g.naive(50); g.zero_vels(); g.treecode1(1.3); g.zero_vels(); g.treecode2(1.3); g.zero_vels(); g.treecode3(1.3)
    if g.has_fastsumm:
        g.zero_vels(); g.fastsumm(1.3)
    lo, hi = g.shard_particle_range(n, 1, 3)
    g.init_driver(); g.set_shard(1, 3); g.make_tree_range(0, lo, hi); g.finish_tree(0); g.upward(0)
    g.make_tree_range(1, lo, hi); g.finish_tree(1); g.set_build_range(1, lo, hi); g.refine(1); g.upward(1)
    if g.has_fastsumm:
        g.zero_vels(); g.fastsumm(1.3)
    print(phys, "ok", g.stats(), flush=True)
    g.close()
# concurrent builds + the legacy (-o omitted) equivalents
g = GpuSession("grav3d", 50000, 50000, order=-1)
g.init_driver(); g.make_trees(); g.refine(0); g.upward(0); g.zero_vels(); g.treecode2(1.2); g.zero_vels(); g.treecode3(1.2)
print("legacy ok", g.stats(), flush=True)
g.close()
