"""relative rms deviation of the fast (product) dual-tree result from the strict oracle: python tools/dev_vs_oracle.py [N] [theta]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from onbody_b200.api import GpuSession, ARITH_FAST
from oracle.refapi import PortSession
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000
theta = float(sys.argv[2]) if len(sys.argv) > 2 else 1.4
for phys in ("grav3d", "vort3d"):
    o = PortSession(phys, N, N); o.init_driver(); o.make_tree(0); o.upward(0); o.make_tree(1); o.refine(1); o.upward(1); o.zero_vels(); o.fastsumm(theta)
    uo = o.parts(1)["u"].astype(np.float64)
    g = GpuSession(phys, N, N, arith=ARITH_FAST); g.init_driver(); g.make_trees(); g.upward(0); g.refine(1); g.upward(1); g.zero_vels(); g.fastsumm(theta)
    ug = g.parts(1, ("u",))["u"].astype(np.float64)
    print("%s N=%d theta=%g split_target=%s: rel rms deviation from the strict oracle %.3e" % (phys, N, theta, os.environ.get("ONB_P2P_SPLIT_TARGET", "default"), np.sqrt(((ug - uo) ** 2).sum() / (uo ** 2).sum())), flush=True)
