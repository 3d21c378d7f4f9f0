import sys, time
sys.path.insert(0, '/root/repo')
from onbody_b200.api import GpuSession
for n in (1000000,):
    g = GpuSession("grav3d", n, n); g.init_driver()
    for it in range(3):
        t=time.time(); g.make_tree(0); t1=time.time(); g.upward(0); t2=time.time(); g.make_tree(1); t3=time.time(); g.refine(1); g.upward(1); t4=time.time()
        g.zero_vels(); g.fastsumm(1.4); t5=time.time()
        print("N=%d iter %d: tree0 %.4f upward0 %.4f tree1 %.4f refine+up1 %.4f fastsumm %.4f attempts %s pool %s/%s" % (n, it, t1-t, t2-t1, t3-t2, t4-t3, t5-t4, g.phase_ms("dtt_attempts"), g.phase_ms("dtt_pool_used"), g.phase_ms("dtt_pool_cap")), flush=True)
