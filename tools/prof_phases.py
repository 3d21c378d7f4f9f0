import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from onbody_b200.api import GpuSession, driver_inputs
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
x, r, s = driver_inputs("grav3d", N, True)
dx = torch.from_numpy(x).cuda(); dr = torch.from_numpy(r).cuda(); ds = torch.from_numpy(s).cuda()
g = GpuSession("grav3d", N, N)
def step(tag):
    t = {}
    torch.cuda.synchronize()
    g.set_sources_ptr(N, dx.data_ptr(), dr.data_ptr(), ds.data_ptr()); g.set_targets_ptr(N, dx.data_ptr(), dr.data_ptr())
    t0 = time.perf_counter()
    for name, fn in (("src_tree", lambda: g.make_tree(0)), ("upward", lambda: g.upward(0)), ("tgt_tree", lambda: g.make_tree(1)),
                     ("refine", lambda: g.refine(1)), ("tgt_eq", lambda: g.upward(1)), ("fastsumm", lambda: g.fastsumm(1.4))):
        a = time.perf_counter(); fn(); t[name] = (time.perf_counter() - a) * 1e3
    tot = (time.perf_counter() - t0) * 1e3
    print(tag, "total %.1f" % tot, " ".join("%s %.1f" % kv for kv in t.items()), "| gpu: lists %.1f p2p %.1f down %.1f eval %.1f" % (g.phase_ms("lists"), g.phase_ms("p2p"), g.phase_ms("downward"), g.phase_ms("eval")), flush=True)
for i in range(6): step("step%d" % i)
