import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from onbody_b200.api import GpuSession, driver_inputs
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
x, r, s = driver_inputs("grav3d", N, True)
hx = torch.from_numpy(x).pin_memory(); hr = torch.from_numpy(r).pin_memory(); hs = torch.from_numpy(s).pin_memory()
hu = torch.empty((3, N), dtype=torch.float32).pin_memory()
g = GpuSession("grav3d", N, N)
def step(tag):
    t = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for name, fn in (("h2d_s", lambda: g.set_sources_ptr(N, hx.data_ptr(), hr.data_ptr(), hs.data_ptr())), ("h2d_t", lambda: g.set_targets_ptr(N, hx.data_ptr(), hr.data_ptr())),
                     ("src_tree", lambda: g.make_tree(0)), ("upward", lambda: g.upward(0)), ("tgt_tree", lambda: g.make_tree(1)),
                     ("refine", lambda: g.refine(1)), ("tgt_eq", lambda: g.upward(1)), ("fastsumm", lambda: g.fastsumm(1.4)), ("d2h", lambda: g.results_into(hu.data_ptr()))):
        a = time.perf_counter(); fn(); t[name] = (time.perf_counter() - a) * 1e3
    tot = (time.perf_counter() - t0) * 1e3
    print(tag, "total %.1f" % tot, " ".join("%s %.1f" % kv for kv in t.items()), flush=True)
for i in range(12): step("e2e%d" % i)
