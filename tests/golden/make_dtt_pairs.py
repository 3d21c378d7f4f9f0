"""Writes tests/golden/dtt_pairs.json: exact dual-tree pair counts of `ongrav3d -n=N -t=1.4 -o=4 -b=128` (seed 12345)
from the GPU list builder (run on a B200; the lists are bit-identical to the reference's, see test_gpu_parity.py)."""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from onbody_b200.api import GpuSession
res = {}
for n in (50000, 100000, 200000, 500000, 1000000, 10000000):
    g = GpuSession("grav3d", n, n)
    g.init_driver(); g.make_tree(0); g.upward(0); g.make_tree(1); g.refine(1); g.upward(1)
    g.zero_vels(); g.fastsumm(1.4)
    res[n] = g.last_pairs(); print(n, res[n], g.stats(), flush=True)
    g.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/dtt_pairs.json", "w"))
json.dump(res, open(os.path.join(HERE, "dtt_pairs.json"), "w"))
