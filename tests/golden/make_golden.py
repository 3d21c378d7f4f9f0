"""Generates tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/strict, compiled in place from
/root/reference by oracle/Makefile). Run in the build container only:  python tests/golden/make_golden.py

Every entry is an FNV-1a-64 hash over the raw little-endian bytes of an array the reference produced (the same hash
SURVEY.md section 4 quotes), or a scalar it returned. tests/test_oracle.py pins the CPU restatement (oracle/port) to
these on any machine; tests/test_gpu_parity.py pins the CUDA path to them on the GPU box.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.refapi import RefSession, fnv1a64  # noqa: E402

CASES = [("grav3d", 20000, 1.2), ("vort3d", 20000, 1.2), ("vortgrad3d", 20000, 1.2), ("vort2d", 20000, 1.2), ("vort2dtr", 20000, 1.2),
         ("grav3d", 100000, 1.11111), ("grav3d", 5000, 1.4), ("grav3d", 129, 1.4), ("grav3d", 100, 1.4)]


def h(a):
    return "%016x" % fnv1a64(a)


def one(physics, n, theta):
    out = {"physics": physics, "n": n, "theta": theta}
    s = RefSession(physics, n, n, build="strict")
    s.init_driver()
    s.make_tree(0)
    p = s.parts(0); t = s.tree(0)
    out["srcs.x"] = h(p["x"]); out["srcs.s"] = h(p["s"]); out["srcs.r"] = h(p["r"])
    for k in ("num", "ioffset", "nc", "ns", "nr", "x", "s", "pr"):
        out["stree." + k] = h(t[k])
    out["stree.levels"] = t["levels"]; out["stree.numnodes"] = t["numnodes"]
    s.upward(0)
    e = s.parts(2)
    out["eqsrcs.x"] = h(e["x"]); out["eqsrcs.s"] = h(e["s"]); out["eqsrcs.r"] = h(e["r"])
    out["eqsrcs.s0.sum"] = float(e["s"][0].astype(np.float64).sum())
    s.make_tree(1); s.refine(1)
    p = s.parts(1); t = s.tree(1)
    out["targs.x"] = h(p["x"]); out["targs.gidx"] = h(p["gidx"]); out["targs.gidx.first8"] = [int(v) for v in p["gidx"][:8]]
    for k in ("num", "ioffset", "nc", "ns", "nr", "x", "pr"):
        out["ttree." + k] = h(t[k])
    s.upward(1)
    out["eqtargs.x"] = h(s.parts(3)["x"])
    if n <= 20000:
        tsk = max(1, n // 400)
        out["tskip"] = tsk
        s.zero_vels(); out["naive.flops"] = s.naive(tsk); u = s.parts(1)["u"]; out["naive.u"] = h(u); out["naive.u0"] = [float(v) for v in u[:, 0]]
        for name in ("treecode1", "treecode2", "treecode3"):
            s.zero_vels(); out[name + ".flops"] = getattr(s, name)(theta); u = s.parts(1)["u"]
            out[name + ".u"] = h(u); out[name + ".u0"] = [float(v) for v in u[:, 0]]
        if s.has_fastsumm:
            s.zero_vels(); s.fastsumm(theta); u = s.parts(1)["u"]
            out["fastsumm.u"] = h(u); out["fastsumm.u0"] = [float(v) for v in u[:, 0]]; out["fastsumm.equ"] = h(s.parts(3)["u"])
    return out


def legacy(physics="grav3d", n=20000, theta=1.2):
    """-o omitted (order = -1): refineTree(srcs) + calcEquivalents (barneshut.hpp:946-1061), treecode2/3 over them"""
    out = {"physics": physics, "n": n, "theta": theta}
    s = RefSession(physics, n, n, order=-1, build="strict")
    s.init_driver(); s.make_tree(0); s.refine(0); s.upward(0); s.make_tree(1)
    e, t = s.parts(2), s.tree(0)
    out["srcs.x"] = h(s.parts(0)["x"]); out["eqsrcs.x"] = h(e["x"]); out["eqsrcs.r"] = h(e["r"]); out["eqsrcs.s"] = h(e["s"])
    out["stree.epnum"] = h(t["epnum"]); out["stree.epoffset"] = h(t["epoffset"])
    for name in ("treecode2", "treecode3"):
        s.zero_vels(); out[name + ".flops"] = getattr(s, name)(theta); out[name + ".u"] = h(s.parts(1)["u"])
    return out


if __name__ == "__main__":
    if "--legacy-only" in sys.argv:      # add the legacy case to the existing file without touching the others
        with open(os.path.join(HERE, "golden.json")) as f:
            g = json.load(f)
        g["legacy"] = legacy()
        with open(os.path.join(HERE, "golden.json"), "w") as f:
            json.dump(g, f, indent=1)
        print("wrote the legacy case")
        sys.exit(0)
    res = [one(*c) for c in CASES]
    # SURVEY.md section 4 golden values, generated independently by the survey from the same reference (N=1e5)
    survey = {"grav3d_100000": {"targs.gidx": "4dabb0d8b604b1af", "srcs.x0": "fc790cc62bf536c0", "stree.nr": "d236dfc82c723b6a",
                                "stree.num": "453ba41b3a30f5f9", "treecode2.gflop": 29.555, "treecode3.gflop": 39.098, "treecode1.gflop": 25.484},
              "dtt_counts_1e5_t1.4": {"sltl": 52086, "sbtl": 3988, "sltb": 3989, "sbtb": 17570, "tlc": 782, "bpc": 780},
              "dtt_counts_1e6_t1.4": {"sltl": 612957, "sbtl": 57780, "sltb": 57774, "sbtb": 297403, "tlc": 7813}}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump({"cases": res, "survey": survey, "legacy": legacy()}, f, indent=1)
    print("wrote", len(res), "cases")
