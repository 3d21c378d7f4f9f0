import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    """the oracle restatement is test infrastructure: build it on demand (gcc only, a few seconds)"""
    from oracle.refapi import port_lib_path
    if not os.path.exists(port_lib_path()):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])


def bits_equal(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def rel_rms(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / ((b ** 2).sum() + 1e-300)))


def run_phases(s, theta, evals=True, tskip=None):
    """the drivers' phase sequence (ongrav3d.cpp:600-908) on any session object; returns every intermediate array"""
    out = {}
    s.init_driver()
    s.make_tree(0); p = s.parts(0); t = s.tree(0)
    out.update({"srcs.x": p["x"], "srcs.s": p["s"], "srcs.r": p["r"]})
    for k in ("num", "ioffset", "nc", "ns", "nr", "x", "s", "pr"):
        out["stree." + k] = t[k]
    out["stree.levels"] = t["levels"]; out["stree.numnodes"] = t["numnodes"]
    s.upward(0); e = s.parts(2)
    out.update({"eqsrcs.x": e["x"], "eqsrcs.s": e["s"], "eqsrcs.r": e["r"]})
    s.make_tree(1); s.refine(1); p = s.parts(1); t = s.tree(1)
    out.update({"targs.x": p["x"], "targs.gidx": p["gidx"]})
    for k in ("num", "ioffset", "nc", "ns", "nr", "x", "pr"):
        out["ttree." + k] = t[k]
    s.upward(1); out["eqtargs.x"] = s.parts(3)["x"]
    if evals:
        n = s.ntarg
        tsk = tskip or max(1, n // 400)
        s.zero_vels(); out["naive.flops"] = s.naive(tsk); out["naive.u"] = s.parts(1)["u"]
        for name in ("treecode1", "treecode2", "treecode3"):
            s.zero_vels(); out[name + ".flops"] = getattr(s, name)(theta); out[name + ".u"] = s.parts(1)["u"]
        if s.has_fastsumm:
            s.zero_vels(); s.fastsumm(theta); out["fastsumm.u"] = s.parts(1)["u"]; out["fastsumm.equ"] = s.parts(3)["u"]
    return out


def check_against_golden(out, case, hashfn):
    bad = []
    for k, v in case.items():
        if k not in out:
            continue
        got = out[k]
        if isinstance(got, np.ndarray):
            if "%016x" % hashfn(got) != v:
                bad.append(k)
        elif isinstance(v, float):
            if float(got) != v:
                bad.append("%s (%r != %r)" % (k, float(got), v))
        elif got != v:
            bad.append("%s (%r != %r)" % (k, got, v))
    return bad
