"""Phase-by-phase GPU-vs-oracle diagnostic (run on a B200: `python tools/gpu_diag.py [physics] [N] [theta]`).

Not a pytest file: it prints, for every phase of the hot path, whether the CUDA result is bit-identical to the
oracle's, and when a phase differs it re-runs the downstream phases from the ORACLE's state so that one GPU
session localises every independent fault. The pytest parity tests (tests/test_gpu_*.py) assert the same things.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.refapi import RefSession, PortSession, ref_available  # noqa: E402
from onbody_b200.api import GpuSession, ARITH_STRICT, ARITH_FAST  # noqa: E402


def same(a, b):
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def report(name, a, b, out):
    if a is None or b is None:
        return True
    ok = same(a, b)
    msg = "OK   " if ok else "DIFF "
    extra = ""
    if not ok and a.shape == b.shape:
        nd = int((a != b).sum())
        first = np.argwhere(a != b)[:3].tolist()
        af = a.astype(np.float64); bf = b.astype(np.float64)
        den = np.sqrt((bf ** 2).sum()) + 1e-300
        extra = " ndiff=%d/%d first=%s rel_rms=%.3e maxabs=%.3e" % (nd, a.size, first, np.sqrt(((af - bf) ** 2).sum()) / den, np.abs(af - bf).max())
    elif not ok:
        extra = " shape %s vs %s" % (a.shape, b.shape)
    line = "%s %-28s%s" % (msg, name, extra)
    print(line, flush=True)
    out.append(line)
    return ok


def cmp_parts(tag, g, o, out, keys=("x", "r", "s", "u", "gidx")):
    ok = True
    for k in keys:
        if g.get(k) is not None and o.get(k) is not None:
            ok &= report("%s.%s" % (tag, k), g[k], o[k], out)
    return ok


def cmp_tree(tag, g, o, out, eq=False):
    ok = True
    for k in ("num", "ioffset", "ns", "nc", "nr", "x", "s", "pr") + (("epoffset", "epnum") if eq else ()):
        ok &= report("%s.%s" % (tag, k), g[k], o[k], out)
    return ok


def rel_rms(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / ((b ** 2).sum() + 1e-300)))


def main():
    physics = sys.argv[1] if len(sys.argv) > 1 else "grav3d"
    N = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20000
    theta = float(sys.argv[3]) if len(sys.argv) > 3 else 1.2
    Oracle = RefSession if ref_available(physics) else PortSession
    print("oracle:", Oracle.__name__, "physics", physics, "N", N, "theta", theta, flush=True)
    out = []
    o = Oracle(physics, N, N)
    o.init_driver()
    init_s = o.parts(0); init_t = o.parts(1)

    g = GpuSession(physics, N, N, arith=ARITH_STRICT)
    g.init_driver()
    cmp_parts("init.srcs", g.parts(0), init_s, out, ("x", "r", "s"))

    # ---- source tree
    t0 = time.time(); g.make_tree(0); tg = time.time() - t0
    o.make_tree(0)
    ok_st = cmp_parts("srcs_tree", g.parts(0), o.parts(0), out, ("x", "r", "s"))
    ok_st &= cmp_tree("stree", g.tree(0), o.tree(0), out)
    print("   gpu tree %.1f ms (host wall %.1f ms) build stats %s" % (g.phase_ms("tree"), tg * 1e3, g.build_stats()), flush=True)
    if not ok_st:
        print("   -> reloading oracle source tree into the GPU session", flush=True)
        ps = o.parts(0); g.set_sources(ps["x"], ps["r"], ps["s"]); g.load_tree(0, o.tree(0))
    # ---- upward
    g.upward(0); o.upward(0)
    ok_up = cmp_parts("eqsrcs", g.parts(2), o.parts(2), out, ("x", "r", "s"))
    cmp_tree("stree+eq", g.tree(0), o.tree(0), out, eq=True)
    # ---- target tree + refine + upward
    g.make_tree(1); o.make_tree(1)
    ok_tt = cmp_parts("targs_tree", g.parts(1), o.parts(1), out, ("x", "r", "gidx"))
    ok_tt &= cmp_tree("ttree", g.tree(1), o.tree(1), out)
    g.refine(1); o.refine(1)
    ok_rf = cmp_parts("targs_refined", g.parts(1), o.parts(1), out, ("x", "r", "gidx"))
    print("   refine %.2f ms, tie sorts %d" % (g.phase_ms("refine"), g.build_stats()["tie_sorts"]), flush=True)
    if not (ok_tt and ok_rf):
        print("   -> reloading oracle target tree/order into the GPU session", flush=True)
        pt = o.parts(1); g.set_targets(pt["x"], pt["r"]); g.load_tree(1, o.tree(1))
    g.upward(1); o.upward(1)
    cmp_parts("eqtargs", g.parts(3), o.parts(3), out, ("x", "r"))

    if os.environ.get("DIAG_TREE_ONLY"):
        print("   upward %.2f ms" % g.phase_ms("upward"))
        bad = [l for l in out if not l.startswith("OK")]
        print("SUMMARY (tree only) %s N=%d: %d checks, %d not OK" % (physics, N, len(out), len(bad)))
        return
    # ---- evaluations, strict arithmetic: bit-exact expected
    tsk = max(1, N // 400)
    res = {}
    for name in ("naive", "treecode3", "fastsumm", "treecode2", "treecode1"):
        if name == "fastsumm" and not o.has_fastsumm:
            continue
        try:
            g.zero_vels(); o.zero_vels()
            if name == "naive":
                fg = g.naive(tsk); fo = o.naive(tsk)
            elif name == "fastsumm":
                g.fastsumm(theta); o.fastsumm(theta); fg = fo = 0.0
            else:
                fg = getattr(g, name)(theta); fo = getattr(o, name)(theta)
            ug = g.parts(1, ("u",))["u"]; uo = o.parts(1)["u"]
            report("strict.%s.u" % name, ug, uo, out)
            if fg != fo:
                line = "DIFF strict.%s.flops gpu %r oracle %r" % (name, fg, fo); print(line); out.append(line)
            print("   %s: gpu stats %s pairs %d eval %.2f ms (lists %.2f p2p %.2f down %.2f)" % (
                name, g.stats(), g.last_pairs(), g.phase_ms("eval"), g.phase_ms("lists"), g.phase_ms("p2p"), g.phase_ms("downward")), flush=True)
            if name == "fastsumm":
                report("strict.fastsumm.equ", g.parts(3, ("u",))["u"], o.parts(3)["u"], out)
            res[name] = uo.copy()
        except Exception as e:  # keep going: later phases are independent
            line = "FAIL strict.%s: %s" % (name, e); print(line, flush=True); out.append(line)
    if isinstance(o, PortSession):
        print("   oracle stats (last):", o.stats())

    # ---- evaluations, fast arithmetic: tolerance
    gf = GpuSession(physics, N, N, arith=ARITH_FAST)
    ps = o.parts(0); pt = o.parts(1)
    gf.set_sources(ps["x"], ps["r"], ps["s"]); gf.load_tree(0, o.tree(0)); gf.upward(0)
    gf.set_targets(pt["x"], pt["r"]); gf.load_tree(1, o.tree(1)); gf.upward(1)
    for name in ("naive", "treecode3", "fastsumm", "treecode2", "treecode1"):
        if name not in res:
            continue
        try:
            gf.zero_vels()
            if name == "naive":
                gf.naive(tsk)
            elif name == "fastsumm":
                gf.fastsumm(theta)
            else:
                getattr(gf, name)(theta)
            ug = gf.parts(1, ("u",))["u"]
            sel = slice(None, None, tsk) if name == "naive" else slice(None)
            d = rel_rms(ug[:, sel], res[name][:, sel])
            pairs = gf.last_pairs(); ms = gf.phase_ms("p2p")
            line = "%s fast.%s rel_rms_vs_oracle=%.3e  pairs=%d p2p=%.3f ms -> %.1f Gpairs/s" % (
                "OK   " if d < 2e-6 else "WARN ", name, d, pairs, ms, pairs / max(ms, 1e-9) * 1e-6)
            print(line, flush=True); out.append(line)
        except Exception as e:
            line = "FAIL fast.%s: %s" % (name, e); print(line, flush=True); out.append(line)
    print("fp32 peak (measured) %.2f TFLOP/s" % gf.measure_fp32_peak())
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/diag_%s_%d.txt" % (physics, N), "w") as f:
        f.write("\n".join(out) + "\n")
    bad = [l for l in out if not l.startswith("OK")]
    print("SUMMARY %s N=%d: %d checks, %d not OK" % (physics, N, len(out), len(bad)))


if __name__ == "__main__":
    main()
