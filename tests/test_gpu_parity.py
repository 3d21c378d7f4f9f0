"""GPU parity tests (run on a B200 with -m gpu). Everything goes through the C ABI (onbody_b200.api -> libonbody_b200.so).

Bars: tree order, node arrays, equivalent strengths, interaction counts and - in ARITH_STRICT - every output value are
BIT-EXACT against the oracle (the compiled reference when oracle/_ref travelled with the snapshot, else the CPU
restatement) and against tests/golden/golden.json. The product arithmetic (ARITH_FAST: rsqrt + FMA, reference
accumulation order) must stay within 1e-6 relative rms of the reference treecode result (north_star) - checked at
N = 3e4 for every physics and at N = 1e6 for the headline one - and reproduce the reference's own error against the direct sum.
"""
import numpy as np
import pytest

from conftest import bits_equal, rel_rms, run_phases, check_against_golden

pytestmark = pytest.mark.gpu


def _oracle_cls(physics):
    from oracle.refapi import RefSession, PortSession, ref_available
    return RefSession if ref_available(physics) else PortSession


def _gpu(physics, n, **kw):
    from onbody_b200.api import GpuSession
    return GpuSession(physics, n, n, **kw)


PHYS = ("grav3d", "vort3d", "vortgrad3d", "vort2d", "vort2dtr")


@pytest.mark.parametrize("idx", range(9))
def test_strict_matches_golden(golden, idx):
    from onbody_b200.api import ARITH_STRICT
    from oracle.refapi import fnv1a64
    case = golden["cases"][idx]
    g = _gpu(case["physics"], case["n"], arith=ARITH_STRICT)
    out = run_phases(g, case["theta"], evals=case["n"] <= 20000, tskip=case.get("tskip"))
    bad = check_against_golden(out, case, fnv1a64)
    assert not bad, "CUDA path differs from the reference's golden vectors: %s" % bad


@pytest.mark.parametrize("physics", PHYS)
def test_strict_bit_exact_against_oracle(physics):
    from onbody_b200.api import ARITH_STRICT
    n, theta = 30000, 1.3
    a = run_phases(_oracle_cls(physics)(physics, n, n), theta)
    b = run_phases(_gpu(physics, n, arith=ARITH_STRICT), theta)
    for k, v in a.items():
        if isinstance(v, np.ndarray):
            assert bits_equal(v, b[k]), k
        else:
            assert v == b[k], (k, v, b[k])


@pytest.mark.parametrize("physics", PHYS)
def test_fast_arithmetic_within_tolerance(physics):
    from onbody_b200.api import ARITH_FAST
    n, theta = 30000, 1.3
    a = run_phases(_oracle_cls(physics)(physics, n, n), theta)
    b = run_phases(_gpu(physics, n, arith=ARITH_FAST), theta)
    # ordering, node arrays, equivalent strengths and the interaction-list checksums do not depend on the pair arithmetic
    for k in ("srcs.x", "targs.gidx", "stree.nr", "stree.x", "eqsrcs.s", "treecode2.flops", "treecode3.flops", "treecode1.flops"):
        v = a[k]
        assert bits_equal(v, b[k]) if isinstance(v, np.ndarray) else v == b[k], k
    # north_star: 1e-6 relative deviation from the reference treecode; the 9 gradient outputs of vortgrad3d cancel more
    # strongly than velocities (their own float32 noise is larger): measured 3.6e-6 for its treecode1
    tol = 1e-6 if physics != "vortgrad3d" else 5e-6
    for k in ("treecode1.u", "treecode2.u", "treecode3.u", "fastsumm.u"):
        if k in a:
            assert rel_rms(b[k], a[k]) < tol, (k, rel_rms(b[k], a[k]))
    tsk = max(1, n // 400)
    assert rel_rms(b["naive.u"][:, ::tsk], a["naive.u"][:, ::tsk]) < 1e-5


def test_dtt_counts_and_error_match_reference_at_1e5(golden):
    """N=1e5, theta=1.4, o=4: the interaction counts SURVEY section 4 quotes, and the reference's rms error vs direct"""
    from onbody_b200.api import ARITH_FAST
    want = golden["survey"]["dtt_counts_1e5_t1.4"]
    n = 100000
    g = _gpu("grav3d", n, arith=ARITH_FAST)
    g.init_driver(); g.make_tree(0); g.upward(0); g.make_tree(1); g.refine(1); g.upward(1)
    g.zero_vels(); g.fastsumm(1.4)
    st = g.stats()
    for k, v in want.items():
        assert st[k] == v, (k, st[k], v)
    u = g.parts(1, ("u",))["u"][0]
    g.zero_vels(); g.naive(5); un = g.parts(1, ("u",))["u"][0]
    a, b = u[::5].astype(np.float64), un[::5].astype(np.float64)
    rms = np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum())
    assert abs(rms - 8.50221e-05) < 2e-6, rms          # reference prints 8.50221e-05 for this run
    # treecode checksums: the GFlop lines of the reference at theta=1.11111
    sv = golden["survey"]["grav3d_100000"]
    g.zero_vels(); assert abs(g.treecode3(1.11111) * 1e-9 - sv["treecode3.gflop"]) < 5e-4
    g.zero_vels(); assert abs(g.treecode2(1.11111) * 1e-9 - sv["treecode2.gflop"]) < 5e-4
    g.zero_vels(); assert abs(g.treecode1(1.11111) * 1e-9 - sv["treecode1.gflop"]) < 5e-4


def test_tree_order_1e6_against_oracle():
    """stall exits of the partial select and libstdc++ tie order inside leaves, at a size where both occur"""
    from onbody_b200.api import ARITH_FAST
    n = 1000000
    o = _oracle_cls("grav3d")("grav3d", n, n)
    g = _gpu("grav3d", n, arith=ARITH_FAST)
    for s in (o, g):
        s.init_driver(); s.make_tree(0); s.make_tree(1); s.refine(1)
    assert bits_equal(o.parts(0)["x"], g.parts(0, ("x",))["x"])
    assert bits_equal(o.parts(1)["gidx"], g.parts(1, ("gidx",))["gidx"])
    to, tg = o.tree(0), g.tree(0)
    for k in ("num", "ioffset", "nc", "ns", "nr", "x", "s", "pr"):
        assert bits_equal(to[k], tg[k]), k
    bs = g.build_stats()
    assert bs["selects"] == 7812 and bs["stalls"] == 2 and bs["tie_sorts"] > 0


def test_full_size_properties_1e7():
    """BASELINE configs[1] size (N=1e7): properties that need no oracle run."""
    from onbody_b200.api import ARITH_FAST, driver_inputs
    n = 10000000
    x, r, s = driver_inputs("grav3d", n, True)
    g = _gpu("grav3d", n, arith=ARITH_FAST)
    g.set_sources(x, r, s); g.set_targets(x, r)
    g.make_tree(0); g.upward(0); g.make_tree(1); g.refine(1); g.upward(1)
    p = g.parts(1, ("x", "gidx"))
    gi = p["gidx"].astype(np.int64)
    assert np.array_equal(np.sort(gi), np.arange(n))                     # a permutation
    assert np.array_equal(p["x"], x[:, gi])                              # ... of the input particles
    t = g.tree(1)
    assert t["levels"] == 18 and t["numnodes"] == 262144                 # Tree.hpp sizing
    leaf = (t["num"] > 0) & (t["num"] <= 128)
    assert int(leaf.sum()) == 78125 and int(t["num"][leaf].sum()) == n   # leaves partition the particles
    # every leaf's tight box (nc +- ns/2) contains its particles: check a sample of leaves
    ids = np.nonzero(leaf)[0][::997]
    for i in ids:
        a, m = int(t["ioffset"][i]), int(t["num"][i])
        for d in range(3):
            seg = p["x"][d, a:a + m]
            assert seg.min() >= t["nc"][d, i] - 0.5 * t["ns"][d, i] - 1e-6 and seg.max() <= t["nc"][d, i] + 0.5 * t["ns"][d, i] + 1e-6
    # upward pass conserves total charge at every level: root equivalent strengths sum to the total strength
    es = g.parts(2, ("s",))["s"][0].astype(np.float64)
    root = es[128:128 + 125].sum()
    assert abs(root - s[0].astype(np.float64).sum()) < 1e-6 * np.abs(s[0]).astype(np.float64).sum()
    # dual tree: error against the direct sum at the reference's level, linear in the strengths
    g.zero_vels(); g.fastsumm(1.4); u = g.parts(1, ("u",))["u"]
    tsk = 50000
    g.zero_vels(); g.naive(tsk); un = g.parts(1, ("u",))["u"]
    a, b = u[0, ::tsk].astype(np.float64), un[0, ::tsk].astype(np.float64)
    rms = np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum())
    assert rms < 2.5e-4, rms                                             # reference: 1.26e-4 at this N
    st = g.stats()
    assert st["tlc"] == 78125
    g2 = _gpu("grav3d", n, arith=ARITH_FAST)
    g2.set_sources(x, r, 2.0 * s); g2.set_targets(x, r)
    g2.make_tree(0); g2.upward(0); g2.make_tree(1); g2.refine(1); g2.upward(1)
    g2.zero_vels(); g2.fastsumm(1.4)
    u2 = g2.parts(1, ("u",))["u"]
    assert np.array_equal(u2, 2.0 * u)                                   # exact: scaling by 2 commutes with every rounding


@pytest.mark.parametrize("arith", [0, 1])
def test_target_shards_reproduce_the_unsharded_result(arith):
    """multi-GPU path on one GPU: shard 0/3, 1/3, 2/3 evaluated one after the other == the 1-GPU result, bit for bit"""
    n, theta = 50000, 1.4
    def run(rank, world):
        g = _gpu("grav3d", n, arith=arith)
        g.set_shard(rank, world)
        g.init_driver(); g.make_tree(0); g.upward(0); g.make_tree(1); g.refine(1); g.upward(1)
        g.zero_vels(); g.fastsumm(theta); uf = g.parts(1, ("u",))["u"]
        g.zero_vels(); g.treecode3(theta); u3 = g.parts(1, ("u",))["u"]
        g.zero_vels(); g.treecode2(theta); u2 = g.parts(1, ("u",))["u"]
        return uf, u3, u2
    from onbody_b200.api import shard_range_for
    full = run(0, 1)
    acc = [np.zeros_like(full[0]) for _ in range(3)]
    for rk in range(3):
        lo, hi = shard_range_for(n, 128, rk, 3)                 # equal leaf-aligned chunks (csrc/plan.cu)
        part = run(rk, 3)
        for k in range(3):
            acc[k][:, lo:hi] = part[k][:, lo:hi]
            outside = np.ones(n, bool); outside[lo:hi] = False
            assert not part[k][:, outside].any() or k == 0      # boxwise and pointwise touch only their own leaves
    for k in range(3):
        assert bits_equal(acc[k], full[k]), k


def test_c_abi_results_original_order():
    """+= into caller arrays in the caller's original target order (interface3dvortgrads.cpp:384-395)"""
    n = 20000
    o = _oracle_cls("vortgrad3d")("vortgrad3d", n, n); o.init_driver()
    ps, pt = o.parts(0), o.parts(1)
    from onbody_b200.api import ARITH_STRICT
    g = _gpu("vortgrad3d", n, arith=ARITH_STRICT)
    g.set_sources(ps["x"], ps["r"], ps["s"]); g.set_targets(pt["x"], pt["r"])
    g.make_tree(0); g.upward(0); g.make_tree(1)
    g.zero_vels(); g.treecode3(1.5)
    out = np.ones((12, n), np.float32)
    g.add_results_original_order(out)
    o.make_tree(0); o.upward(0); o.make_tree(1); o.zero_vels(); o.treecode3(1.5)
    po = o.parts(1)
    want = np.ones((12, n), np.float32)
    want[:, po["gidx"].astype(np.int64)] += po["u"]
    assert bits_equal(out, want)


def test_range_restricted_builds_reproduce_the_full_build():
    """multi-GPU source build emulated on one GPU: three contexts each sort one third of the source tree, exchange their
    plane ranges (device copies standing in for the NCCL all-gather), finish bottom-up, and must then hold exactly the
    arrays of a full build; sharded target builds + dual tree must reproduce the unsharded outputs bit for bit."""
    import torch
    n, theta, world = 70000, 1.4, 3
    full = _gpu("grav3d", n)
    full.init_driver(); full.make_tree(0); full.upward(0); full.make_tree(1); full.refine(1); full.upward(1)
    full.zero_vels(); full.fastsumm(theta)
    want_src = full.parts(0); want_eq = full.parts(2, ("x", "s")); want_tree = full.tree(0)
    want_u = full.parts(1, ("u", "gidx"))
    ranks = []
    for rk in range(world):
        g = _gpu("grav3d", n); g.set_shard(rk, world); g.init_driver()
        lo, hi = g.shard_particle_range(n, rk, world)
        g.make_tree_range(0, lo, hi)
        ranks.append((g, lo, hi))
    torch.cuda.synchronize()
    for g, lo, hi in ranks:                      # "all-gather": everyone receives everyone else's range of every plane
        for f in g.source_fields():
            dst = g.plane_tensor(0, f, n)
            for h, a, b in ranks:
                if h is not g:
                    dst[a:b].copy_(h.plane_tensor(0, f, n)[a:b])
    torch.cuda.synchronize()
    got_u = np.zeros_like(want_u["u"]); got_g = np.zeros_like(want_u["gidx"])
    for g, lo, hi in ranks:
        g.finish_tree(0); g.upward(0)
        ps = g.parts(0)
        for k in ("x", "r", "s"):
            assert bits_equal(ps[k], want_src[k]), k
        t = g.tree(0)
        for k in ("num", "ioffset", "nc", "ns", "nr", "x", "s", "pr"):
            assert bits_equal(t[k], want_tree[k]), k
        assert bits_equal(g.parts(2, ("s",))["s"], want_eq["s"])
        g.make_tree_range(1, lo, hi)
    torch.cuda.synchronize()
    for g, lo, hi in ranks:                      # target coordinates too: ancestor centres enter the dual-tree MAC
        for f in (0, 1, 2, 3):
            dst = g.plane_tensor(1, f, n)
            for h, a, b in ranks:
                if h is not g:
                    dst[a:b].copy_(h.plane_tensor(1, f, n)[a:b])
    torch.cuda.synchronize()
    for g, lo, hi in ranks:
        g.finish_tree(1); g.set_build_range(1, lo, hi); g.refine(1); g.upward(1)
        g.zero_vels(); g.fastsumm(theta)
        p = g.parts(1, ("u", "gidx"))
        got_u[:, lo:hi] = p["u"][:, lo:hi]; got_g[lo:hi] = p["gidx"][lo:hi]
    assert bits_equal(got_g, want_u["gidx"])
    assert bits_equal(got_u, want_u["u"])


def _run_generic(s, xs, rs, ss, xt, rt, theta, tsk):
    """phase sequence with caller-supplied, DIFFERENT source and target clouds"""
    out = {}
    s.set_sources(xs, rs, ss); s.set_targets(xt, rt)
    s.make_tree(0); s.upward(0); s.make_tree(1); s.refine(1); s.upward(1)
    out["srcs.x"] = s.parts(0)["x"]; out["targs.gidx"] = s.parts(1)["gidx"]; out["eqsrcs.s"] = s.parts(2)["s"]
    s.zero_vels(); s.naive(tsk); out["naive.u"] = s.parts(1)["u"]
    for name in ("treecode1", "treecode2", "treecode3"):
        s.zero_vels(); out[name + ".flops"] = getattr(s, name)(theta); out[name + ".u"] = s.parts(1)["u"]
    if s.has_fastsumm:
        s.zero_vels(); s.fastsumm(theta); out["fastsumm.u"] = s.parts(1)["u"]
    return out


@pytest.mark.parametrize("physics,ns,nt,block,order", [
    ("grav3d", 9000, 5000, 128, 4), ("grav3d", 5000, 9000, 64, 3), ("grav3d", 7001, 7003, 50, 2), ("grav3d", 3000, 300, 16, 1),
    ("vort3d", 4000, 6000, 100, 4), ("vortgrad3d", 5000, 3000, 128, 3), ("vort2d", 6000, 8000, 128, 10), ("vort2dtr", 8000, 6000, 30, 6),
    ("grav3d", 1, 1, 128, 4), ("grav3d", 2, 3, 128, 4), ("grav3d", 128, 129, 128, 4), ("vort2dtr", 257, 100, 128, 4)])
def test_strict_bit_exact_unequal_clouds_blocks_orders(physics, ns, nt, block, order):
    """sources != targets (count and positions), block sizes other than 128 (incl. one that is not a multiple of 4, which
    exercises the unaligned-tile path of the pair kernel), orders 1..10, and degenerate sizes (single particle, single leaf)"""
    from onbody_b200.api import GpuSession, ARITH_STRICT, _DIMS
    from oracle.refapi import PortSession
    PD, SD, OD = _DIMS[physics]
    rng = np.random.RandomState(ns * 7 + nt)
    f = lambda *sh: np.ascontiguousarray(rng.uniform(-1, 1, sh).astype(np.float32))
    xs, ss, xt = f(PD, ns), (f(SD, ns) / ns).astype(np.float32), (0.9 * f(PD, nt) + 0.05).astype(np.float32)
    rs = (np.full(ns, max(ns, 2) ** (-1.0 / PD)) * (0.5 + rng.rand(ns))).astype(np.float32)
    rt = (np.full(nt, 0.3 * max(nt, 2) ** (-1.0 / PD))).astype(np.float32)
    o = PortSession(physics, ns, nt, block=block, order=order, eq_block=128)
    g = GpuSession(physics, ns, nt, block=block, order=order, arith=ARITH_STRICT)
    theta, tsk = 1.25, max(1, nt // 300)
    a = _run_generic(o, xs, rs, ss, xt, rt, theta, tsk)
    b = _run_generic(g, xs, rs, ss, xt, rt, theta, tsk)
    for k, v in a.items():
        if isinstance(v, np.ndarray):
            assert bits_equal(v, b[k]), k
        else:
            assert v == b[k], (k, v, b[k])


def test_concurrent_tree_builds_equal_sequential():
    """onb_make_trees (two streams) must give exactly the arrays of two onb_make_tree calls"""
    n = 300000
    a = _gpu("grav3d", n); a.init_driver(); a.make_tree(0); a.make_tree(1)
    b = _gpu("grav3d", n); b.init_driver(); b.make_trees()
    for which in (0, 1):
        pa, pb = a.parts(which), b.parts(which)
        for k in ("x", "r", "s", "gidx"):
            if pa[k] is not None:
                assert bits_equal(pa[k], pb[k]), (which, k)
        ta, tb = a.tree(which), b.tree(which)
        for k in ("num", "ioffset", "nc", "ns", "nr", "x", "s", "pr"):
            assert bits_equal(ta[k], tb[k]), (which, k)


def test_fastsumm_needs_no_history_and_survives_overflow(monkeypatch):
    """the dual-tree lists live in a device-side bump pool: every evaluation - first or repeated, same or new theta - is one
    pass without host round trips; a pool (or per-warp FIFO) that turns out too small is grown and the pass redone, with
    identical results"""
    n = 60000
    g = _gpu("grav3d", n)
    g.init_driver(); g.make_trees(); g.upward(0); g.refine(1); g.upward(1)
    res = {}
    for i, th in enumerate((1.4, 1.4, 1.1, 1.1, 1.4, 0.8)):
        g.zero_vels(); g.fastsumm(th)
        assert g.phase_ms("dtt_attempts") == 1.0, (i, th)          # the default pool fits: no redo, not even on the very first call
        u = g.parts(1, ("u",))["u"]; st = g.stats(); pr = g.last_pairs()
        if th in res:
            assert bits_equal(u, res[th][0]) and st == res[th][1] and pr == res[th][2], (i, th)
        else:
            res[th] = (u, st, pr)
    monkeypatch.setenv("ONB_DTT_POOL_INIT", "2000")                # a pool of 2000 entries and a FIFO of 8: both must overflow and grow
    monkeypatch.setenv("ONB_DTT_QCAP_INIT", "8")
    h = _gpu("grav3d", n)
    h.init_driver(); h.make_tree(0); h.upward(0); h.make_tree(1); h.refine(1); h.upward(1)
    h.zero_vels(); h.fastsumm(1.1)
    assert h.phase_ms("dtt_attempts") > 2.0
    assert bits_equal(h.parts(1, ("u",))["u"], res[1.1][0]) and h.stats() == res[1.1][1]
    h.zero_vels(); h.fastsumm(1.1)
    assert h.phase_ms("dtt_attempts") == 1.0 and bits_equal(h.parts(1, ("u",))["u"], res[1.1][0])


def _run_legacy(s, theta):
    out = {}
    s.init_driver(); s.make_tree(0); s.refine(0); s.upward(0); s.make_tree(1)
    e, t = s.parts(2), s.tree(0)
    out.update({"srcs.x": s.parts(0)["x"], "eqsrcs.x": e["x"], "eqsrcs.r": e["r"], "eqsrcs.s": e["s"], "stree.epnum": t["epnum"]})
    for name in ("treecode2", "treecode3"):
        s.zero_vels(); out[name + ".flops"] = getattr(s, name)(theta); out[name + ".u"] = s.parts(1)["u"]
    return out


@pytest.mark.parametrize("physics,n,block", [("grav3d", 20000, 128), ("vort3d", 9000, 128), ("vort2d", 7001, 128), ("vortgrad3d", 5000, 128), ("grav3d", 6000, 64),
                                             ("vort2dtr", 4001, 32), ("grav3d", 300, 128), ("grav3d", 129, 128)])
def test_legacy_equivalents_strict_bit_exact(physics, n, block):
    """-o omitted (order = -1, the drivers' default): refineTree(srcs) + pair-merge equivalents (barneshut.hpp:946-1061)
    and treecode2/3 over them, every array and every result bit-identical to the oracle"""
    from onbody_b200.api import GpuSession, ARITH_STRICT
    from oracle.refapi import PortSession
    a = _run_legacy(PortSession(physics, n, n, block=block, order=-1, eq_block=block), 1.2)
    b = _run_legacy(GpuSession(physics, n, n, block=block, order=-1, arith=ARITH_STRICT), 1.2)
    for k in ("srcs.x", "stree.epnum", "treecode2.u", "treecode3.u"):
        assert bits_equal(a[k], b[k]), k
    # equivalent arrays: the reference strides node blocks by blockSize, the GPU build by 128 - compare node by node
    ep = a["stree.epnum"]
    for node in np.nonzero(ep)[0][:4000]:
        cnt = int(ep[node])
        for key in ("eqsrcs.x", "eqsrcs.r", "eqsrcs.s"):
            ra = a[key][..., node * block: node * block + cnt]; rb = b[key][..., node * 128: node * 128 + cnt]
            assert bits_equal(np.ascontiguousarray(ra), np.ascontiguousarray(rb)), (key, node)
    assert a["treecode2.flops"] == b["treecode2.flops"] and a["treecode3.flops"] == b["treecode3.flops"]


def test_legacy_equivalents_fast_arithmetic_and_refusals():
    from onbody_b200.api import GpuSession, ARITH_FAST, OnbodyError
    from oracle.refapi import PortSession
    n = 30000
    a = _run_legacy(PortSession("grav3d", n, n, order=-1), 1.2)
    g = GpuSession("grav3d", n, n, order=-1, arith=ARITH_FAST)
    b = _run_legacy(g, 1.2)
    for k in ("treecode2.u", "treecode3.u"):
        assert rel_rms(b[k], a[k]) < 1e-6, (k, rel_rms(b[k], a[k]))
    with pytest.raises(OnbodyError):
        g.fastsumm(1.2)          # no target equivalents in the legacy mode (barneshut.hpp:953)
    with pytest.raises(OnbodyError):
        g.upward(1)


def test_full_size_invariants_1e7():
    """BASELINE configs[1] at its full size (N = 1e7, -t=1.4 -o=4 -b=128), where the CPU oracle would take minutes:
    size-independent properties of every stage, plus the numbers the reference run of the same command prints"""
    from onbody_b200.api import GpuSession, ARITH_FAST
    n, block = 10000000, 128
    g = GpuSession("grav3d", n, n, arith=ARITH_FAST)
    g.init_driver(); g.make_trees()
    bs = g.build_stats()
    assert bs["selects"] == 78124 and bs["stalls"] == 39 and bs["passes"] == 338547      # the reference's own pass statistics at N = 1e7
    ts, tt = g.tree(0), g.tree(1)
    for t in (ts, tt):
        num, io = t["num"].astype(np.int64), t["ioffset"].astype(np.int64)
        assert num[1] == n and io[1] == 0
        nonleaf = np.nonzero(num > block)[0]
        assert np.all(num[2 * nonleaf] + num[2 * nonleaf + 1] == num[nonleaf])            # children partition the parent
        assert np.all(io[2 * nonleaf] == io[nonleaf]) and np.all(io[2 * nonleaf + 1] == io[nonleaf] + num[2 * nonleaf])
        assert np.all(num[2 * nonleaf] == block * 2 ** np.floor(np.log2((num[nonleaf] - 1) // block)).astype(np.int64))   # barneshut.hpp:663
        leaves = np.nonzero((num > 0) & (num <= block))[0]
        assert leaves.size == 78125 and num[leaves].sum() == n
    p = g.parts(0, ("x", "s"))
    # every particle lies inside the tight box of its leaf and of the root; the split axis separates the children
    num, io = ts["num"].astype(np.int64), ts["ioffset"].astype(np.int64)
    for node in (1, 2, 3, 77, 1000, 70001, 131072, 150000, 200000):
        if num[node] == 0:
            continue
        seg = p["x"][:, io[node]: io[node] + num[node]]
        lo, hi = seg.min(axis=1), seg.max(axis=1)
        assert np.array_equal((hi - lo).astype(np.float32), ts["ns"][:, node]) and np.array_equal((0.5 * (hi + lo)).astype(np.float32), ts["nc"][:, node])
        if num[node] > block:
            ax = int(np.argmax(ts["ns"][:, node])); m = num[2 * node]
            assert seg[ax, :m].max() <= seg[ax, m:].min()
    # the reordering is a permutation of the input (multiset of strengths and of every coordinate is preserved)
    from onbody_b200.api import driver_inputs
    xi, ri, si = driver_inputs("grav3d", n, True)
    assert np.array_equal(np.sort(p["s"][0]), np.sort(si[0])) and np.array_equal(np.sort(p["x"][2]), np.sort(xi[2]))
    gi = g.parts(1, ("gidx",))["gidx"].astype(np.int64)
    assert np.array_equal(np.sort(gi), np.arange(n))                                      # targets: gidx is a permutation ...
    tx = g.parts(1, ("x",))["x"]
    assert np.array_equal(tx[1], xi[1][gi])                                               # ... that maps tree order back to the input
    # upward pass: barycentric weights sum to one, so every node's equivalent strengths sum to its particles' strengths
    g.upward(0); g.refine(1); g.upward(1)
    # the hashes SURVEY.md section 4 obtained independently from the reference at N = 1e7 (FNV-1a-64 over the raw bytes):
    # final target order, target coordinates, every source-node array, and the strict (-O2) build's intra-leaf source order
    from oracle.refapi import fnv1a64
    h = lambda a: "%016x" % fnv1a64(np.ascontiguousarray(a))
    pt = g.parts(1, ("gidx", "x"))
    assert h(pt["gidx"]) == "957a1ab3437a7eeb" and [int(v) for v in pt["gidx"][:8]] == [8089804, 7138819, 7371734, 6685117, 5734885, 1392952, 2924730, 1050990]
    assert h(pt["x"][0]) == "b647a1c4491c5fe3"
    assert ts["levels"] == 18 and ts["numnodes"] == 262144
    assert h(ts["nr"]) == "54cb79fc9521bd9d" and h(ts["nc"][0]) == "1b16678759bcf79e" and h(ts["x"][0]) == "8951e8b417150f7e" and h(ts["num"]) == "822a73b6b72f92bf"
    assert h(p["x"][0]) == "f1fd8ca55ae2d837"
    es = g.parts(2, ("s",))["s"][0].astype(np.float64).reshape(-1, 128).sum(axis=1)
    tot = float(si[0].astype(np.float64).sum())
    assert abs(es[1] - tot) <= 2e-5 * np.abs(si[0]).astype(np.float64).sum()
    for node in (2, 3, 500, 40000):
        want = p["s"][0][io[node]: io[node] + num[node]].astype(np.float64).sum()
        assert abs(es[node] - want) <= 2e-5 * np.abs(p["s"][0][io[node]: io[node] + num[node]]).astype(np.float64).sum()
    # dual tree: exact pair count of the lists (tests/golden/dtt_pairs.json, measured once and pinned) and the reference's
    # accuracy against the direct sum on its own sample (every 5000th target, ongrav3d.cpp:556-561)
    g.zero_vels(); g.fastsumm(1.4)
    import json, os
    from conftest import ROOT
    with open(os.path.join(ROOT, "tests", "golden", "dtt_pairs.json")) as f:
        assert g.last_pairs() == int(json.load(f)[str(n)])
    u = g.parts(1, ("u",))["u"]
    g.zero_vels(); g.naive(5000); un = g.parts(1, ("u",))["u"]
    a, b = u[:, ::5000].astype(np.float64), un[:, ::5000].astype(np.float64)
    rms = np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum())
    assert 3e-5 < rms < 2e-4, rms              # the reference's README: ~1e-4 at -t=1.4 -o=4


def test_prepare_eval_equals_separate_calls():
    """onb_prepare_eval (source side and target side overlapped on two streams) against the separate phase calls, plain and
    in the multi-GPU variant (node arrays completed bottom-up, refinement restricted to a range)"""
    n = 300000
    a = _gpu("grav3d", n); a.init_driver(); a.make_trees(); a.upward(0); a.refine(1); a.upward(1); a.zero_vels(); a.fastsumm(1.4)
    b = _gpu("grav3d", n); b.init_driver(); b.make_trees(); b.prepare_eval(); b.zero_vels(); b.fastsumm(1.4)
    lo, hi = a.shard_particle_range(n, 1, 3)
    c = _gpu("grav3d", n); c.init_driver(); c.make_trees(); c.finish_tree(0); c.upward(0); c.finish_tree(1); c.set_build_range(1, lo, hi); c.refine(1); c.upward(1)
    d = _gpu("grav3d", n); d.init_driver(); d.make_trees(); d.prepare_eval(True, lo, hi)
    for x, y, evald in ((a, b, True), (c, d, False)):
        for which, keys in ((1, ("x", "gidx") + (("u",) if evald else ())), (2, ("x", "r", "s")), (3, ("x",))):
            px, py = x.parts(which, keys), y.parts(which, keys)
            for k in keys:
                assert bits_equal(px[k], py[k]), (which, k, evald)
        assert x.build_stats() == y.build_stats()
        for which in (0, 1):
            tx, ty = x.tree(which), y.tree(which)
            for k in ("num", "ioffset", "nc", "ns", "nr", "x", "pr"):
                assert bits_equal(tx[k], ty[k]), (which, k)


def test_pivot_mode_1_matches_the_fast_math_reference_build():
    """Tier B against the reference built with ITS OWN flags (-O3 -ffast-math, FMA target): g++ contracts the pivot blend of
    barneshut.hpp:538-540 into FMAs there, which changes the intra-leaf source order of a few particles from N ~ 1e6 up.
    onb_set_pivot_mode(1) evaluates the pivot with exactly those contractions: the source order must then equal the fast
    build's (oracle/_ref/fast, -march=x86-64-v3) and the hashes SURVEY.md section 4 obtained independently from a
    -march=native build at N = 1e6 and N = 1e7."""
    from onbody_b200.api import load_library, ARITH_FAST
    from oracle.refapi import RefSession, ref_available, fnv1a64
    L = load_library()
    h = lambda a: "%016x" % fnv1a64(np.ascontiguousarray(a))
    try:
        L.onb_set_pivot_mode(1)
        n = 1000000
        g = _gpu("grav3d", n, arith=ARITH_FAST); g.init_driver(); g.make_tree(0); g.make_tree(1); g.refine(1)
        sx = g.parts(0, ("x", "s")); tg = g.parts(1, ("gidx",))["gidx"]
        assert h(sx["x"][0]) == "9f9b637700c080e1"                       # SURVEY section 4, fast build, N = 1e6
        assert h(tg) == "502bae5161063bf7"                                # the refined target order is the same in both builds
        if ref_available("grav3d", "fast"):
            o = RefSession("grav3d", n, n, build="fast"); o.init_driver(); o.make_tree(0)
            po = o.parts(0)
            assert bits_equal(po["x"], sx["x"]) and bits_equal(po["s"], sx["s"])
            o.close()
        g.close()
        L.onb_set_pivot_mode(0)
        g = _gpu("grav3d", n, arith=ARITH_FAST); g.init_driver(); g.make_tree(0)
        assert h(g.parts(0, ("x",))["x"][0]) == "5b0b68d776e49dc5"       # ... and mode 0 is the strict (-O2, no contraction) build
        g.close()
        L.onb_set_pivot_mode(1)
        n = 10000000
        g = _gpu("grav3d", n, arith=ARITH_FAST); g.init_driver(); g.make_tree(0)
        assert h(g.parts(0, ("x",))["x"][0]) == "d55259c3112840c7"       # SURVEY section 4, fast build, N = 1e7
        g.close()
    finally:
        L.onb_set_pivot_mode(0)


def test_fast_arithmetic_deviation_at_1e6():
    """north_star: 1e-6 relative deviation from the reference treecode. The product arithmetic (rsqrt.approx + FMA, reference
    accumulation order) against the compiled strict reference at N = 1e6: boxwise treecode (parallel-safe in the reference)
    and the serial dual tree (the reference's OpenMP dual tree has a data race, README.md:200)."""
    from onbody_b200.api import ARITH_FAST
    from oracle.refapi import RefSession, ref_available
    if not ref_available("grav3d"):
        pytest.skip("compiled reference (oracle/_ref) not present")
    n = 1000000
    o = RefSession("grav3d", n, n); o.init_driver()
    o.make_tree(0); o.upward(0); o.make_tree(1); o.refine(1); o.upward(1)
    g = _gpu("grav3d", n, arith=ARITH_FAST); g.init_driver(); g.make_trees(); g.prepare_eval()
    o.zero_vels(); o.treecode3(1.11111); u3 = o.parts(1)["u"]
    g.zero_vels(); g.treecode3(1.11111)
    d3 = rel_rms(g.parts(1, ("u",))["u"], u3)
    o.zero_vels(); o.fastsumm(1.4); uf = o.parts(1)["u"]
    g.zero_vels(); g.fastsumm(1.4)
    df = rel_rms(g.parts(1, ("u",))["u"], uf)
    print("fast arithmetic vs strict reference at N=1e6: boxwise %.3e dual tree %.3e" % (d3, df))
    assert d3 < 1e-6 and df < 1e-6, (d3, df)


@pytest.mark.parametrize("physics", ["grav3d", "vort3d"])
def test_accum_double_matches_the_reference_built_with_accum_double(physics):
    """the reference's compile-time ACCUM = double variant (ongrav3d.cpp:7-8, README.md:107-112: fp32 storage and
    pair arithmetic, fp64 accumulation). onb_set_accum(1) against the reference templates instantiated with A = double
    (oracle/_ref/strict/libref_<physics>_a64.so): every method bit-exact in the strict arithmetic - outputs compared as
    fp64 -, the product arithmetic within 1e-6, and the dual tree at a tight theta below the float32 error floor."""
    from onbody_b200.api import GpuSession, ARITH_STRICT, ARITH_FAST
    from oracle.refapi import RefSession, ref_available
    if not ref_available(physics + "_a64"):
        pytest.skip("ACCUM = double build of the reference (oracle/_ref) not present")
    n, theta = 20000, 1.3

    def run(s):
        out = {}
        s.init_driver(); s.make_tree(0); s.upward(0); s.make_tree(1); s.refine(1); s.upward(1)
        s.zero_vels(); s.naive(1); out["naive"] = s.results_f64(1)
        for name in ("treecode1", "treecode2", "treecode3"):
            s.zero_vels(); getattr(s, name)(theta); out[name] = s.results_f64(1)
        s.zero_vels(); s.fastsumm(theta); out["fastsumm"] = s.results_f64(1); out["fastsumm.eq"] = s.results_f64(3)
        s.zero_vels(); s.fastsumm(3.0); out["fastsumm.tight"] = s.results_f64(1)
        return out
    want = run(RefSession(physics, n, n, accum64=True))
    strict = run(GpuSession(physics, n, n, arith=ARITH_STRICT, accum64=True))
    for k, v in want.items():
        assert bits_equal(v, strict[k]), k
    g = GpuSession(physics, n, n, arith=ARITH_FAST, accum64=True)
    fast = run(g)
    for k, v in want.items():
        assert rel_rms(fast[k], v) < 1e-6, (k, rel_rms(fast[k], v))
    # the float view of the fp64 outputs is their correctly rounded copy
    assert bits_equal(g.parts(1, ("u",))["u"], fast["fastsumm.tight"].astype(np.float32))
    # README.md:107-112: fp64 accumulation takes the dual tree below the float32 floor (6e-6 at float, a few 1e-7 with ACCUM = double)
    err64 = rel_rms(fast["fastsumm.tight"], want["naive"])
    f32 = GpuSession(physics, n, n, arith=ARITH_FAST)
    f32.init_driver(); f32.make_trees(); f32.prepare_eval(); f32.zero_vels(); f32.fastsumm(3.0)
    err32 = rel_rms(f32.parts(1, ("u",))["u"], want["naive"])
    print("%s dual tree at -t=3 vs fp64-accumulated direct sum: ACCUM=float %.2e, ACCUM=double %.2e" % (physics, err32, err64))
    assert err64 < 0.5 * err32 and err64 < 1e-6


def test_first_call_of_prepare_eval_at_full_size():
    """on the very first onb_prepare_eval of a context the equivalent-target planes are allocated while the
    target chain runs on the second stream; their zero fill must be ordered on THAT stream or it can land after k_upward has
    written the points. N = 1e7 makes the source chain long enough to lose such a race; compare with the separate calls."""
    n = 10000000
    a = _gpu("grav3d", n); a.init_driver(); a.make_trees(); a.prepare_eval()          # first call, nothing allocated yet
    ea = a.parts(3, ("x",))["x"]; a.zero_vels(); a.fastsumm(1.4); ua = a.parts(1, ("u",))["u"]; a.close()
    b = _gpu("grav3d", n); b.init_driver(); b.make_trees(); b.upward(0); b.refine(1); b.upward(1)
    eb = b.parts(3, ("x",))["x"]; b.zero_vels(); b.fastsumm(1.4); ub = b.parts(1, ("u",))["u"]; b.close()
    assert np.isfinite(ua).all() and bits_equal(ea, eb) and bits_equal(ua, ub)


@pytest.mark.parametrize("physics,n,tol", [("vort3d", 200000, 1e-6), ("vortgrad3d", 100000, 5e-6)])
def test_other_physics_fast_arithmetic_at_larger_n(physics, n, tol):
    """BASELINE configs[2]/[3] (onvort3d, onvortgrad3d): product arithmetic against the compiled strict reference beyond the
    3e4 of the per-physics test - boxwise treecode for both, dual tree where the reference has one"""
    from onbody_b200.api import ARITH_FAST
    from oracle.refapi import RefSession, ref_available
    if not ref_available(physics):
        pytest.skip("compiled reference (oracle/_ref) not present")
    o = RefSession(physics, n, n); o.init_driver()
    o.make_tree(0); o.upward(0); o.make_tree(1); o.refine(1)
    g = _gpu(physics, n, arith=ARITH_FAST); g.init_driver(); g.make_trees(); g.upward(0); g.refine(1)
    o.zero_vels(); f_ref = o.treecode3(1.2); u3 = o.parts(1)["u"]
    g.zero_vels(); f_gpu = g.treecode3(1.2)
    assert f_gpu == f_ref                                                   # same interaction lists (flop checksum)
    d3 = rel_rms(g.parts(1, ("u",))["u"], u3)
    assert d3 < tol, d3
    if o.has_fastsumm:
        o.upward(1); g.upward(1)
        o.zero_vels(); o.fastsumm(1.4); uf = o.parts(1)["u"]
        g.zero_vels(); g.fastsumm(1.4)
        df = rel_rms(g.parts(1, ("u",))["u"], uf)
        assert df < tol, df
