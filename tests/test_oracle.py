"""CPU tests of the oracle: the restatement (oracle/port) against the golden vectors generated from the unmodified
reference, and - when the compiled reference is present - against the reference itself, array for array."""
import numpy as np
import pytest

from conftest import bits_equal, run_phases, check_against_golden
from oracle.refapi import PortSession, RefSession, ref_available, fnv1a64, PHYSICS


@pytest.mark.parametrize("idx", range(9))
def test_port_matches_golden(golden, idx):
    case = golden["cases"][idx]
    s = PortSession(case["physics"], case["n"], case["n"])
    out = run_phases(s, case["theta"], evals=case["n"] <= 20000, tskip=case.get("tskip"))
    bad = check_against_golden(out, case, fnv1a64)
    assert not bad, "port differs from the reference's golden vectors: %s" % bad
    assert "%016x" % fnv1a64(out["eqsrcs.s"]) == case["eqsrcs.s"]


def test_survey_golden_values(golden):
    """the values SURVEY.md section 4 quotes were produced independently from the same reference"""
    sv = golden["survey"]["grav3d_100000"]
    s = PortSession("grav3d", 100000, 100000)
    s.init_driver(); s.make_tree(0)
    assert "%016x" % fnv1a64(s.parts(0)["x"][0]) == sv["srcs.x0"]
    t = s.tree(0)
    assert "%016x" % fnv1a64(t["nr"]) == sv["stree.nr"]
    assert "%016x" % fnv1a64(t["num"]) == sv["stree.num"]
    s.upward(0); s.make_tree(1); s.refine(1)
    assert "%016x" % fnv1a64(s.parts(1)["gidx"]) == sv["targs.gidx"]
    bs = s.build_stats()
    assert bs["selects"] == 781            # 781 selects at N=1e5


def test_dtt_counts_match_survey(golden):
    """dual-tree interaction counts at N=1e5, theta=1.4 (from the reference with dostats=true)"""
    want = golden["survey"]["dtt_counts_1e5_t1.4"]
    s = PortSession("grav3d", 100000, 100000)
    s.init_driver(); s.make_tree(0); s.upward(0); s.make_tree(1); s.refine(1); s.upward(1)
    s.zero_vels(); s.fastsumm(1.4)
    st = s.stats()
    for k, v in want.items():
        assert st[k] == v, (k, st[k], v)
    # and the printed error metric of the reference for this run: rms 8.50221e-05
    u = s.parts(1)["u"][0].copy()
    s.zero_vels(); s.naive(5); un = s.parts(1)["u"][0]
    a = u[::5].astype(np.float32); b = un[::5].astype(np.float32)
    rms = np.sqrt(((a - b) ** 2).sum(dtype=np.float32) / (b ** 2).sum(dtype=np.float32))
    assert abs(rms - 8.50221e-05) < 2e-8


@pytest.mark.skipif(not ref_available("grav3d"), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("physics", PHYSICS)
def test_port_equals_compiled_reference(physics):
    n, theta = 6000, 1.3
    a = run_phases(RefSession(physics, n, n), theta)
    b = run_phases(PortSession(physics, n, n), theta)
    for k, v in a.items():
        if isinstance(v, np.ndarray):
            assert bits_equal(v, b[k]), k
        else:
            assert v == b[k], k


@pytest.mark.skipif(not ref_available("grav3d"), reason="compiled reference (oracle/_ref) not present")
def test_port_equals_reference_tree_1e6():
    """N=1e6: stall exits of the partial select (duplicate float keys) and libstdc++ tie order in refineLeaf"""
    n = 1000000
    r = RefSession("grav3d", n, n); p = PortSession("grav3d", n, n)
    for s in (r, p):
        s.init_driver(); s.make_tree(0); s.make_tree(1); s.refine(1)
    assert bits_equal(r.parts(0)["x"], p.parts(0)["x"])
    assert bits_equal(r.parts(1)["gidx"], p.parts(1)["gidx"])
    tr, tp = r.tree(0), p.tree(0)
    for k in ("num", "ioffset", "nc", "ns", "nr", "x", "s", "pr"):
        assert bits_equal(tr[k], tp[k]), k
    assert p.build_stats()["stalls"] == 2 and p.refine_tie_sorts() > 0
    # the hashes SURVEY.md section 4 obtained independently from the reference at this size (FNV-1a-64 over the raw bytes)
    h = lambda a: "%016x" % fnv1a64(np.ascontiguousarray(a))
    assert h(p.parts(1)["gidx"]) == "502bae5161063bf7" and h(p.parts(1)["x"][0]) == "f6c62a65e19a00c1"
    assert tp["levels"] == 14 and tp["numnodes"] == 16384
    assert h(tp["nr"]) == "16e0e91d02b96887" and h(tp["nc"][0]) == "f5df6737c4298b86" and h(tp["x"][0]) == "0b659ba5898c0aa6" and h(tp["num"]) == "ccbcffa83332a668"
    assert h(p.parts(0)["x"][0]) == "5b0b68d776e49dc5"          # intra-leaf source order of the strict (-O2) build


def test_reference_cli_binary_matches_golden_stdout():
    """the stock driver binary (when built): its 'particle 0 vel' and GFlop lines are the known answers of SURVEY section 4"""
    import os, subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "bin", "ongrav3d")
    if not os.path.exists(exe):
        pytest.skip("reference driver binary not built")
    out = subprocess.run([exe, "-n=20000", "-t=1.2", "-o=4", "-b=128"], stdout=subprocess.PIPE, text=True, timeout=300).stdout
    assert "error in fastsumm (max/rms)" in out and "[fast total]" in out and "[onbody naive]" in out


def _run_legacy(s, theta):
    """the drivers' sequence when -o is omitted (order = -1): refineTree(srcs) + calcEquivalents, then treecode2/3
    (ongrav3d.cpp:617-659); the dual tree has no target equivalents in that mode (barneshut.hpp:953)"""
    out = {}
    s.init_driver(); s.make_tree(0); s.refine(0); s.upward(0); s.make_tree(1)
    e, t = s.parts(2), s.tree(0)
    out.update({"srcs.x": s.parts(0)["x"], "eqsrcs.x": e["x"], "eqsrcs.r": e["r"], "eqsrcs.s": e["s"], "stree.epnum": t["epnum"], "stree.epoffset": t["epoffset"]})
    for name in ("treecode2", "treecode3"):
        s.zero_vels(); out[name + ".flops"] = getattr(s, name)(theta); out[name + ".u"] = s.parts(1)["u"]
    return out


@pytest.mark.skipif(not ref_available("grav3d"), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("physics,n,block", [("grav3d", 6100, 128), ("vort3d", 5000, 128), ("vort2d", 7001, 128), ("vortgrad3d", 3000, 128),
                                             ("grav3d", 5003, 64), ("vort2dtr", 4001, 32), ("grav3d", 300, 128), ("grav3d", 129, 128)])
def test_port_legacy_equivalents_equal_compiled_reference(physics, n, block):
    """-o omitted: pair-merge equivalents (barneshut.hpp:946-1061) of the restatement against the reference itself,
    including block sizes other than 128 (odd counts on the right spine) and trees of one or two levels"""
    a = _run_legacy(RefSession(physics, n, n, block=block, order=-1, eq_block=block), 1.2)
    b = _run_legacy(PortSession(physics, n, n, block=block, order=-1, eq_block=block), 1.2)
    for k, v in a.items():
        if isinstance(v, np.ndarray):
            assert bits_equal(v, b[k]), k
        else:
            assert v == b[k], k
    assert 0 < int(a["stree.epnum"][1]) <= block


def test_port_legacy_golden(golden):
    """known answers of the legacy path generated from the compiled reference (tests/golden/make_golden.py)"""
    case = golden.get("legacy")
    if not case:
        pytest.skip("golden.json has no legacy case")
    out = _run_legacy(PortSession(case["physics"], case["n"], case["n"], order=-1), case["theta"])
    for k, v in case.items():
        if k in out and isinstance(out[k], np.ndarray):
            assert "%016x" % fnv1a64(out[k]) == v, k
        elif k in out:
            assert float(out[k]) == v, k
