"""GPU tests of the drop-in boundary: the Fortran-style entry points of the two shim libraries against the reference's own
entry points (compiled in place into oracle/_ref), and the C++ drivers' stdout against the reference driver's."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import rel_rms, ROOT

pytestmark = pytest.mark.gpu
FP = C.POINTER(C.c_float)


def _p(a):
    return a.ctypes.data_as(FP)


def _ref(name):
    p = os.path.join(ROOT, "oracle", "_ref", "strict", "libref_%s.so" % name)
    if not os.path.exists(p):
        pytest.skip("compiled reference %s not present" % name)
    return C.CDLL(p)


def _inputs3(n, seed=3):
    rng = np.random.RandomState(seed)
    f = lambda *s: np.ascontiguousarray(rng.uniform(-1, 1, s).astype(np.float32))
    return dict(x=f(3, n), s=(f(3, n) / n).astype(np.float32), r=np.full(n, n ** (-1.0 / 3), np.float32), t=f(3, n))


@pytest.mark.parametrize("fn", ["external_vel_solver_f_", "external_vel_direct_f_"])
def test_bh3dvortgrads_entry_points(fn):
    ours = C.CDLL(os.path.join(ROOT, "onbody_b200", "libbh3dvortgrads_b200.so"))
    ref = _ref("vortgrad3d")
    ns, nt = (30000, 20000) if "solver" in fn else (3000, 2000)
    d = _inputs3(max(ns, nt))
    outs = []
    for lib in (ref, ours):
        f = getattr(lib, fn); f.restype = C.c_float
        o = np.full((12, nt), 0.25, np.float32)            # pre-filled: the entry points must ADD
        a = [C.byref(C.c_int(ns))] + [_p(d["x"][k, :ns].copy()) for k in range(3)] + [_p(d["s"][k, :ns].copy()) for k in range(3)] + [_p(d["r"][:ns].copy())]
        a += [C.byref(C.c_int(nt))] + [_p(d["t"][k, :nt].copy()) for k in range(3)]
        rows = [np.ascontiguousarray(o[k]) for k in range(12)]
        flops = f(*(a + [_p(r) for r in rows]))
        outs.append((np.stack(rows), flops))
    (uo, fo), (ug, fg) = outs
    assert fg == fo                                         # the flop estimate is a checksum of the interaction lists
    assert rel_rms(ug - 0.25, uo - 0.25) < 5e-6
    assert np.abs(ug - 0.25).max() > 0


@pytest.mark.parametrize("fn,reflib,tr", [("external_vel_solver_f_", "vort2d", False), ("external_vel_direct_f_", "vort2d", False),
                                          ("external_vel_solver_tr_f_", "iface2dvorttr", True), ("external_vel_direct_tr_f_", "iface2dvorttr", True)])
def test_bh2dvort_entry_points(fn, reflib, tr):
    ours = C.CDLL(os.path.join(ROOT, "onbody_b200", "libbh2dvort_b200.so"))
    ref = _ref(reflib)
    ns, nt = (40000, 30000) if "solver" in fn else (3000, 2000)
    rng = np.random.RandomState(5)
    f32 = lambda a: np.ascontiguousarray(a.astype(np.float32))
    sx, sy, tx, ty = (f32(rng.uniform(-1, 1, max(ns, nt))) for _ in range(4))
    ss = f32(rng.uniform(-1, 1, ns) / ns); sr = f32(np.full(ns, ns ** -0.5)); trr = f32(np.full(nt, 0.5 * nt ** -0.5))
    outs = []
    for lib in (ref, ours):
        f = getattr(lib, fn); f.restype = C.c_float
        tu = np.full(nt, 1.5, np.float32); tv = np.full(nt, -0.5, np.float32)
        a = [C.byref(C.c_int(ns)), _p(sx[:ns].copy()), _p(sy[:ns].copy()), _p(ss), _p(sr), C.byref(C.c_int(nt)), _p(tx[:nt].copy()), _p(ty[:nt].copy())]
        if tr:
            a.append(_p(trr))
        flops = f(*(a + [_p(tu), _p(tv)]))
        outs.append((tu - 1.5, tv + 0.5, flops))
    (uo, vo, fo), (ug, vg, fg) = outs
    assert fg == fo
    assert rel_rms(ug, uo) < 5e-6 and rel_rms(vg, vo) < 5e-6


def _parse(out):
    d = {}
    cur = None
    for line in out.splitlines():
        m = re.match(r"\[onbody (\w+)\]", line)
        if m:
            cur = m.group(1)
        m = re.match(r"\s+particle 0 vel (.*)", line)
        if m and cur:
            d[cur + ".vel"] = [float(v) for v in m.group(1).split()]
        m = re.match(r"\s+GFlop: ([\d.]+)", line)
        if m and cur:
            d[cur + ".gflop"] = m.group(1)
        m = re.match(r"error in (\w+) \(max/rms\):\s+(\S+) / (\S+)", line)
        if m:
            d[m.group(1) + ".err"] = (float(m.group(2)), float(m.group(3)))
    return d


@pytest.mark.parametrize("exe,args", [("ongrav3d", ["-n=20000", "-t=1.2", "-o=4", "-b=128"]), ("onvort3d", ["-n=20000", "-t=1.2", "-o=4"]),
                                      ("onvort2d", ["-n=20000", "-t=1.3", "-o=4"]), ("onvortgrad3d", ["-n=10000", "-t=1.2", "-o=4"])])
def test_driver_stdout_matches_reference_driver(exe, args):
    ours = os.path.join(ROOT, "onbody_b200", "bin", exe)
    ref = os.path.join(ROOT, "oracle", "_ref", "bin", exe)
    if not os.path.exists(ref):
        pytest.skip("reference driver binary not present")
    env = dict(os.environ, OMP_NUM_THREADS="1")          # race-free dual tree in the reference (README.md:200)
    a = subprocess.run([ref] + args, stdout=subprocess.PIPE, text=True, timeout=900, env=env).stdout
    b = subprocess.run([ours] + args, stdout=subprocess.PIPE, text=True, timeout=300).stdout
    for key in ("error in", "[onbody naive]", "Done."):
        assert key in b
    A, Bd = _parse(a), _parse(b)
    assert set(A) == set(Bd), (sorted(A), sorted(Bd))
    for k, v in A.items():
        if k.endswith(".gflop") and not k.startswith("naive"):
            assert v == Bd[k], (k, v, Bd[k])              # printed digits identical: same interaction lists
        elif k.endswith(".vel"):
            assert np.allclose(v, Bd[k], rtol=2e-4, atol=0), (k, v, Bd[k])
        elif k.endswith(".err"):
            assert abs(v[1] - Bd[k][1]) <= 0.1 * v[1] + 1e-7, (k, v, Bd[k])   # same accuracy as the reference (rms within 10 %)


def test_driver_flag_quirks():
    exe = os.path.join(ROOT, "onbody_b200", "bin", "ongrav3d")
    assert subprocess.run([exe, "-t1=1.2"], stderr=subprocess.PIPE, stdout=subprocess.PIPE).returncode == 1      # ongrav3d.cpp:491 bug kept
    assert subprocess.run([exe, "-h"], stderr=subprocess.PIPE, stdout=subprocess.PIPE).returncode == 1


def test_driver_default_order_runs_legacy_equivalents():
    """-o omitted (the reference's default, order = -1): treecode / treecode2 / treecode3 over the pair-merge equivalents
    print the reference's lines; the dual tree is skipped with a note (the reference builds no target equivalents there)"""
    ours = os.path.join(ROOT, "onbody_b200", "bin", "ongrav3d")
    ref = os.path.join(ROOT, "oracle", "_ref", "bin", "ongrav3d")
    if not os.path.exists(ref):
        pytest.skip("reference driver binary not present")
    args = ["-n=20000", "-t=1.2"]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    a = subprocess.run([ref] + args, stdout=subprocess.PIPE, text=True, timeout=900, env=env).stdout
    r = subprocess.run([ours] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0 and "dual-tree method is skipped" in r.stderr
    assert "with equivalent particles and theta" in r.stdout and "refine within leaf nodes" in r.stdout
    A, Bd = _parse(a), _parse(r.stdout)
    for k, v in A.items():
        if k.startswith("fast"):
            continue
        assert k in Bd, k
        if k.endswith(".gflop") and not k.startswith("naive"):
            assert v == Bd[k], (k, v, Bd[k])
        elif k.endswith(".vel"):
            assert np.allclose(v, Bd[k], rtol=2e-4, atol=0), (k, v, Bd[k])
        elif k.endswith(".err"):
            assert abs(v[1] - Bd[k][1]) <= 0.1 * v[1] + 1e-7, (k, v, Bd[k])


@pytest.mark.parametrize("exe,n", [("run3dvortgrads", 20000), ("run2dvort", 20000)])
def test_reference_abi_test_programs_run_against_the_shims(exe, n):
    """the reference's own ABI test programs (main3dvortgrads.cpp:136-206, main2dvort.cpp:103-155) are "the ABI's
    only tests". The same unmodified main(), linked once against the reference's interface*.cpp (oracle/_ref/bin/<exe>) and
    once against the B200 shim library (<exe>_b200, oracle/Makefile), must report the same accuracy of the fast solver
    against the direct solver - both of which it calls through the Fortran-style entry points."""
    ref = os.path.join(ROOT, "oracle", "_ref", "bin", exe)
    ours = ref + "_b200"
    if not (os.path.exists(ref) and os.path.exists(ours)):
        pytest.skip("reference ABI test programs not present (built where /root/reference exists)")
    arg = ["-n=%d" % n]
    a = subprocess.run([ref] + arg, stdout=subprocess.PIPE, text=True, timeout=900).stdout
    b = subprocess.run([ours] + arg, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert b.returncode == 0, b.stderr

    def errs(txt):
        return {k: float(re.search(k + r" error in fast solver:\s*(\S+)", txt).group(1)) for k in ("rms", "max")}
    ea, eb = errs(a), errs(b.stdout)
    assert a.splitlines()[0] == b.stdout.splitlines()[0]                      # "Running main... with n sources and n targets"
    for key in ("external_vel_solver_f_:", "external_vel_direct_f_:"):
        assert key in b.stdout
    assert abs(eb["rms"] - ea["rms"]) <= 0.25 * ea["rms"], (ea, eb)           # same solver accuracy (the numbers sit near float32 noise)
    assert eb["max"] <= 2.0 * ea["max"], (ea, eb)


def test_driver_multi_gpu_mode_prints_what_one_gpu_prints():
    """ongrav3d -g=3: three contexts, one host thread each, the library's communicator behind the same phase calls (here the
    loopback transport on one GPU; on a multi-GPU box the same flag uses NCCL). Every printed result line must equal the
    single-GPU run's digit for digit - the distributed evaluation is bit-identical."""
    exe = os.path.join(ROOT, "onbody_b200", "bin", "ongrav3d")
    args = ["-n=60000", "-t=1.3", "-o=4", "-b=128"]
    a = subprocess.run([exe] + args, stdout=subprocess.PIPE, text=True, timeout=300).stdout
    env = dict(os.environ, ONBODY_B200_LOOPBACK="1")
    r = subprocess.run([exe, "-g=3"] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    assert "loopback communicator" in r.stderr
    A, B = _parse(a), _parse(r.stdout)
    assert set(A) == set(B)
    for k, v in A.items():
        if k.endswith(".vel") or k.endswith(".err"):
            assert v == B[k], (k, v, B[k])
        elif k.endswith(".gflop") and not k.startswith("naive"):
            assert abs(float(v) - float(B[k])) <= 2e-3, (k, v, B[k])      # the ranks' float flop counts are added up: last printed digit
