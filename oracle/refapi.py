"""ctypes front-end for the compiled reference (oracle/_ref/*/libref_<physics>.so).

TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may
import this module. The libraries are the UNMODIFIED reference templates compiled in place by
oracle/Makefile (see oracle/ref/hooks_common.hpp); they travel to the GPU box prebuilt.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

PHYSICS = ("grav3d", "vort3d", "vortgrad3d", "vort2d", "vort2dtr")
# strength_mode for oref_init_driver: the drivers that call wave_strengths() (onvort3d.cpp:593, onvortgrad3d.cpp:355)
_WAVE = {"grav3d": 0, "vort3d": 1, "vortgrad3d": 1, "vort2d": 0, "vort2dtr": 0}


def ref_lib_path(physics, build="strict"):
    return os.path.join(_HERE, "_ref", build, "libref_%s.so" % physics)


def ref_available(physics="grav3d", build="strict"):
    return os.path.exists(ref_lib_path(physics, build))


_f32p = C.POINTER(C.c_float)
_u64p = C.POINTER(C.c_uint64)


def _fp(a):
    return None if a is None else a.ctypes.data_as(_f32p)


def _up(a):
    return None if a is None else a.ctypes.data_as(_u64p)


class _NS:
    """maps lib.oref_xxx -> lib.<prefix>xxx so one front-end drives the reference hooks and the port"""

    def __init__(self, lib, prefix):
        object.__setattr__(self, "_lib", lib)
        object.__setattr__(self, "_prefix", prefix)

    def __getattr__(self, name):
        assert name.startswith("oref_")
        return getattr(self._lib, self._prefix + name[5:])


class RefSession:
    """One source set + one target set run through the reference's own call sequence."""

    _prefix = "oref_"

    def _open(self, physics, build):
        return C.CDLL(ref_lib_path(physics, build))

    def _create(self, L, physics, nsrc, ntarg, block, eq_block, order):
        L.oref_create.restype = C.c_void_p
        L.oref_create.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int]
        return L.oref_create(nsrc, ntarg, block, eq_block, order)

    def __init__(self, physics, nsrc, ntarg, block=128, order=4, eq_block=128, build="strict", accum64=False):
        # accum64: the reference templates instantiated with ACCUM = double (libref_<physics>_a64.so, grav3d and vort3d)
        self.accum64 = accum64
        self.rawlib = self._open(physics + ("_a64" if accum64 else ""), build)
        self.lib = _NS(self.rawlib, self._prefix)
        L = self.lib
        L.oref_destroy.argtypes = [C.c_void_p]
        L.oref_init_driver.argtypes = [C.c_void_p, C.c_int]
        L.oref_set_sources.argtypes = [C.c_void_p, _f32p, _f32p, _f32p]
        L.oref_set_targets.argtypes = [C.c_void_p, _f32p, _f32p]
        for fn in ("oref_make_tree", "oref_refine", "oref_upward"):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_int]
        L.oref_zero_vels.argtypes = [C.c_void_p]
        L.oref_naive.restype = C.c_float
        L.oref_naive.argtypes = [C.c_void_p, C.c_uint64]
        for fn in ("oref_treecode1", "oref_treecode2", "oref_treecode3"):
            getattr(L, fn).restype = C.c_float
            getattr(L, fn).argtypes = [C.c_void_p, C.c_float]
        L.oref_fastsumm.restype = C.c_int
        L.oref_fastsumm.argtypes = [C.c_void_p, C.c_float, C.c_int]
        L.oref_count.restype = C.c_uint64
        L.oref_count.argtypes = [C.c_void_p, C.c_int]
        L.oref_get_parts.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _f32p, _f32p, _u64p]
        L.oref_tree_shape.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oref_get_tree.argtypes = [C.c_void_p, C.c_int] + [_f32p] * 6 + [_u64p] * 4
        self.physics = physics
        self.nsrc, self.ntarg = int(nsrc), int(ntarg)
        self.h = self._create(L, physics, nsrc, ntarg, block, eq_block, order)
        pd, sd, od, hf = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._dims(L, pd, sd, od, hf)
        self.PD, self.SD, self.OD, self.has_fastsumm = pd.value, sd.value, od.value, bool(hf.value)

    def _dims(self, L, pd, sd, od, hf):
        L.oref_dims(C.byref(pd), C.byref(sd), C.byref(od), C.byref(hf))

    def close(self):
        if self.h:
            self.lib.oref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs
    def init_driver(self):
        self.lib.oref_init_driver(self.h, _WAVE[self.physics])

    def set_sources(self, x, r, s):
        x = np.ascontiguousarray(x, np.float32); r = np.ascontiguousarray(r, np.float32)
        s = np.ascontiguousarray(s, np.float32)
        assert x.shape == (self.PD, self.nsrc) and s.shape == (self.SD, self.nsrc)
        self.lib.oref_set_sources(self.h, _fp(x), _fp(r), _fp(s))

    def set_targets(self, x, r):
        x = np.ascontiguousarray(x, np.float32); r = np.ascontiguousarray(r, np.float32)
        assert x.shape == (self.PD, self.ntarg)
        self.lib.oref_set_targets(self.h, _fp(x), _fp(r))

    # ---- phases (which: 0 sources, 1 targets)
    def make_tree(self, which): self.lib.oref_make_tree(self.h, which)
    def refine(self, which): self.lib.oref_refine(self.h, which)
    def upward(self, which): self.lib.oref_upward(self.h, which)
    def zero_vels(self): self.lib.oref_zero_vels(self.h)
    def naive(self, tskip=1): return self.lib.oref_naive(self.h, tskip)
    def treecode1(self, theta): return self.lib.oref_treecode1(self.h, theta)
    def treecode2(self, theta): return self.lib.oref_treecode2(self.h, theta)
    def treecode3(self, theta): return self.lib.oref_treecode3(self.h, theta)

    def fastsumm(self, theta, parallel=False):
        rc = self.lib.oref_fastsumm(self.h, theta, int(parallel))
        if rc != 0:
            raise RuntimeError("%s has no dual-tree method in the reference" % self.physics)

    # ---- outputs
    def parts(self, which):
        """which: 0 srcs, 1 targs, 2 eqsrcs, 3 eqtargs -> dict of arrays"""
        n = int(self.lib.oref_count(self.h, which))
        src = which in (0, 2)
        x = np.zeros((self.PD, n), np.float32); r = np.zeros(n, np.float32)
        s = np.zeros((self.SD, n), np.float32) if src else None
        u = np.zeros((self.OD, n), np.float32) if not src else None
        g = np.full(n, np.iinfo(np.uint64).max, np.uint64) if which == 1 else None
        self.lib.oref_get_parts(self.h, which, _fp(x), _fp(r), _fp(s), _fp(u), _up(g))
        return {"n": n, "x": x, "r": r, "s": s, "u": u, "gidx": g}

    def results_f64(self, which=1):
        """outputs in the accumulator's precision: [OD, n] float64 (which: 1 targets, 3 equivalent targets)"""
        n = int(self.lib.oref_count(self.h, which))
        u = np.zeros((self.OD, n), np.float64)
        self.rawlib.oref_get_u64.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        self.rawlib.oref_get_u64(self.h, which, u.ctypes.data_as(C.POINTER(C.c_double)))
        return u

    def tree(self, which):
        lev, nn = C.c_int(), C.c_int()
        self.lib.oref_tree_shape(self.h, which, C.byref(lev), C.byref(nn))
        n = nn.value
        out = {"levels": lev.value, "numnodes": n,
               "x": np.zeros((self.PD, n), np.float32), "nc": np.zeros((self.PD, n), np.float32),
               "ns": np.zeros((self.PD, n), np.float32), "nr": np.zeros(n, np.float32),
               "pr": np.zeros(n, np.float32), "s": np.zeros((self.SD, n), np.float32),
               "ioffset": np.zeros(n, np.uint64), "num": np.zeros(n, np.uint64),
               "epoffset": np.zeros(n, np.uint64), "epnum": np.zeros(n, np.uint64)}
        self.lib.oref_get_tree(self.h, which, _fp(out["x"]), _fp(out["nc"]), _fp(out["ns"]), _fp(out["nr"]),
                               _fp(out["pr"]), _fp(out["s"]), _up(out["ioffset"]), _up(out["num"]),
                               _up(out["epoffset"]), _up(out["epnum"]))
        return out


def port_lib_path():
    return os.path.join(_HERE, "libonbody_oracle.so")


class PortSession(RefSession):
    """Same front-end over the CPU restatement oracle/port/oracle_port.cpp (libonbody_oracle.so)."""

    _prefix = "oport_"

    def _open(self, physics, build):
        return C.CDLL(port_lib_path())

    def _create(self, L, physics, nsrc, ntarg, block, eq_block, order):
        L.oref_create.restype = C.c_void_p
        L.oref_create.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int]
        return L.oref_create(PHYSICS.index(physics), nsrc, ntarg, block, eq_block, order)

    def _dims(self, L, pd, sd, od, hf):
        L.oref_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.oref_dims(self.h, C.byref(pd), C.byref(sd), C.byref(od), C.byref(hf))

    def stats(self):
        """sltp sbtp | sltl sbtl sltb sbtb tlc lpc bpc of the last treecode / dual-tree call"""
        out = (C.c_uint64 * 9)()
        self.rawlib.oport_get_stats.argtypes = [C.c_void_p, C.c_uint64 * 9]
        self.rawlib.oport_get_stats(self.h, out)
        k = ("sltp", "sbtp", "sltl", "sbtl", "sltb", "sbtb", "tlc", "lpc", "bpc")
        return dict(zip(k, [int(v) for v in out]))

    def build_stats(self):
        out = (C.c_uint64 * 4)()
        self.rawlib.oport_get_build_stats.argtypes = [C.c_void_p, C.c_uint64 * 4]
        self.rawlib.oport_get_build_stats(self.h, out)
        return dict(zip(("selects", "passes", "stalls", "scanned"), [int(v) for v in out]))

    def refine_tie_sorts(self):
        self.rawlib.oport_refine_tie_sorts.restype = C.c_uint64
        self.rawlib.oport_refine_tie_sorts.argtypes = [C.c_void_p]
        return int(self.rawlib.oport_refine_tie_sorts(self.h))


def fnv1a64(a):
    """FNV-1a-64 over the raw little-endian bytes (the hash SURVEY.md section 4 quotes)."""
    data = np.ascontiguousarray(a).view(np.uint8)
    if os.path.exists(port_lib_path()):
        lib = C.CDLL(port_lib_path())
        lib.oport_fnv1a64.restype = C.c_uint64
        lib.oport_fnv1a64.argtypes = [C.c_void_p, C.c_uint64]
        return int(lib.oport_fnv1a64(data.ctypes.data, data.size))
    h = np.uint64(1469598103934665603)
    prime = np.uint64(1099511628211)
    # vectorising FNV is not possible (serial dependency): chunked pure-python loop is fine up to ~1e6 bytes,
    # for bigger arrays use the C helper in the oracle port (oracle_fnv1a64)
    with np.errstate(over="ignore"):
        for b in data.tobytes():
            h = (h ^ np.uint64(b)) * prime
    return int(h)
