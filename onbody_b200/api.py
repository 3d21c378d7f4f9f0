"""ctypes mirror of include/onbody_b200.h.

``GpuSession`` exposes the same phase methods, in the same order and with the same meaning, as the reference
drivers call them (make_tree / refine / upward / naive / treecode1,2,3 / fastsumm); tests drive it side by side
with the oracle's session objects. Nothing here computes: every method is one call through the C ABI.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PHYSICS = ("grav3d", "vort3d", "vortgrad3d", "vort2d", "vort2dtr")
_WAVE = {"grav3d": 0, "vort3d": 1, "vortgrad3d": 1, "vort2d": 0, "vort2dtr": 0}
_DIMS = {"grav3d": (3, 1, 3), "vort3d": (3, 3, 3), "vortgrad3d": (3, 3, 12), "vort2d": (2, 1, 2), "vort2dtr": (2, 1, 2)}

ARITH_FAST, ARITH_STRICT = 0, 1

_f32p = C.POINTER(C.c_float)
_u64p = C.POINTER(C.c_uint64)


class OnbodyError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "libonbody_b200.so")


_lib = None


def load_library():
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise OnbodyError("%s is missing: run `python -m onbody_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback." % p)
    L = C.CDLL(p)
    L.onb_create.restype = C.c_void_p
    L.onb_create.argtypes = [C.c_int, C.c_int]
    L.onb_destroy.argtypes = [C.c_void_p]
    L.onb_error.restype = C.c_char_p
    L.onb_error.argtypes = [C.c_void_p]
    L.onb_last_create_error.restype = C.c_char_p
    L.onb_set_params.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.onb_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
    L.onb_set_sources.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.onb_set_targets.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.onb_driver_inputs.argtypes = [C.c_int, C.c_uint64, C.c_int, _f32p, _f32p, _f32p]
    for fn in ("onb_make_tree", "onb_refine", "onb_upward"):
        getattr(L, fn).argtypes = [C.c_void_p, C.c_int]
    L.onb_make_tree_range.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]
    L.onb_finish_tree.argtypes = [C.c_void_p, C.c_int]
    L.onb_make_trees.argtypes = [C.c_void_p]
    L.onb_make_trees_range.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]
    L.onb_set_build_range.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]
    L.onb_prepare_eval.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]
    L.onb_shard_particle_range.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, _u64p, _u64p]
    L.onb_device_ptr.restype = C.c_void_p
    L.onb_device_ptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.onb_zero_vels.argtypes = [C.c_void_p]
    L.onb_naive.argtypes = [C.c_void_p, C.c_uint64, _f32p]
    for fn in ("onb_treecode1", "onb_treecode2", "onb_treecode3"):
        getattr(L, fn).argtypes = [C.c_void_p, C.c_float, _f32p]
    L.onb_fastsumm.argtypes = [C.c_void_p, C.c_float]
    L.onb_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.onb_set_async_inputs.argtypes = [C.c_void_p, C.c_int]
    L.onb_set_sliced_inputs.argtypes = [C.c_void_p, C.c_int]
    L.onb_count.restype = C.c_uint64
    L.onb_count.argtypes = [C.c_void_p, C.c_int]
    L.onb_get_parts.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.onb_add_results_original_order.argtypes = [C.c_void_p, C.c_void_p]
    L.onb_tree_shape.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.onb_get_tree.argtypes = [C.c_void_p, C.c_int] + [_f32p] * 6 + [_u64p] * 4
    L.onb_get_stats.argtypes = [C.c_void_p, C.c_uint64 * 9]
    L.onb_phase_ms.restype = C.c_double
    L.onb_phase_ms.argtypes = [C.c_void_p, C.c_char_p]
    L.onb_timer_start.argtypes = [C.c_void_p]
    L.onb_timer_stop_ms.restype = C.c_double
    L.onb_timer_stop_ms.argtypes = [C.c_void_p]
    L.onb_last_pairs.restype = C.c_uint64
    L.onb_last_pairs.argtypes = [C.c_void_p]
    L.onb_launch_count.restype = C.c_uint64
    L.onb_launch_count.argtypes = [C.c_void_p]
    L.onb_measure_fp32_peak.restype = C.c_double
    L.onb_measure_fp32_peak.argtypes = [C.c_void_p]
    L.onb_load_tree.argtypes = [C.c_void_p, C.c_int, C.c_int] + [_f32p] * 6 + [_u64p] * 2
    L.onb_get_build_stats.argtypes = [C.c_void_p, C.c_uint64 * 5]
    vpp = C.POINTER(C.c_void_p)
    L.onb_set_sources_planes.argtypes = [C.c_void_p, C.c_uint64, vpp, C.c_void_p, vpp]
    L.onb_set_targets_planes.argtypes = [C.c_void_p, C.c_uint64, vpp, C.c_void_p]
    L.onb_add_results_planes.argtypes = [C.c_void_p, vpp]
    L.onb_get_shard_results.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, _u64p, _u64p]
    L.onb_comm_unique_id.argtypes = [C.c_void_p, C.c_uint64]
    L.onb_comm_init_rank.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
    L.onb_comm_init_all.argtypes = [vpp, C.c_int]
    L.onb_comm_init_loopback.argtypes = [vpp, C.c_int]
    L.onb_comm_destroy.argtypes = [C.c_void_p]
    L.onb_comm_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
    L.onb_set_memory_mode.argtypes = [C.c_void_p, C.c_int]
    L.onb_set_accum.argtypes = [C.c_void_p, C.c_int]
    L.onb_get_results_f64.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.onb_device_memory.argtypes = [C.c_void_p, _u64p, _u64p]
    L.onb_shard_range_for.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, _u64p, _u64p]
    L.onb_set_pivot_mode.argtypes = [C.c_int]
    if os.environ.get("ONB_P2P_TPT"):          # tuning knob: targets per thread of the pair kernel (1, 2, 4; +16 = scalar arithmetic)
        L.onb_set_p2p_tpt(int(os.environ["ONB_P2P_TPT"]))
    _lib = L
    return L


def _fp(a):
    return None if a is None else a.ctypes.data_as(_f32p)


def _up(a):
    return None if a is None else a.ctypes.data_as(_u64p)


def driver_inputs(physics, n, sources=True):
    """The reference drivers' synthetic inputs (mt19937(12345), Parts.hpp:99-109, wave_strengths :169-176),
    generated on the host by the product library. Returns x [PD,n], r [n], s [SD,n] (s None for targets)."""
    L = load_library()
    PD, SD, _ = _DIMS[physics]
    x = np.empty((PD, n), np.float32); r = np.empty(n, np.float32)
    s = np.empty((SD, n), np.float32) if sources else None
    rc = L.onb_driver_inputs(PHYSICS.index(physics), n, _WAVE[physics], _fp(x), _fp(r), _fp(s))
    if rc != 0:
        raise OnbodyError("onb_driver_inputs failed (%d)" % rc)
    return x, r, s


MEM_NORMAL, MEM_LEAN = 0, 1
UNIQUE_ID_BYTES = 128


def comm_unique_id():
    """128 bytes rank 0 ships to the other ranks (ncclGetUniqueId inside the library)"""
    L = load_library()
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    rc = L.onb_comm_unique_id(buf, UNIQUE_ID_BYTES)
    if rc != 0:
        raise OnbodyError("onb_comm_unique_id failed (%d): is libnccl.so.2 loadable?" % rc)
    return buf.raw


def _join(sessions, fn, what):
    L = load_library()
    arr = (C.c_void_p * len(sessions))(*[s.h for s in sessions])
    rc = getattr(L, fn)(arr, len(sessions))
    if rc != 0:
        raise OnbodyError("%s failed (%d): %s" % (what, rc, L.onb_error(sessions[0].h).decode()))
    for i, s in enumerate(sessions):
        s.rank, s.world = i, len(sessions)


def comm_init_all(sessions):
    """one process, one session per GPU: ncclCommInitAll; then drive each session from its own host thread"""
    _join(sessions, "onb_comm_init_all", "onb_comm_init_all")


def comm_init_loopback(sessions):
    """test transport: sessions of this process (any devices, also one) joined by device copies instead of NCCL"""
    _join(sessions, "onb_comm_init_loopback", "onb_comm_init_loopback")


def shard_range_for(n, block, rank, nranks):
    lo, hi = C.c_uint64(), C.c_uint64()
    load_library().onb_shard_range_for(n, block, rank, nranks, C.byref(lo), C.byref(hi))
    return int(lo.value), int(hi.value)


class GpuSession:
    """One source set + one target set on one B200, phase by phase."""

    def __init__(self, physics, nsrc=None, ntarg=None, block=128, order=4, arith=ARITH_FAST, device=0, accum64=False, **_ignored):
        self.lib = load_library()
        self.physics = physics
        self.PD, self.SD, self.OD = _DIMS[physics]
        self.has_fastsumm = physics != "vortgrad3d"
        self.nsrc, self.ntarg = nsrc, ntarg
        self.block, self.rank, self.world = block, 0, 1
        self.h = self.lib.onb_create(PHYSICS.index(physics), device)
        if not self.h:
            raise OnbodyError("onb_create failed: %s" % self.lib.onb_last_create_error().decode())
        self._chk(self.lib.onb_set_params(self.h, block, order, arith))
        if accum64:            # the reference's ACCUM = double (ongrav3d.cpp:8): fp32 pair arithmetic, fp64 accumulation and outputs
            self._chk(self.lib.onb_set_accum(self.h, 1))

    def results_f64(self, which=1):
        n = int(self.lib.onb_count(self.h, which))
        u = np.zeros((self.OD, n), np.float64)
        self._chk(self.lib.onb_get_results_f64(self.h, which, u.ctypes.data))
        return u

    def _chk(self, rc):
        if rc != 0:
            raise OnbodyError("onbody_b200 error %d: %s" % (rc, self.lib.onb_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.onb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs (host arrays; may be pinned torch tensors' numpy views or raw pointers)
    def init_driver(self):
        n = self.nsrc
        x, r, s = driver_inputs(self.physics, n, True)
        self.set_sources(x, r, s)
        if self.ntarg == n:
            self.set_targets(x, r)          # the drivers copy the engine: targets == sources (ongrav3d.cpp:574-594)
        else:
            tx, tr, _ = driver_inputs(self.physics, self.ntarg, False)
            self.set_targets(tx, tr)

    def set_sources(self, x, r, s):
        x = np.ascontiguousarray(x, np.float32); r = np.ascontiguousarray(r, np.float32); s = np.ascontiguousarray(s, np.float32)
        n = r.shape[0]
        assert x.shape == (self.PD, n) and s.shape == (self.SD, n)
        self.nsrc = n
        self._chk(self.lib.onb_set_sources(self.h, n, x.ctypes.data, r.ctypes.data, s.ctypes.data))

    def set_targets(self, x, r):
        x = np.ascontiguousarray(x, np.float32); r = np.ascontiguousarray(r, np.float32)
        n = r.shape[0]
        assert x.shape == (self.PD, n)
        self.ntarg = n
        self._chk(self.lib.onb_set_targets(self.h, n, x.ctypes.data, r.ctypes.data))

    def set_sources_ptr(self, n, x_ptr, r_ptr, s_ptr):
        self.nsrc = n
        self._chk(self.lib.onb_set_sources(self.h, n, x_ptr, r_ptr, s_ptr))

    def set_targets_ptr(self, n, x_ptr, r_ptr):
        self.ntarg = n
        self._chk(self.lib.onb_set_targets(self.h, n, x_ptr, r_ptr))

    def set_async_inputs(self, on):
        self._chk(self.lib.onb_set_async_inputs(self.h, 1 if on else 0))

    def set_sliced_inputs(self, on):
        self._chk(self.lib.onb_set_sliced_inputs(self.h, 1 if on else 0))

    def set_shard(self, rank, nranks):
        self._chk(self.lib.onb_set_shard(self.h, rank, nranks))
        self.rank, self.world = rank, nranks

    # ---- multi-GPU: the communicator lives in the library (csrc/comm.cu); with one attached the same phase calls run distributed
    def comm_init_rank(self, rank, nranks, unique_id):
        self._chk(self.lib.onb_comm_init_rank(self.h, rank, nranks, unique_id, len(unique_id)))
        self.rank, self.world = rank, nranks

    def comm_destroy(self):
        self._chk(self.lib.onb_comm_destroy(self.h))
        self.rank, self.world = 0, 1

    def comm_info(self):
        v = [C.c_int() for _ in range(4)]
        self.lib.onb_comm_info(self.h, *[C.byref(x) for x in v])
        return {"rank": v[0].value, "nranks": v[1].value, "transport": ("none", "nccl", "loopback")[v[2].value], "nccl_version": v[3].value}

    def set_memory_mode(self, mode):
        self._chk(self.lib.onb_set_memory_mode(self.h, mode))

    def device_memory(self):
        u, t = C.c_uint64(), C.c_uint64()
        self._chk(self.lib.onb_device_memory(self.h, C.byref(u), C.byref(t)))
        return int(u.value), int(t.value)

    def shard_results_into(self, u_ptr, stride=0):
        """device -> host copy of this rank's shard of the target outputs: u[d*stride + i] for i in [lo, hi) (stride 0 = n, the
        full [OD][n] layout; a caller holding only its shard passes the shard length and a pointer shifted by -lo); returns (lo, hi)"""
        lo, hi = C.c_uint64(), C.c_uint64()
        self._chk(self.lib.onb_get_shard_results(self.h, u_ptr, stride, C.byref(lo), C.byref(hi)))
        return int(lo.value), int(hi.value)

    def set_planes_ptr(self, n, x_ptrs, r_ptr, s_ptrs):
        """sources and targets from raw per-plane pointers (host or device)"""
        self.nsrc = self.ntarg = n
        X = (C.c_void_p * len(x_ptrs))(*x_ptrs); S = (C.c_void_p * len(s_ptrs))(*s_ptrs)
        self._chk(self.lib.onb_set_sources_planes(self.h, n, X, r_ptr, S))
        self._chk(self.lib.onb_set_targets_planes(self.h, n, X, r_ptr))

    def set_sources_planes(self, xs, r, ss):
        """one array per plane (what the Fortran-style entry points receive)"""
        n = r.shape[0]; self.nsrc = n
        X = (C.c_void_p * len(xs))(*[a.ctypes.data for a in xs]); S = (C.c_void_p * len(ss))(*[a.ctypes.data for a in ss])
        self._chk(self.lib.onb_set_sources_planes(self.h, n, X, r.ctypes.data, S))

    def set_targets_planes(self, xs, r):
        n = r.shape[0]; self.ntarg = n
        X = (C.c_void_p * len(xs))(*[a.ctypes.data for a in xs])
        self._chk(self.lib.onb_set_targets_planes(self.h, n, X, r.ctypes.data))

    def add_results_planes(self, planes):
        P = (C.c_void_p * len(planes))(*[a.ctypes.data if hasattr(a, "ctypes") else int(a) for a in planes])
        self._chk(self.lib.onb_add_results_planes(self.h, P))

    # ---- phases
    def make_tree(self, which): self._chk(self.lib.onb_make_tree(self.h, which))
    def make_tree_range(self, which, lo, hi): self._chk(self.lib.onb_make_tree_range(self.h, which, lo, hi))
    def finish_tree(self, which): self._chk(self.lib.onb_finish_tree(self.h, which))
    def make_trees(self): self._chk(self.lib.onb_make_trees(self.h))
    def make_trees_range(self, slo, shi, tlo, thi): self._chk(self.lib.onb_make_trees_range(self.h, slo, shi, tlo, thi))
    def prepare_eval(self, finish=False, tgt_lo=0, tgt_hi=2**63): self._chk(self.lib.onb_prepare_eval(self.h, 1 if finish else 0, tgt_lo, tgt_hi))
    def set_build_range(self, which, lo, hi): self._chk(self.lib.onb_set_build_range(self.h, which, lo, hi))

    def shard_particle_range(self, n, rank, nranks):
        lo, hi = C.c_uint64(), C.c_uint64()
        self._chk(self.lib.onb_shard_particle_range(self.h, n, rank, nranks, C.byref(lo), C.byref(hi)))
        return int(lo.value), int(hi.value)

    def device_ptr(self, which, field):
        """raw device pointer of a particle plane (field 0..2 x[d], 3 r, 4..6 s[d]) for zero-copy collectives"""
        return self.lib.onb_device_ptr(self.h, which, field)

    def plane_tensor(self, which, field, n):
        """the plane as a torch CUDA tensor aliasing the library's memory (plumbing for torch.distributed collectives)"""
        import torch

        class _P:
            pass
        p = _P()
        p.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(self.device_ptr(which, field)), False), "version": 2}
        return torch.as_tensor(p, device="cuda")

    def source_fields(self):
        return list(range(self.PD)) + [3] + [4 + d for d in range(self.SD)]

    def refine(self, which): self._chk(self.lib.onb_refine(self.h, which))
    def upward(self, which): self._chk(self.lib.onb_upward(self.h, which))
    def zero_vels(self): self._chk(self.lib.onb_zero_vels(self.h))

    def naive(self, tskip=1):
        f = C.c_float(); self._chk(self.lib.onb_naive(self.h, tskip, C.byref(f))); return f.value

    def treecode1(self, theta):
        f = C.c_float(); self._chk(self.lib.onb_treecode1(self.h, theta, C.byref(f))); return f.value

    def treecode2(self, theta):
        f = C.c_float(); self._chk(self.lib.onb_treecode2(self.h, theta, C.byref(f))); return f.value

    def treecode3(self, theta):
        f = C.c_float(); self._chk(self.lib.onb_treecode3(self.h, theta, C.byref(f))); return f.value

    def fastsumm(self, theta, parallel=False):
        self._chk(self.lib.onb_fastsumm(self.h, theta))

    # ---- outputs
    def parts(self, which, want=("x", "r", "s", "u", "gidx")):
        n = int(self.lib.onb_count(self.h, which))
        src = which in (0, 2)
        x = np.zeros((self.PD, n), np.float32) if "x" in want else None
        r = np.zeros(n, np.float32) if "r" in want else None
        s = np.zeros((self.SD, n), np.float32) if (src and "s" in want) else None
        u = np.zeros((self.OD, n), np.float32) if (not src and "u" in want) else None
        g = np.full(n, np.iinfo(np.uint64).max, np.uint64) if (which == 1 and "gidx" in want) else None
        ptr = lambda a: None if a is None else a.ctypes.data
        self._chk(self.lib.onb_get_parts(self.h, which, ptr(x), ptr(r), ptr(s), ptr(u), ptr(g)))
        return {"n": n, "x": x, "r": r, "s": s, "u": u, "gidx": g}

    def results_into(self, u_ptr):
        """device -> host copy of the target outputs [OD][n] (tree order) into caller memory"""
        self._chk(self.lib.onb_get_parts(self.h, 1, None, None, None, u_ptr, None))

    def add_results_original_order(self, u):
        assert u.dtype == np.float32 and u.shape == (self.OD, self.ntarg) and u.flags.c_contiguous
        self._chk(self.lib.onb_add_results_original_order(self.h, u.ctypes.data))

    def tree(self, which):
        lev, nn = C.c_int(), C.c_int()
        self._chk(self.lib.onb_tree_shape(self.h, which, C.byref(lev), C.byref(nn)))
        n = nn.value
        out = {"levels": lev.value, "numnodes": n,
               "x": np.zeros((self.PD, n), np.float32), "nc": np.zeros((self.PD, n), np.float32),
               "ns": np.zeros((self.PD, n), np.float32), "nr": np.zeros(n, np.float32),
               "pr": np.zeros(n, np.float32), "s": np.zeros((self.SD, n), np.float32),
               "ioffset": np.zeros(n, np.uint64), "num": np.zeros(n, np.uint64),
               "epoffset": np.zeros(n, np.uint64), "epnum": np.zeros(n, np.uint64)}
        self._chk(self.lib.onb_get_tree(self.h, which, _fp(out["x"]), _fp(out["nc"]), _fp(out["ns"]), _fp(out["nr"]),
                                        _fp(out["pr"]), _fp(out["s"]), _up(out["ioffset"]), _up(out["num"]),
                                        _up(out["epoffset"]), _up(out["epnum"])))
        return out

    def load_tree(self, which, t):
        """test support: install an oracle-built tree (particles must already be set in tree order)"""
        c = lambda a, dt: np.ascontiguousarray(a, dt)
        self._chk(self.lib.onb_load_tree(self.h, which, int(t["levels"]), _fp(c(t["x"], np.float32)), _fp(c(t["nc"], np.float32)),
                                         _fp(c(t["ns"], np.float32)), _fp(c(t["nr"], np.float32)), _fp(c(t["pr"], np.float32)),
                                         _fp(c(t["s"], np.float32)), _up(c(t["ioffset"], np.uint64)), _up(c(t["num"], np.uint64))))

    def stats(self):
        out = (C.c_uint64 * 9)()
        self.lib.onb_get_stats(self.h, out)
        k = ("sltp", "sbtp", "sltl", "sbtl", "sltb", "sbtb", "tlc", "lpc", "bpc")
        return dict(zip(k, [int(v) for v in out]))

    def build_stats(self):
        out = (C.c_uint64 * 5)()
        self._chk(self.lib.onb_get_build_stats(self.h, out))
        return dict(zip(("selects", "passes", "stalls", "scanned", "tie_sorts"), [int(v) for v in out]))

    def timer_start(self): self._chk(self.lib.onb_timer_start(self.h))
    def timer_stop_ms(self): return float(self.lib.onb_timer_stop_ms(self.h))
    def phase_ms(self, name): return float(self.lib.onb_phase_ms(self.h, name.encode()))
    def last_pairs(self): return int(self.lib.onb_last_pairs(self.h))
    def launch_count(self): return int(self.lib.onb_launch_count(self.h))
    def measure_fp32_peak(self): return float(self.lib.onb_measure_fp32_peak(self.h))
