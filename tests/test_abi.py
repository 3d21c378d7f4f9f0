"""CPU tests of the drop-in boundary: the C-ABI libraries load, export every symbol their headers declare, and fail
loudly (no fallback) when there is no GPU."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    from onbody_b200.api import lib_path
    if not os.path.exists(lib_path()):
        from onbody_b200 import build
        build.build()
    return lib_path()


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:onb|external_vel)_[a-z0-9_]+)\s*\(", src)))


def test_core_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_built())
    names = _declared("onbody_b200.h")
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "include/onbody_b200.h declares %s but the library does not export it" % n


@pytest.mark.parametrize("header,libname", [("onbody_bh2dvort.h", "libbh2dvort_b200.so"), ("onbody_bh3dvortgrads.h", "libbh3dvortgrads_b200.so")])
def test_shim_libraries_export_the_reference_entry_points(header, libname):
    _built()
    path = os.path.join(ROOT, "onbody_b200", libname)
    if not os.path.exists(os.path.join(ROOT, "include", header)):
        pytest.skip("shim not written yet")
    lib = ctypes.CDLL(path)
    for n in _declared(header):
        assert hasattr(lib, n), n


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from onbody_b200.api import GpuSession, OnbodyError
    with pytest.raises(OnbodyError):
        GpuSession("grav3d", 1000, 1000)


def test_product_never_touches_the_oracle():
    """oracle/ is checker-only: nothing under onbody_b200/ or include/ may mention it"""
    for base in ("onbody_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            if "/build" in dp or "__pycache__" in dp:
                continue
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    for needle in ("import oracle", "from oracle", "oracle/", "oracle_port", "libonbody_oracle", "oport_", "oref_", "libref_", "refapi"):
                        assert needle not in txt, "%s references the oracle (%s)" % (os.path.join(dp, f), needle)


def test_driver_inputs_match_reference_generator(golden):
    """onb_driver_inputs (host code of the product) reproduces Parts::random_in_cube(std::mt19937) exactly"""
    _built()
    import numpy as np
    from onbody_b200.api import driver_inputs
    from oracle.refapi import PortSession
    for phys in ("grav3d", "vort3d", "vort2d"):
        x, r, s = driver_inputs(phys, 5000, True)
        o = PortSession(phys, 5000, 5000); o.init_driver(); p = o.parts(0)
        assert np.array_equal(x, p["x"]) and np.array_equal(r, p["r"]) and np.array_equal(s, p["s"])


def test_bench_reference_arm_runs():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fast", "libref_grav3d.so")):
        pytest.skip("compiled reference not present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--n-ref", "50000"],
                         stdout=subprocess.PIPE, text=True, timeout=600).stdout.strip().splitlines()[-1]
    import json
    rec = json.loads(out)
    assert rec["impl"] == "reference" and rec["value"] > 0 and rec["cpu_baseline"]["kind"] == "reference"


@pytest.mark.parametrize("headers", [("onbody_b200.h", "onbody_bh2dvort.h"), ("onbody_b200.h", "onbody_bh3dvortgrads.h")])
def test_headers_are_plain_c(headers, tmp_path):
    """the boundary is a C ABI: the headers must compile as C99 (no C++ types, no default arguments) - what a cgo / Fortran /
    ctypes binding generator would consume. (The two shim headers declare the same names with different arities, exactly as
    the reference's two libraries do, so they are checked separately.)"""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text("".join('#include "%s"\n' % h for h in headers) + "int main(void) { return 0; }\n")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-fsyntax-only", str(src)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


def test_communicator_entry_points_without_a_gpu():
    """host-only parts of the multi-GPU ABI work (and fail loudly) without a device: the partition arithmetic answers, the
    communicator refuses null contexts instead of crashing"""
    import ctypes as C
    lib = ctypes.CDLL(_built())
    lib.onb_shard_chunk_for.restype = C.c_uint64
    lib.onb_shard_chunk_for.argtypes = [C.c_uint64, C.c_int, C.c_int]
    assert lib.onb_shard_chunk_for(10 ** 9, 128, 8) == 976563 * 128          # ceil(7812500 leaves / 8) leaves per rank
    lib.onb_comm_init_all.argtypes = [C.c_void_p, C.c_int]
    assert lib.onb_comm_init_all(None, 2) != 0
    lib.onb_comm_init_loopback.argtypes = [C.c_void_p, C.c_int]
    assert lib.onb_comm_init_loopback(None, 2) != 0
