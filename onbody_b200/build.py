"""Builds the CUDA library in-tree with nvcc for sm_100a only (cross-compiles without a GPU).

    python -m onbody_b200.build            # build what is out of date
    python -m onbody_b200.build --force
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = "/usr/bin/g++"   # the image's CXX env points at a gcc without libgomp; use the system one

CORE_SOURCES = ["context.cu", "p2p.cu", "tree.cu", "bary.cu", "traverse.cu", "pointwise.cu", "scan.cu", "mem.cu", "plan.cu", "comm.cu", "dist.cu"]
CORE_LIB = os.path.join(HERE, "libonbody_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-ccbin", HOSTCXX, "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr", "-Xptxas", "-v", "-cudart", "static",
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, log=None):
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        with open(log, "a") as f:
            f.write("$ " + " ".join(cmd) + "\n" + res.stdout + "\n")
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return res.stdout


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    log = os.path.join(HERE, "build", "build.log")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "onbody_b200.h"))
    objs = []
    for src in CORE_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(HERE, "build", src + ".o")
        if force or _newer(o, [s] + headers):
            out = _run([NVCC] + NVCC_FLAGS + ["-c", s, "-o", o], log)
            if verbose:
                print(out)
        objs.append(o)
    if force or _newer(CORE_LIB, objs):
        _run([NVCC, "-shared", "-cudart", "static", "-ccbin", HOSTCXX, "-o", CORE_LIB] + objs +
             ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl", "-lpthread"], log)
    build_hosts(force)
    return CORE_LIB


def build_hosts(force=False):
    """the drop-in shim libraries and the C++ drivers (host code only, linked against the core library)"""
    host_dir = os.path.join(CSRC, "host")
    if not os.path.isdir(host_dir):
        return
    log = os.path.join(HERE, "build", "build.log")
    inc = ["-I" + os.path.join(ROOT, "include")]
    link = ["-L" + HERE, "-lonbody_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + HERE]
    for name, srcs in (("libbh2dvort_b200.so", ["iface2dvort.cpp"]), ("libbh3dvortgrads_b200.so", ["iface3dvortgrads.cpp"])):
        paths = [os.path.join(host_dir, s) for s in srcs]
        if not all(os.path.exists(p) for p in paths):
            continue
        out = os.path.join(HERE, name)
        if force or _newer(out, paths + [CORE_LIB]):
            _run([HOSTCXX, "-std=c++14", "-O2", "-fPIC", "-shared", "-o", out] + paths + inc + link, log)
    bindir = os.path.join(HERE, "bin")
    os.makedirs(bindir, exist_ok=True)
    for exe, src in (("ongrav3d", "ongrav3d.cpp"), ("onvort3d", "onvort3d.cpp"), ("onvort2d", "onvort2d.cpp"),
                     ("onvortgrad3d", "onvortgrad3d.cpp")):
        p = os.path.join(host_dir, src)
        if not os.path.exists(p):
            continue
        out = os.path.join(bindir, exe)
        deps = [p, CORE_LIB] + [os.path.join(host_dir, f) for f in os.listdir(host_dir) if f.endswith(".hpp")]
        if force or _newer(out, deps):
            _run([HOSTCXX, "-std=c++14", "-O2", "-o", out, p] + inc +
                 ["-L" + HERE, "-lonbody_b200", "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath," + HERE], log)


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", lib)
