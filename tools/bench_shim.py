"""The drop-in Fortran-style entry point external_vel_solver_f_ (3-D vortex velocity + gradients, interface3dvortgrads.cpp:247-416)
timed end to end from HOST arrays - the call an existing caller of the reference makes - through libbh3dvortgrads_b200.so, next to
the reference's own entry point (oracle/_ref/fast, OpenMP, all host threads) on the same inputs. Prints one JSON line.
    python tools/bench_shim.py [N] [reps]"""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
FP = C.POINTER(C.c_float)
rng = np.random.RandomState(7)
f = lambda *s: np.ascontiguousarray(rng.uniform(0, 1, s).astype(np.float32))
sx, ss, tx = f(3, N), (f(3, N) / N).astype(np.float32), f(3, N)
sr = np.full(N, 1.0 / np.sqrt(N), np.float32)
p = lambda a: a.ctypes.data_as(FP)


def call(lib, out):
    fn = lib.external_vel_solver_f_; fn.restype = C.c_float
    n = C.c_int(N)
    a = [C.byref(n)] + [p(sx[k]) for k in range(3)] + [p(ss[k]) for k in range(3)] + [p(sr), C.byref(n)] + [p(tx[k]) for k in range(3)] + [p(out[k]) for k in range(12)]
    t0 = time.perf_counter(); flops = fn(*a); return time.perf_counter() - t0, flops


ours = C.CDLL(os.path.join(ROOT, "onbody_b200", "libbh3dvortgrads_b200.so"))
out_g = [np.zeros(N, np.float32) for _ in range(12)]
call(ours, out_g)                                   # first call: context creation, allocations
ts = []
for _ in range(reps):
    for o in out_g: o[:] = 0
    t, fl = call(ours, out_g); ts.append(t)
rec = {"entry_point": "external_vel_solver_f_ (libbh3dvortgrads_b200.so)", "n": N, "reps": reps, "seconds": min(ts), "seconds_all": [round(v, 4) for v in ts],
       "flops_reported": fl, "host_bytes_in": 10 * N * 4, "host_bytes_out": 12 * N * 4}
refp = os.path.join(ROOT, "oracle", "_ref", "fast", "libref_vortgrad3d.so")
if os.path.exists(refp) and not os.environ.get("ONB_NO_REF"):
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    ref = C.CDLL(refp)
    out_r = [np.zeros(N, np.float32) for _ in range(12)]
    t, flr = call(ref, out_r)
    a = np.stack(out_g).astype(np.float64); b = np.stack(out_r).astype(np.float64)
    rec.update({"reference_seconds": t, "reference_threads": os.cpu_count(), "speedup": t / min(ts), "flops_reported_reference": flr,
                "rel_rms_vs_reference_velocities": float(np.sqrt(((a[:3] - b[:3]) ** 2).sum() / (b[:3] ** 2).sum())),
                "rel_rms_vs_reference_all_12": float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))})
print(json.dumps(rec))
