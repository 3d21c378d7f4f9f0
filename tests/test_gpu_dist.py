"""The distributed hot path (csrc/comm.cu, dist.cu) verified on ONE GPU: R contexts on cuda:0, one host thread each, joined
by the library's loopback transport (device copies + a host barrier where a multi-GPU run has NCCL). Everything but the
transport is the code an 8-GPU run executes: partition plan, range-restricted builds, leaf-record exchange, own / straddling
upward passes with the strength exchange, sparse planes of the lean memory mode, sharded evaluation. Every array must equal
the single-context run bit for bit. (tools/check_multi.py repeats the comparison over real NCCL on 2-8 GPUs.)"""
import threading

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


def _gpu(physics, n, **kw):
    from onbody_b200.api import GpuSession
    return GpuSession(physics, n, n, **kw)


def _on_all(sessions, fn):
    """fn(rank, session) on one thread per session; the collectives inside the library rendezvous across the threads"""
    errs, outs = [], [None] * len(sessions)

    def work(i):
        try:
            outs[i] = fn(i, sessions[i])
        except BaseException as e:      # noqa: BLE001 - reported below
            errs.append((i, repr(e)))
    ths = [threading.Thread(target=work, args=(i,), daemon=True) for i in range(len(sessions))]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=600)
    assert not errs, errs
    assert all(not t.is_alive() for t in ths), "a rank is stuck in a collective"
    return outs


def _reference_run(physics, n, theta, inputs):
    g = _gpu(physics, n)
    g.set_sources(*inputs); g.set_targets(inputs[0], inputs[1])
    g.make_trees(); g.prepare_eval()
    out = {"src": g.parts(0), "eq": g.parts(2, ("x", "r", "s")), "stree": g.tree(0), "ttree": g.tree(1)}
    tg = g.parts(1, ("x", "r", "gidx"))
    out["eqt"] = g.parts(3, ("x",))["x"]
    if physics != "vortgrad3d":          # the reference has no dual tree for the gradient kernel (onvortgrad3d.cpp:264)
        g.zero_vels(); g.fastsumm(theta); out["fast"] = g.parts(1, ("u",))["u"]; out["fast_stats"] = g.stats(); out["pairs"] = g.last_pairs()
    g.zero_vels(); g.treecode3(theta); out["tc3"] = g.parts(1, ("u",))["u"]
    g.zero_vels(); g.treecode2(theta); out["tc2"] = g.parts(1, ("u",))["u"]
    g.zero_vels(); g.naive(7); out["naive"] = g.parts(1, ("u",))["u"]
    out["tg"] = tg
    g.close()
    return out


@pytest.mark.timeout(900)
@pytest.mark.parametrize("physics,n,world,lean", [("grav3d", 70000, 3, False), ("grav3d", 70000, 2, True), ("vort3d", 40000, 4, False),
                                                   ("vort2dtr", 30000, 3, False), ("grav3d", 300000, 8, True), ("grav3d", 700, 8, False),
                                                   ("vortgrad3d", 20000, 2, False)])
def test_loopback_ranks_reproduce_the_single_context_run(physics, n, world, lean):
    from onbody_b200.api import driver_inputs, comm_init_loopback, shard_range_for, MEM_LEAN
    theta = 1.4
    inputs = driver_inputs(physics, n, True)
    want = _reference_run(physics, n, theta, inputs)
    has_fast = physics != "vortgrad3d"
    sess = [_gpu(physics, n) for _ in range(world)]
    if lean:
        for s in sess:
            s.set_memory_mode(MEM_LEAN)
    comm_init_loopback(sess)
    assert sess[1].comm_info()["transport"] == "loopback" and sess[1].comm_info()["nranks"] == world

    def step(rank, g):
        res = {}
        g.set_sources(*inputs); g.set_targets(inputs[0], inputs[1])
        g.make_trees()
        res["stree"] = g.tree(0); res["ttree"] = g.tree(1)
        if not lean:
            g.prepare_eval()
            res["src"] = g.parts(0); res["eq"] = g.parts(2, ("x", "r", "s"))
        else:
            g.prepare_eval()          # releases the SoA source planes once the tiles exist
        res["tg"] = g.parts(1, ("x", "r", "gidx"))
        res["eqt"] = g.parts(3, ("x",))["x"]
        if has_fast:
            g.zero_vels(); g.fastsumm(theta); res["fast"] = g.parts(1, ("u",))["u"]; res["fast_stats"] = g.stats(); res["pairs"] = g.last_pairs()
        g.zero_vels(); g.treecode3(theta); res["tc3"] = g.parts(1, ("u",))["u"]
        g.zero_vels(); g.treecode2(theta); res["tc2"] = g.parts(1, ("u",))["u"]
        g.zero_vels(); g.naive(7); res["naive"] = g.parts(1, ("u",))["u"]
        return res
    outs = _on_all(sess, step)
    pairs = 0
    for rank, res in enumerate(outs):
        lo, hi = shard_range_for(n, 128, rank, world)
        for k in ("num", "ioffset", "nc", "ns", "nr", "x", "s", "pr"):
            assert bits_equal(res["stree"][k], want["stree"][k]), ("stree", k, rank)
        for k in ("num", "ioffset", "nc", "ns", "nr", "x", "pr"):
            assert bits_equal(res["ttree"][k], want["ttree"][k]), ("ttree", k, rank)
        if not lean:
            for k in ("x", "r", "s"):
                assert bits_equal(res["src"][k], want["src"][k]), ("src", k, rank)
                assert bits_equal(res["eq"][k], want["eq"][k]), ("eq", k, rank)
        assert bits_equal(res["tg"]["gidx"][lo:hi], want["tg"]["gidx"][lo:hi]), rank
        assert bits_equal(res["tg"]["x"][:, lo:hi], want["tg"]["x"][:, lo:hi]), rank
        for k in ("fast", "tc3", "tc2", "naive"):
            if k in res:
                assert bits_equal(res[k][:, lo:hi], want[k][:, lo:hi]), (k, rank)
        if has_fast:
            pairs += res["pairs"]
    if has_fast:
        assert pairs >= want["pairs"]          # ancestors of a shard boundary are evaluated by both neighbours
    for s in sess:
        s.close()


@pytest.mark.timeout(600)
def test_loopback_separate_calls_and_second_step():
    """the drivers' call sequence (make_tree, upward, make_tree, refine, upward) with a communicator attached, twice in a
    row with different particles (a time-stepping caller): second step must not see anything of the first"""
    from onbody_b200.api import driver_inputs, comm_init_loopback, shard_range_for
    n, world, theta = 50000, 3, 1.2
    x, r, s = driver_inputs("grav3d", n, True)
    steps = [(x, r, s), (np.ascontiguousarray(x[:, ::-1] * 0.97), r, np.ascontiguousarray(s[:, ::-1]))]
    wants = []
    for inp in steps:
        g = _gpu("grav3d", n)
        g.set_sources(*inp); g.set_targets(inp[0], inp[1])
        g.make_tree(0); g.upward(0); g.make_tree(1); g.refine(1); g.upward(1)
        g.zero_vels(); g.fastsumm(theta)
        wants.append((g.parts(1, ("u", "gidx")), g.parts(2, ("s",))["s"])); g.close()
    sess = [_gpu("grav3d", n) for _ in range(world)]
    comm_init_loopback(sess)

    def run(rank, g):
        outs = []
        g.set_sliced_inputs(True)        # every context pulls only its 1/world slice of the inputs; the library all-gathers them
        for inp in steps:
            g.set_sources(*inp); g.set_targets(inp[0], inp[1])
            g.make_tree(0); g.upward(0); g.make_tree(1); g.refine(1); g.upward(1)
            g.zero_vels(); g.fastsumm(theta)
            outs.append((g.parts(1, ("u", "gidx")), g.parts(2, ("s",))["s"]))
        return outs
    outs = _on_all(sess, run)
    for rank, per_step in enumerate(outs):
        lo, hi = shard_range_for(n, 128, rank, world)
        for (got, eqs), (want, weqs) in zip(per_step, wants):
            assert bits_equal(got["u"][:, lo:hi], want["u"][:, lo:hi]) and bits_equal(got["gidx"][lo:hi], want["gidx"][lo:hi]), rank
            assert bits_equal(eqs, weqs), rank
    for s_ in sess:
        s_.close()


def test_device_results_added_in_place():
    """onb_add_results_planes with DEVICE pointers: one kernel, += in the caller's order, no host round trip"""
    import torch
    n = 30000
    g = _gpu("vort3d", n); g.init_driver()
    g.make_trees(); g.upward(0); g.zero_vels(); g.treecode3(1.3)
    p = g.parts(1, ("u", "gidx"))
    host = np.full((3, n), 2.0, np.float32)
    g.add_results_original_order(host)
    dev = torch.full((3, n), 2.0, dtype=torch.float32, device="cuda")
    g.add_results_planes([dev[d].data_ptr() for d in range(3)])
    torch.cuda.synchronize()
    want = np.full((3, n), 2.0, np.float32); want[:, p["gidx"].astype(np.int64)] += p["u"]
    assert bits_equal(host, want) and bits_equal(dev.cpu().numpy(), want)
    g.close()


@pytest.mark.timeout(600)
def test_loopback_accum_double_and_refusals():
    """ACCUM = double through the distributed path (fp64 outputs of each rank's shard equal the single-context run), and the
    combinations the library refuses loudly: legacy equivalents with a communicator, ACCUM = double in the lean memory mode"""
    from onbody_b200.api import GpuSession, OnbodyError, driver_inputs, comm_init_loopback, shard_range_for, MEM_LEAN
    n, world, theta = 40000, 3, 1.3
    inputs = driver_inputs("grav3d", n, True)
    ref = GpuSession("grav3d", n, n, accum64=True)
    ref.set_sources(*inputs); ref.set_targets(inputs[0], inputs[1]); ref.make_trees(); ref.prepare_eval()
    ref.zero_vels(); ref.fastsumm(theta); want = ref.results_f64(1)
    ref.zero_vels(); ref.treecode3(theta); want3 = ref.results_f64(1); ref.close()
    sess = [GpuSession("grav3d", n, n, accum64=True) for _ in range(world)]
    comm_init_loopback(sess)

    def run(rank, g):
        g.set_sources(*inputs); g.set_targets(inputs[0], inputs[1]); g.make_trees(); g.prepare_eval()
        g.zero_vels(); g.fastsumm(theta); a = g.results_f64(1)
        g.zero_vels(); g.treecode3(theta); b = g.results_f64(1)
        return a, b
    for rank, (a, b) in enumerate(_on_all(sess, run)):
        lo, hi = shard_range_for(n, 128, rank, world)
        assert bits_equal(a[:, lo:hi], want[:, lo:hi]) and bits_equal(b[:, lo:hi], want3[:, lo:hi]), rank
    for s in sess:
        s.close()
    legacy = [GpuSession("grav3d", n, n, order=-1) for _ in range(2)]
    comm_init_loopback(legacy)
    legacy[0].set_sources(*inputs); legacy[0].set_targets(inputs[0], inputs[1])
    with pytest.raises(OnbodyError):
        legacy[0].make_tree(0)                       # refused before any collective is entered
    for s in legacy:
        s.close()
    lean = [GpuSession("grav3d", n, n, accum64=True) for _ in range(2)]
    for s in lean:
        s.set_memory_mode(MEM_LEAN)
    comm_init_loopback(lean)
    with pytest.raises(OnbodyError):
        lean[0].set_targets(inputs[0], inputs[1])
    for s in lean:
        s.close()


@pytest.mark.timeout(600)
def test_loopback_other_block_size_order_and_unequal_clouds():
    """the distributed path away from the defaults: block size 64, order 3, different source and target clouds of unequal size"""
    from onbody_b200.api import GpuSession, driver_inputs, comm_init_loopback, shard_range_for
    ns, nt, world, theta, block, order = 45000, 31000, 3, 1.3, 64, 3
    xs, rs, ss = driver_inputs("grav3d", ns, True)
    rng = np.random.RandomState(11)
    xt = np.ascontiguousarray(rng.uniform(-1, 1, (3, nt)).astype(np.float32)); rt = np.full(nt, nt ** (-1.0 / 3), np.float32)

    def run(g):
        g.set_sources(xs, rs, ss); g.set_targets(xt, rt); g.make_trees(); g.prepare_eval()
        g.zero_vels(); g.fastsumm(theta); a = g.parts(1, ("u", "gidx"))
        g.zero_vels(); g.treecode3(theta); b = g.parts(1, ("u",))["u"]
        return a, b, g.parts(2, ("s",))["s"]
    ref = GpuSession("grav3d", ns, nt, block=block, order=order)
    want = run(ref); ref.close()
    sess = [GpuSession("grav3d", ns, nt, block=block, order=order) for _ in range(world)]
    comm_init_loopback(sess)
    for rank, (a, b, eq) in enumerate(_on_all(sess, lambda r, g: run(g))):
        lo, hi = shard_range_for(nt, block, rank, world)
        assert bits_equal(a["u"][:, lo:hi], want[0]["u"][:, lo:hi]) and bits_equal(a["gidx"][lo:hi], want[0]["gidx"][lo:hi]), rank
        assert bits_equal(b[:, lo:hi], want[1][:, lo:hi]) and bits_equal(eq, want[2]), rank
    for s_ in sess:
        s_.close()
