"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic in fts_b200.distributed:

* batch sharding: contiguous slices cover the batch exactly once; the gather reassembles them;
* single large 3D integral: the level loop + all-gather + stop test used on the GPUs is driven
  here with a CPU stand-in for the two device steps (the oracle's rank-aware level partial), and
  must reproduce the oracle's unsharded adaptive_integrate_3D result on every rank.
"""
import ctypes as C
import math
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions(fts):
    from fts_b200.distributed import shard_range, shard_indices
    for n in (0, 1, 7, 1000):
        for world in (1, 2, 3, 8):
            for layout in ("contiguous", "cyclic"):
                got = np.concatenate([np.arange(n)[shard_indices(n, r, world, layout)] for r in range(world)])
                assert np.array_equal(np.sort(got), np.arange(n))
    # a linear cost ramp (a sorted sweep): cyclic dealing balances it, contiguous slices do not
    cost = np.linspace(1.0, 3.0, 1 << 12)
    cyc = [cost[shard_indices(cost.size, r, 8, "cyclic")].sum() for r in range(8)]
    con = [cost[shard_indices(cost.size, r, 8, "contiguous")].sum() for r in range(8)]
    assert max(cyc) / min(cyc) < 1.002 and max(con) / min(con) > 2.0
    for n in (0, 1, 7, 8, 1 << 20, 65537):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    for p in (os.path.join(ROOT, "fasttanhsinhquadrature.jl_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fts_b200 import api, distributed
    from fts_oracle import FAMILIES as F, Oracle
    orc = Oracle("strict")
    L = 4
    oc = orc.cache_tensor("f64", 3, L)
    low = np.array([-2.0, -1.0, 0.0]); up = np.array([3.0, 2.0, 4.0])
    rtol, atol = 1e-7, 0.0
    fn = orc.lib.fts_or_adaptive3d_level_partial_f64
    fn.restype = C.c_double
    fn.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_int] * 3
    p = np.zeros(8)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)

    def level_fn(level, r, w):  # CPU stand-in for fts_adaptive_sharded_level_dev
        s = fn(F["F3_EXP"], ptr(p), ptr(low), ptr(up), ptr(oc["ix"]), ptr(oc["iw"]), ptr(oc["xs"]), ptr(oc["ws"]), level, r, w)
        return torch.tensor([s, 0.0], dtype=torch.float64)

    def all_gather(t):
        g = torch.zeros(world * 2, dtype=torch.float64)
        dist.all_gather_into_tensor(g, t)
        return g.view(world, 2)

    state = dict(total=0.0, old=0.0, value=0.0, err=0.0, done=False, level=0, h=0.0, levels_run=0)

    def step_fn(level, parts):  # CPU stand-in for fts_adaptive_sharded_step_dev (same arithmetic)
        if state["done"]:
            return
        state["levels_run"] += 1
        state["total"] += float(parts[:, 0].sum() + parts[:, 1].sum())
        h = oc["tm"][0] * 0.5 if level == 0 else state["h"] * 0.5
        scale = np.prod(0.5 * (up - low))
        res = scale * h ** 3 * state["total"]
        state["h"] = h
        state["level"] = level
        if level > 0:
            state["err"] = abs(res - state["old"])
            if state["err"] <= max(atol, rtol * abs(res)):
                state["done"] = True
        state["old"] = state["value"] = res

    distributed.run_sharded_levels(L, rank, world, level_fn, step_fn, all_gather)

    # batch sharding with the gather, compute step replaced by a CPU function of the local slice
    real_quad = api.quad
    api.quad = lambda f, lo, hi, params=None, **kw: np.asarray(params)[:, 0] * 2.0 + np.asarray(lo)
    try:
        n = 1001
        alpha = np.arange(n, dtype=np.float64).reshape(n, 1)
        lows = np.linspace(0, 1, n)
        f = api.builtin("gauss")
        full = distributed.quad_sharded(f, lows, np.ones(n), params=alpha)
        local, (a, b) = distributed.quad_sharded(f, lows, np.ones(n), params=alpha, gather=False)
        full_cyc = distributed.quad_sharded(f, lows, np.ones(n), params=alpha, layout="cyclic")
        loc_cyc, where_cyc = distributed.quad_sharded(f, lows, np.ones(n), params=alpha, gather=False, layout="cyclic")
        assert where_cyc == (rank, n, world) and np.array_equal(loc_cyc, (alpha[:, 0] * 2.0 + lows)[rank::world])
        assert np.array_equal(full_cyc, alpha[:, 0] * 2.0 + lows)
    finally:
        api.quad = real_quad
    ref = orc.adaptive_nd("f64", 3, F["F3_EXP"], None, low, up, rtol, atol, L, oc)
    out.put((rank, state["value"], state["level"], state["done"], state["levels_run"], ref.value, ref.level, ref.converged,
             bool(np.array_equal(full, alpha[:, 0] * 2.0 + lows)), (a, b), local.shape[0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharded_integral_and_batch():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    v0, v1 = res[0][1], res[1][1]
    assert v0 == v1                                   # identical arithmetic on every rank
    for (_rank, value, level, done, levels_run, ref_value, ref_level, ref_conv, gathered_ok, (a, b), nlocal) in res:
        assert level == ref_level and done == ref_conv
        assert levels_run == ref_level + 1            # later levels are no-ops after convergence
        assert abs(value - ref_value) <= 1e-13 * abs(ref_value)
        assert gathered_ok and nlocal == b - a
    assert res[0][9][1] == res[1][9][0] and res[0][9][0] == 0 and res[1][9][1] == 1001
