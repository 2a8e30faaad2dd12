"""Multi-GPU paths (one process per GPU, torch.distributed for the plumbing).

The reference has no parallelism beyond SIMD lanes; the two decompositions below are
what BASELINE.json's north_star asks for:

* :func:`shard_range` / :func:`quad_sharded` — a batch of independent integrals is cut into
  contiguous slices, one per rank.  There is NO data-path collective: each rank integrates its
  slice; results are gathered only if the caller asks for the full vector.
* :func:`adaptive_integrate_sharded` — ONE large 2D/3D adaptive integral: every rank sums its
  share of a level's new cells (flat cell index dealt cyclically over ranks x CTAs), the ranks
  exchange their compensated (s, c) partials with an all-gather of 2 elements, and each rank
  evaluates the same stop test on the device (adaptive.jl:693-700).  All levels are queued
  without a host round trip; launches after convergence exit immediately.

The orchestration takes its two device steps as callables so that the CPU (gloo, world_size 2)
tests can drive exactly this code with a CPU stand-in for the kernels.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from . import api
from . import _lib as L


def shard_range(batch: int, rank: int, world: int) -> tuple:
    """contiguous slice [start, stop) of `batch` items owned by `rank` (sizes differ by <= 1)"""
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_indices(batch: int, rank: int, world: int, layout: str = "contiguous"):
    """Items of `rank`: a contiguous slice, or — layout="cyclic" — every world-th item starting at
    `rank`.  Cyclic dealing equalises the COST per rank when the refinement depth varies smoothly
    along the batch (a sorted parameter sweep: SURVEY 8e "cost-aware" partitioning) at the price of
    strided host access."""
    if layout == "cyclic":
        return slice(rank, batch, world)
    if layout != "contiguous":
        raise ValueError("layout must be 'contiguous' or 'cyclic'")
    start, stop = shard_range(batch, rank, world)
    return slice(start, stop)


def _dist():
    import torch.distributed as dist
    return dist


def quad_sharded(f, low, up, *, params=None, gather: bool = True, group=None, layout: str = "contiguous", **kw):
    """`quad` over a batch sharded across the ranks of `group` (no data-path collective).
    layout="contiguous": rank r owns a contiguous slice; "cyclic": items r, r+world, ... (equal
    cost per rank for sorted sweeps).  Returns the full result vector, in the caller's order, on
    every rank when gather=True, else the local results and their slice (start, stop[, step])."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    low_a, up_a = np.asarray(low), np.asarray(up)
    dim = f.dim
    batched_bounds = low_a.ndim == (1 if dim == 1 else 2)
    p = None if params is None else np.asarray(params)
    n = low_a.shape[0] if batched_bounds else (p.shape[0] if p is not None and p.ndim >= 1 else 1)
    mine = shard_indices(n, rank, world, layout)
    nloc = len(range(*mine.indices(n)))
    lo_l = np.ascontiguousarray(low_a[mine]) if batched_bounds else low
    up_l = np.ascontiguousarray(up_a[mine]) if batched_bounds else up
    p_l = np.ascontiguousarray(p[mine]) if (p is not None and p.ndim >= 1 and p.shape[0] == n) else params
    local = np.atleast_1d(api.quad(f, lo_l, up_l, params=p_l, **kw)) if nloc > 0 else np.zeros(0)
    if not gather or world == 1:
        where = (mine.start, mine.stop) if layout == "contiguous" else (mine.start, mine.stop, mine.step)
        return (local, where) if not gather else local
    parts = [shard_indices(n, r, world, layout) for r in range(world)]
    sizes = [len(range(*q.indices(n))) for q in parts]
    width = max(sizes)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    buf = torch.zeros(width, dtype=torch.float64, device=dev)
    buf[: local.shape[0]] = torch.from_numpy(np.asarray(local, np.float64))
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    full = np.empty(n, np.float64)
    for o, q, m in zip(out, parts, sizes):
        full[q] = o[:m].cpu().numpy()
    return full


def run_sharded_levels(max_levels: int, rank: int, world: int, level_fn: Callable, step_fn: Callable, all_gather: Callable,
                       shard_from_level: int = 0):
    """The level loop shared by the GPU path and the CPU (gloo) tests.

    level_fn(level, rank, world) -> this rank's partial of the level, a [2] tensor (s, c)
    all_gather(partial)          -> [world, 2] tensor in rank order
    step_fn(level, partials)     -> updates the state (stop test), returns nothing
    Levels below `shard_from_level` are too small to be worth an exchange (level 3 of a 3D
    integral has 3.7k cells): every rank computes them whole (rank 0 of 1) and skips the gather,
    which keeps all ranks bit-identical and saves one collective latency per such level.
    No host synchronisation happens here."""
    for level in range(max_levels + 1):
        if level < shard_from_level:
            step_fn(level, level_fn(level, 0, 1).view(1, 2))
        else:
            step_fn(level, all_gather(level_fn(level, rank, world)))


class PeerMailboxes:
    """NVLink peer-memory mailboxes of the fused exchange (fts_adaptive_sharded_fused_dev): every
    rank owns one small device buffer; its cudaIpc handle is all-gathered once and every peer maps
    it, after which the level kernels store their partials straight into each other's memory."""

    _cache: dict = {}

    def __init__(self, group=None):
        import ctypes as C
        import torch
        dist = _dist()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        nbytes = C.c_size_t()
        L.check(L.lib.fts_peer_mailbox_bytes(self.world, C.byref(nbytes)))
        own = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        L.check(L.lib.fts_peer_alloc(nbytes.value, C.byref(own), handle))
        self.own = own
        self.ptrs = (C.c_void_p * self.world)()
        self.ptrs[self.rank] = own.value
        self.opened = []
        if self.world > 1:
            dev = torch.device("cuda", torch.cuda.current_device())
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=dev)
            allh = torch.zeros(self.world * 64, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine, group=group)
            allh = allh.cpu().numpy().reshape(self.world, 64)
            for r in range(self.world):
                if r == self.rank:
                    continue
                hb = (C.c_ubyte * 64)(*allh[r].tolist())
                ptr = C.c_void_p()
                L.check(L.lib.fts_peer_open(hb, C.byref(ptr)))
                self.ptrs[r] = ptr.value
                self.opened.append(ptr)
            dist.barrier(group=group)
        self.epoch = 0
        self.barriers = 0

    def barrier(self, stream=None) -> None:
        """queue a device-side barrier of the ranks' streams (fts_peer_barrier_dev); every rank must call it"""
        import ctypes as C
        import torch
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        self.barriers += 1
        L.check(L.lib.fts_peer_barrier_dev(self.ptrs, self.rank, self.world, C.c_uint64(self.barriers), C.c_void_p(stream)))

    @classmethod
    def get(cls, group=None) -> "PeerMailboxes":
        key = id(group)
        if key not in cls._cache:
            cls._cache[key] = cls(group)
        return cls._cache[key]


class ShardedPlan:
    """Device buffers and launch closure of ONE sharded 2D/3D integral, so that a caller (bench.py)
    can queue the same integral repeatedly between CUDA events without host work in between.
    `launch()` queues every level on torch's current stream and returns immediately; `result()`
    synchronises and reads the 8-element state."""

    def __init__(self, T, f, low, up, *, rtol=None, atol=0, max_levels: int = 5, cache=None, params=None, group=None,
                 shard_from_level: Optional[int] = None, exchange: str = "peer", world: Optional[int] = None, rank: Optional[int] = None,
                 force_depth: bool = False):
        import torch
        dist = _dist()
        self.dt = dt = api._dt(T)
        if dt == api.Double64:
            raise api.FtsError(L.ERR_UNSUPPORTED, "sharded integrals: float32/float64")
        # world/rank may be forced to 1/0: every rank then computes the WHOLE integral without any
        # exchange (the N = 1 arm of a strong-scaling measurement, run on all GPUs at once)
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.group, self.f, self.exchange, self.max_levels = group, f, exchange, max_levels
        self.options = L.OPT_FORCE_DEPTH if force_depth else 0
        self.rtol_T, self.atol_T = api._resolve_tolerances(dt, rtol, atol)
        self.dim = f.dim
        if cache is None:
            cache = api.adaptive_cache_3D(dt, max_levels=max_levels) if self.dim == 3 else api.adaptive_cache_2D(dt, max_levels=max_levels)
        else:
            api._require_cache_levels(cache, max_levels)
        self.cache = cache
        tdt = torch.float32 if dt == np.dtype(np.float32) else torch.float64
        dev = torch.device("cuda", torch.cuda.current_device())
        self.lo = torch.tensor(np.asarray(low, dt), dtype=tdt, device=dev)
        self.hi = torch.tensor(np.asarray(up, dt), dtype=tdt, device=dev)
        self.p = torch.tensor(np.asarray(params, dt), dtype=tdt, device=dev) if (params is not None and f.nparams) else None
        self.partial = torch.zeros(2, dtype=tdt, device=dev)
        self.gathered = torch.zeros(self.world, 2, dtype=tdt, device=dev)
        self.state = torch.zeros(8, dtype=tdt, device=dev)
        if shard_from_level is None:
            shard_from_level = (max_levels + 1 if exchange == "peer" else 0) if self.world == 1 else (5 if self.dim == 3 else 8)
        self.shard_from_level = shard_from_level
        # the mailboxes belong to the process group; a plan forced to one rank never exchanges
        self.mb = PeerMailboxes.get(group) if (exchange == "peer" and world is None) else None
        if self.mb is None and exchange == "peer":
            self.shard_from_level = max_levels + 1

    def launch(self) -> None:
        import ctypes as C
        import torch
        stream = torch.cuda.current_stream().cuda_stream
        lo, hi, p, state, cache, f = self.lo, self.hi, self.p, self.state, self.cache, self.f
        if self.exchange == "peer":
            mb = self.mb
            epoch = 1
            if mb is not None:
                mb.epoch += 1
                epoch = mb.epoch
            L.check(L.lib.fts_adaptive_sharded_fused_opts_dev(
                cache.handle, f._h, C.c_void_p(lo.data_ptr()), C.c_void_p(hi.data_ptr()), C.c_void_p(p.data_ptr()) if p is not None else None,
                api._ptr(self.rtol_T), api._ptr(self.atol_T), self.max_levels, self.shard_from_level, self.rank, self.world,
                mb.ptrs if mb is not None else None, C.c_uint64(epoch), C.c_void_p(state.data_ptr()), self.options, C.c_void_p(stream)))
            return
        dist = _dist()
        partial, gathered, world, group = self.partial, self.gathered, self.world, self.group

        def level_fn(level, r, w):
            api.sharded_level_dev(cache, f, low=lo.data_ptr(), up=hi.data_ptr(), params=p.data_ptr() if p is not None else 0, level=level,
                                  rank=r, nranks=w, partial=partial.data_ptr(), state=state.data_ptr(), stream=stream)
            return partial

        def all_gather(t):
            if world == 1:
                gathered.copy_(t.view(1, 2))
            else:
                dist.all_gather_into_tensor(gathered.view(-1), t, group=group)
            return gathered

        def step_fn(level, parts):
            api.sharded_step_dev(cache, low=lo.data_ptr(), up=hi.data_ptr(), rtol=self.rtol_T[0], atol=self.atol_T[0], level=level, nranks=parts.shape[0],
                                 partials=parts.data_ptr(), state=state.data_ptr(), stream=stream, options=self.options, max_levels=self.max_levels)

        run_sharded_levels(self.max_levels, self.rank, world, level_fn, step_fn, all_gather, self.shard_from_level)

    def result(self):
        s = self.state.cpu().numpy()  # the only host synchronisation
        if s[5] == 3:
            raise api.FtsError(L.ERR_CUDA, "sharded integral: a rank did not deliver its level partial in time (peer exchange timed out, FTS_B200_PEER_TIMEOUT_MS)")
        return s

    def exchange_wait_ns(self) -> np.ndarray:
        """per-level time this rank's stop test spent waiting for the slowest peer's flag in the last fused call"""
        import ctypes as C
        out = (C.c_uint64 * 64)()
        L.check(L.lib.fts_peer_wait_ns(self.max_levels + 1, out))
        return np.array(out[: self.max_levels + 1], dtype=np.uint64)

    def level_done_ns(self) -> np.ndarray:
        """GPU globaltimer (ns) at which this rank finished the stop test of each level in the last fused call
        (levels < 32); differences of neighbours are the per-level durations including the exchange"""
        import ctypes as C
        out = (C.c_uint64 * 64)()
        L.check(L.lib.fts_peer_wait_ns(64, out))
        return np.array(out[32: 32 + min(self.max_levels + 1, 32)], dtype=np.uint64)


def adaptive_integrate_sharded(T, f, low, up, *, rtol=None, atol=0, max_levels: int = 5, cache=None, params=None, group=None,
                               warn: bool = True, shard_from_level: Optional[int] = None, exchange: str = "peer", force_depth: bool = False):
    """One large adaptive 2D/3D integral over all ranks of `group`.  Returns an AdaptiveResult
    with one element, identical on every rank.

    exchange="peer": the fused path — level kernels store their partials into every rank's
    mailbox over NVLink peer memory, the stop test waits on the device; ONE C call queues all
    levels.  exchange="nccl": level kernel -> NCCL all-gather -> stop-test kernel per level."""
    plan = ShardedPlan(T, f, low, up, rtol=rtol, atol=atol, max_levels=max_levels, cache=cache, params=params, group=group,
                       shard_from_level=shard_from_level, exchange=exchange, force_depth=force_depth)
    plan.launch()
    s = plan.result()
    dt, dim, rank = plan.dt, plan.dim, plan.rank
    conv = s[5] != 0
    flags = np.array([L.FLAG_CONVERGED if conv else (L.FLAG_MAX_LEVELS if max_levels > 0 else 0)], np.int32)
    res = api.AdaptiveResult(np.array([s[3]], dt), np.array([s[4]], dt), np.array([int(s[6])], np.int32), flags)
    if warn and not conv and max_levels > 0 and rank == 0:
        import warnings
        warnings.warn(f"adaptive_integrate_{dim}D reached max_levels without meeting the requested tolerance. max_levels = {max_levels}, "
                      f"estimated_error = {s[4]}, value = {s[3]}", api.ConvergenceWarning, stacklevel=2)
    return res
