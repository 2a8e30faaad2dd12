#!/usr/bin/env python
"""What limits the END-TO-END figure of bench.py when N ranks share one host?

Run under torchrun (one rank per GPU).  Every rank moves the bytes of one e2e step of workload B
(8 MiB of parameters in, 8 MiB of values + 4 MiB of flags out) and all ranks do it AT ONCE:

  copies   plain cudaMemcpyAsync between a pinned host buffer and device memory on two streams (H2D only,
           D2H only, both directions) — the host link's ceiling, no kernel of ours involved;
  e2e      the library's host-buffer call (fts_adaptive_batch) in its in-place mode (the kernel reads and
           writes the pinned buffers over PCIe) and in its staged mode (FTS_B200_ZEROCOPY=0 semantics are
           fixed per process, so the staged figure comes from a pageable-free emulation: explicit copies
           around the device-buffer call);

for several ways of obtaining the host buffers:
  default        cudaHostAlloc through torch (pin_memory)
  affinity       the same after pinning the rank to its own slice of the CPUs (first-touch NUMA placement)
  wc             cudaHostAllocWriteCombined for the buffer the GPU only READS
  hugepage       2 MiB-aligned anonymous memory with MADV_HUGEPAGE, registered with cudaHostRegister

One JSON line per (variant, measurement) on rank 0; times are wall clock per repetition, max over ranks.
"""
import argparse
import ctypes as C
import json
import mmap
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fasttanhsinhquadrature.jl_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

N = 1 << 20
IN_BYTES, OUT_BYTES = N * 8, N * 12


def _load_cudart():
    """the CUDA runtime torch has already loaded (same primary context)"""
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            lib = C.CDLL(name)
            lib.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
            lib.cudaMemcpyAsync.restype = C.c_int
            lib.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
            lib.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
            return lib
        except OSError:
            continue
    raise RuntimeError("libcudart not found")


CUDART = None


def memcpy_async(dst, src, nbytes, kind, stream):
    rc = CUDART.cudaMemcpyAsync(C.c_void_p(dst), C.c_void_p(src), C.c_size_t(nbytes), kind, C.c_void_p(stream))
    if rc:
        raise RuntimeError(f"cudaMemcpyAsync -> {rc}")


def emit(**kw):
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(kw), flush=True)


def max_over_ranks(x, dev, world):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


class HostBuf:
    """a pinned host buffer as a torch uint8 tensor, obtained in one of several ways"""

    def __init__(self, nbytes, how, cudart):
        self.how, self.nbytes, self.cudart, self.ptr, self.keep = how, nbytes, cudart, None, None
        if how in ("default", "affinity"):
            self.t = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        elif how == "wc":
            p = C.c_void_p()
            rc = cudart.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(0x04 | 0x02))  # WriteCombined | Mapped
            if rc:
                raise RuntimeError(f"cudaHostAlloc(WC) -> {rc}")
            self.ptr = p
            arr = np.ctypeslib.as_array((C.c_ubyte * nbytes).from_address(p.value))
            self.t = torch.from_numpy(arr)
        elif how == "hugepage":
            huge = 2 << 20
            size = (nbytes + huge - 1) // huge * huge + huge
            m = mmap.mmap(-1, size, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
            base = C.addressof(C.c_char.from_buffer(m))
            aligned = (base + huge - 1) // huge * huge
            libc = C.CDLL("libc.so.6", use_errno=True)
            libc.madvise(C.c_void_p(aligned), C.c_size_t((nbytes + huge - 1) // huge * huge), 14)  # MADV_HUGEPAGE
            arr = np.ctypeslib.as_array((C.c_ubyte * nbytes).from_address(aligned))
            arr[:] = 0  # touch: pages (hopefully huge) are allocated here, on this rank's CPUs
            rc = cudart.cudaHostRegister(C.c_void_p(aligned), C.c_size_t(nbytes), C.c_uint(0x02))  # Mapped
            if rc:
                raise RuntimeError(f"cudaHostRegister -> {rc}")
            self.ptr, self.keep = C.c_void_p(aligned), m
            self.t = torch.from_numpy(arr)
        else:
            raise ValueError(how)

    def data_ptr(self):
        return self.t.data_ptr()


def copies(dev, world, src, dst, reps):
    d_in = torch.empty(IN_BYTES, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(OUT_BYTES, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    cudart = CUDART
    out = {}
    for mode in ("h2d", "d2h", "both"):
        for timed in (False, True):
            barrier(world)
            t0 = time.perf_counter()
            for _ in range(reps):
                if mode != "d2h":
                    memcpy_async(d_in.data_ptr(), src.data_ptr(), IN_BYTES, 1, s_in.cuda_stream)
                if mode != "h2d":
                    memcpy_async(dst.data_ptr(), d_out.data_ptr(), OUT_BYTES, 2, s_out.cuda_stream)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        dt = max_over_ranks(dt, dev, world)
        nbytes = (IN_BYTES if mode != "d2h" else 0) + (OUT_BYTES if mode != "h2d" else 0)
        out[mode] = {"ms": dt / reps * 1e3, "gbs_all_ranks": nbytes * world * reps / dt / 1e9, "gbs_per_rank": nbytes * reps / dt / 1e9}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=100)
    ap.add_argument("--variants", default="default,affinity,wc,hugepage")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import fts_b200 as fts
    from fts_b200 import _lib as L
    fts.init(local)
    global CUDART
    CUDART = _load_cudart()
    cudart = CUDART
    cpus = sorted(os.sched_getaffinity(0))
    emit(what="host", world=world, cpus_visible=len(cpus), cpu_list=f"{cpus[0]}-{cpus[-1]}",
         numa_nodes=sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")) if os.path.isdir("/sys/devices/system/node") else None,
         thp=open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip() if os.path.exists("/sys/kernel/mm/transparent_hugepage/enabled") else None)

    cache = fts.adaptive_cache_1D(np.float64, max_levels=16)
    gauss = fts.builtin("gauss")
    alpha = 0.5 + 49.5 * np.arange(N, dtype=np.float64) / (N - 1)
    h_lo = torch.tensor([-1.0], dtype=torch.float64).pin_memory()
    h_hi = torch.tensor([1.0], dtype=torch.float64).pin_memory()
    # device-resident reference of the kernel alone
    d_alpha = torch.from_numpy(alpha).to(dev)
    d_lo, d_hi = h_lo.to(dev), h_hi.to(dev)
    d_val = torch.empty(N, dtype=torch.float64, device=dev)
    d_flg = torch.empty(N, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def dev_step():
        fts.adaptive_batch_dev(cache, gauss, low=d_lo.data_ptr(), up=d_hi.data_ptr(), bounds_stride=0, params=d_alpha.data_ptr(), params_stride=1,
                               rtol=1e-10, atol=0.0, max_levels=16, batch=N, value=d_val.data_ptr(), flags=d_flg.data_ptr(), stream=stream,
                               order="reference", options=L.OPT_QUAD_EDGE_RULES)

    for _ in range(5):
        dev_step()
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        dev_step()
    torch.cuda.synchronize()
    emit(what="kernel_only", ms=max_over_ranks(time.perf_counter() - t0, dev, world) / args.reps * 1e3)

    for how in args.variants.split(","):
        if how == "affinity":
            per = max(1, len(cpus) // world)
            mine = cpus[rank * per:(rank + 1) * per] or cpus
            os.sched_setaffinity(0, mine)
        try:
            src = HostBuf(IN_BYTES, how, cudart)
            dst = HostBuf(OUT_BYTES, "default" if how == "wc" else how, cudart)  # the GPU WRITES dst: write-combining is for the read side only
        except Exception as exc:
            emit(what="variant", variant=how, error=str(exc))
            continue
        src.t.view(torch.float64)[:N].copy_(torch.from_numpy(alpha))
        emit(what="copies", variant=how, **copies(dev, world, src, dst, args.reps))
        # the library's host-buffer call on these buffers (in place: the kernel reads src and writes dst over PCIe)
        val_ptr, flg_ptr = dst.data_ptr(), dst.data_ptr() + N * 8

        def e2e_step():
            fts.adaptive_batch_host(cache, gauss, low=h_lo.data_ptr(), up=h_hi.data_ptr(), bounds_stride=0, params=src.data_ptr(), params_stride=1,
                                    rtol=1e-10, atol=0.0, max_levels=16, batch=N, value=val_ptr, flags=flg_ptr, order="reference",
                                    options=L.OPT_QUAD_EDGE_RULES)

        t_w = time.perf_counter()
        while time.perf_counter() - t_w < 1.0:
            e2e_step()
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(args.reps):
            e2e_step()
        torch.cuda.synchronize()
        ms = max_over_ranks(time.perf_counter() - t0, dev, world) / args.reps * 1e3
        ok = bool(np.array_equal(dst.t.view(torch.float64)[:N].numpy(), d_val.cpu().numpy()))
        emit(what="e2e_in_place", variant=how, ms=ms, integrals_per_s_all_ranks=N * world / ms * 1e3, host_gbs_all_ranks=(IN_BYTES + OUT_BYTES) * world / ms / 1e6,
             matches_device_path=ok)
        # staged: explicit copies around the device-buffer call, three streams chained by events (H2D | kernel | D2H per step)
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(args.reps):
            memcpy_async(d_alpha.data_ptr(), src.data_ptr(), IN_BYTES, 1, stream)
            dev_step()
            memcpy_async(val_ptr, d_val.data_ptr(), N * 8, 2, stream)
            memcpy_async(flg_ptr, d_flg.data_ptr(), N * 4, 2, stream)
            torch.cuda.synchronize()
        ms = max_over_ranks(time.perf_counter() - t0, dev, world) / args.reps * 1e3
        emit(what="e2e_staged_serial", variant=how, ms=ms, integrals_per_s_all_ranks=N * world / ms * 1e3)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
