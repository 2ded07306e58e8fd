#!/usr/bin/env python3
"""Pinned-copy rate of the host link with ordinary vs write-combined pinned host memory, all ranks copying at once.
usage: python -m torch.distributed.run --nproc-per-node N tools/probes/wc_pinned_probe.py"""
import ctypes
import glob
import os
import site

import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
path = [p for sp in site.getsitepackages() for p in glob.glob(sp + "/nvidia/cuda_runtime/lib/libcudart.so.12")][0]
rt = ctypes.CDLL(path)
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
N = 256 << 20
d = torch.empty(N, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream(dev)
sp = ctypes.c_void_p(stream.cuda_stream)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def rate(hptr, h2d, reps=8):
    kind = 1 if h2d else 2
    a, b = (hptr, d.data_ptr()) if not h2d else (d.data_ptr(), hptr)
    rt.cudaMemcpyAsync(ctypes.c_void_p(a), ctypes.c_void_p(b), N, kind, sp)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        rt.cudaMemcpyAsync(ctypes.c_void_p(a), ctypes.c_void_p(b), N, kind, sp)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return N / (float(ms) * 1e-3) / 1e9


for name, flags in (("default", 0), ("write-combined", 4), ("portable", 1), ("default", 0), ("write-combined", 4)):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), N, flags)
    assert rc == 0, rc
    ctypes.memset(p, 1, N)                      # touch every page
    up, down = rate(p.value, True), rate(p.value, False)
    if rank == 0:
        print(f"{world} ranks, {name:15s} pinned: H2D {up:6.1f} GB/s  D2H {down:6.1f} GB/s per GPU (max time over ranks)", flush=True)
    rt.cudaFreeHost(p)
if world > 1:
    dist.destroy_process_group()
