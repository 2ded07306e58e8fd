"""Host-side plumbing around the matcher modules: chunked H2D -> kernels -> D2H pipelining on a
ring of CUDA streams, and contiguous sharding of a batch of image pairs over ranks (one process
per GPU, no collective on the data path -- image pairs are independent).
"""
from __future__ import annotations

import os
import re
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def _cpu_list(text: str) -> set:
    out = set()
    for part in text.strip().split(","):
        if part:
            lo, _, hi = part.partition("-")
            out.update(range(int(lo), int(hi or lo) + 1))
    return out


def bind_host_to_gpu(device_index: int) -> dict:
    """One process per GPU: run this process on the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned staging
    buffers are allocated (first touch then places them on that node, so the H2D / D2H copies of the ranks do not cross
    the socket interconnect).  Does nothing where the box exposes no such topology (one node, or numa_node = -1 as in a
    VM).  Returns what it found and did (for the bench line)."""
    info = {"bound": False}
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        info["pci"] = bus
        nodes = [d for d in os.listdir("/sys/devices/system/node") if re.fullmatch(r"node\d+", d)]
        info["numa_nodes"] = len(nodes)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read())
        info["gpu_numa_node"] = node
        if node >= 0 and len(nodes) > 1:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus = _cpu_list(f.read()) & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["bound"], info["cpus"] = True, len(cpus)
    except (OSError, ValueError, AttributeError, RuntimeError) as e:
        info["error"] = f"{type(e).__name__}: {e}"
    return info


def shard_bounds(num_pairs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of the pair batch owned by `rank`; sizes differ by at most one."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size: {rank}/{world_size}")
    base, extra = divmod(num_pairs, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_to_rank0(local: Sequence[torch.Tensor], group=None) -> List[torch.Tensor] | None:
    """Host gather of per-rank result tensors (CPU tensors, ragged along dim 0).  Rank 0 gets the
    concatenation in rank order, other ranks get None.  This is the only cross-rank step of the path."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(local)
    world = dist.get_world_size(group)
    parts = [None] * world if dist.get_rank(group) == 0 else None
    dist.gather_object([t.cpu() for t in local], parts, dst=0, group=group)
    if parts is None:
        return None
    return [torch.cat([p[i] for p in parts], dim=0) for i in range(len(local))]


class HostBatchMatcher:
    """Run a matcher module over HOST-resident image pairs.

    The batch is cut into chunks; chunk c uses stream c % n_streams, so the H2D copy of one chunk,
    the kernels of the previous one and the D2H copy of the one before overlap.  Inputs should be
    pinned (they are pinned on first use otherwise); outputs are pinned host tensors.

    ``depth`` result sets are kept (default 2) and calls rotate through them, so consecutive calls
    overlap as well: call k+1 may start copying in while call k is still copying out.  With
    ``join=True`` (default) the caller's current stream waits for the whole call, i.e. the usual
    torch semantics (results are valid once the current stream is synchronised); with ``join=False``
    nothing waits -- call :meth:`synchronize` (or wait on the returned tensors' producer streams)
    before reading results; a result set is overwritten ``depth`` calls later.

    uint8 host images are accepted (4x less PCIe traffic): the score and integral-image kernels read the bytes
    natively (no widened copy exists), results are identical to passing the same values as float32.

    Any module whose forward(image1, image2) returns a tuple of tensors with the pair batch as their first dimension
    works: the three-output matchers, ``MatchExtractionWrapper`` (matches only: no (K+1)^2 matrix crosses PCIe) or the
    ``WithFilters`` matcher.  The pinned result buffers are shaped after the first chunk's outputs.
    """

    def __init__(self, model: torch.nn.Module, chunk: int = 16, n_streams: int = 4, device=None, depth: int = 2,
                 join: bool = True):
        self.model = model
        self.chunk = int(chunk)
        self.depth = max(1, int(depth))
        self.join = bool(join)
        self.device = torch.device(device) if device is not None else next(model.buffers()).device
        if self.device.type != "cuda":
            raise RuntimeError("HostBatchMatcher needs the model on a CUDA device (no CPU path)")
        self.streams = [torch.cuda.Stream(self.device) for _ in range(n_streams)]
        self._out = [None] * self.depth
        self._calls = 0
        self._chunks = 0          # chunks issued so far: the stream ring keeps turning across calls

    def _outputs(self, slot: int, B: int, outs, n: int):
        """Pinned host buffers of result set `slot`, shaped (B, ...) after the device outputs of one chunk of n pairs."""
        cur = self._out[slot]
        ok = cur is not None and len(cur) == len(outs) and all(
            h.shape[0] == B and h.shape[1:] == o.shape[1:] and h.dtype == o.dtype for h, o in zip(cur, outs))
        if not ok:
            for o in outs:
                if o.dim() == 0 or o.shape[0] != n:
                    raise RuntimeError("HostBatchMatcher: every model output must have the pair batch as its first "
                                       f"dimension, got {tuple(o.shape)} for a chunk of {n} pairs")
            cur = tuple(torch.empty((B,) + tuple(o.shape[1:]), dtype=o.dtype).pin_memory() for o in outs)
            self._out[slot] = cur
        return cur

    def synchronize(self) -> None:
        for s in self.streams:
            s.synchronize()

    @torch.no_grad()
    def __call__(self, image1: torch.Tensor, image2: torch.Tensor):
        if image1.is_cuda or image2.is_cuda:
            raise RuntimeError("HostBatchMatcher takes host tensors; call the module directly for device tensors")
        if not image1.is_pinned():
            image1 = image1.pin_memory()
        if not image2.is_pinned():
            image2 = image2.pin_memory()
        B = image1.shape[0]
        slot = self._calls % self.depth
        self._calls += 1
        host = None
        cur = torch.cuda.current_stream(self.device)
        if self.join:
            for s in self.streams:
                s.wait_stream(cur)
        for ci, lo in enumerate(range(0, B, self.chunk)):
            hi = min(lo + self.chunk, B)
            # the ring does not restart with every call: with one chunk per call (chunk >= batch) consecutive calls still
            # land on different streams, so the copies of call k+1 overlap the kernels of call k
            s = self.streams[self._chunks % len(self.streams)]
            self._chunks += 1
            with torch.cuda.stream(s):
                d1 = image1[lo:hi].to(self.device, non_blocking=True)
                d2 = image2[lo:hi].to(self.device, non_blocking=True)
                outs = self.model(d1, d2)          # uint8 stays uint8: the fused kernels read the bytes natively
                outs = (outs,) if torch.is_tensor(outs) else tuple(outs)
                if host is None:
                    host = self._outputs(slot, B, outs, hi - lo)
                for h, o in zip(host, outs):
                    h[lo:hi].copy_(o, non_blocking=True)
        if self.join:
            for s in self.streams:
                cur.wait_stream(s)
        return host


class GraphedMatcher:
    """One matcher step captured into a CUDA graph and replayed (every C-ABI call is capturable: no allocation, no
    synchronisation, the side streams of the fused matcher fork and join by events).  Worth it at small batches, where
    the ~13 launches and 6 event operations of a step are a visible part of its ~0.15 ms: batch 1 drops to ~0.12 ms.

    ``graphed = GraphedMatcher(model, image1, image2)`` captures for the shapes of the two example tensors;
    ``graphed(image1, image2)`` copies new images into the static input buffers, replays and returns the static output
    tensors (overwritten by the next call).  ``graphed.inputs`` are the static buffers: filling them directly (for example
    as the destination of a host-to-device copy) and calling ``graphed.replay()`` saves the device-to-device copies."""

    def __init__(self, model: torch.nn.Module, image1: torch.Tensor, image2: torch.Tensor, warmup: int = 2):
        if not (image1.is_cuda and image2.is_cuda):
            raise RuntimeError("GraphedMatcher captures device work: the example images must be CUDA tensors")
        self.model = model
        self.inputs = (image1.clone(), image2.clone())
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(image1.device)
        side.wait_stream(torch.cuda.current_stream(image1.device))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(1, warmup)):          # creates the library's side streams and warms the allocator
                model(*self.inputs)
            with torch.cuda.graph(self.graph, stream=side):
                outs = model(*self.inputs)
        torch.cuda.current_stream(image1.device).wait_stream(side)
        self.outputs = (outs,) if torch.is_tensor(outs) else tuple(outs)

    def replay(self):
        self.graph.replay()
        return self.outputs

    def __call__(self, image1: torch.Tensor, image2: torch.Tensor):
        for dst, src in zip(self.inputs, (image1, image2)):
            if src.shape != dst.shape:
                raise RuntimeError(f"GraphedMatcher was captured for images of shape {tuple(dst.shape)}, got {tuple(src.shape)}")
            dst.copy_(src, non_blocking=True)
        return self.replay()
