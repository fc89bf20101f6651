"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL / NVLink).

Only two places of the hot path shard:
  * independent compression blocks: block b -> rank b mod world, no data-path collective;
  * FM-index queries: the index is built on rank 0, its device image is replicated with ONE
    broadcast, and the query batch is split in contiguous chunks (like parListChunk in
    src/Data/FMIndex.hs:417-422); results are concatenated in input order.
A single block's suffix sort stays on one GPU.  The same helpers run on CPU processes with
the gloo backend for the host-side logic tests (no kernels are called there).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import FmInfo


def world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def bind_to_gpu_cpus(gpu_index: int) -> bool:
    """Pin this process to the CPU cores next to its GPU (NVML's ideal affinity), so that pinned
    host buffers are allocated on, and copied over, the GPU's own PCIe root / NUMA node.  One
    process per GPU; call before allocating pinned memory.  Returns False if NVML is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def blocks_of_rank(n_blocks: int, world_size: int, rank: int):
    """Round-robin block assignment: block b belongs to rank b mod world."""
    return list(range(rank, n_blocks, world_size))


def compress_blocks_sharded(texts, with_mtf: bool = True, ctx=None, packed: bool = False):
    """Config 5 (multi-block compression): this rank compresses blocks rank, rank + world, ...
    with the pipelined multi-block call; no data-path collective.  Returns [(block index, result)]
    for the blocks this rank owns: CompressedBlock records (tc_blocks_encode), or with
    packed=True one block container per block (tc_blocks_encode_packed, a third of the bytes
    over PCIe and several blocks in flight per GPU)."""
    from . import block
    ws, rk = world()
    mine = blocks_of_rank(len(texts), ws, rk)
    fn = block.compress_blocks_packed if packed else block.compress_blocks
    return list(zip(mine, fn([texts[b] for b in mine], with_mtf, ctx)))


def query_slice(q: int, world_size: int, rank: int):
    """Contiguous chunk [lo, hi) of q queries for `rank` (sizes differ by at most one)."""
    base, rem = divmod(q, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_in_order(local: np.ndarray, q: int):
    """All ranks contribute their contiguous chunk; every rank gets the full array back in
    input order (order must match src/Data/FMIndex.hs:423,553)."""
    import torch
    import torch.distributed as dist
    ws, rk = world()
    if ws == 1:
        return local
    backend = dist.get_backend()
    dev = "cuda" if backend == "nccl" else "cpu"
    sizes = [query_slice(q, ws, r)[1] - query_slice(q, ws, r)[0] for r in range(ws)]
    mx = max(sizes) if sizes else 0
    buf = torch.zeros(mx, dtype=torch.int64, device=dev)
    buf[: local.size] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.int64)).to(dev)
    outs = [torch.zeros(mx, dtype=torch.int64, device=dev) for _ in range(ws)]
    dist.all_gather(outs, buf)
    return np.concatenate([o[: sizes[r]].cpu().numpy() for r, o in enumerate(outs)])


class ReplicatedFM:
    """A tc_fm handle opened over a torch-owned device buffer (the broadcast image)."""

    def __init__(self, ctx, handle, keepalive):
        self.ctx, self.h, self._keep = ctx, handle, keepalive
        self.info = FmInfo()
        ctx.L.tc_fm_get_info(handle, C.byref(self.info))

    def close(self):
        if self.h is not None and self.h.value:
            self.ctx.L.tc_fm_free(self.h)
            self.h = C.c_void_p(None)
        self._keep = None


class _CudaView:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, addr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (addr, False), "version": 2}


def build_replicated(ctx, d_text, n: int, sa_sample_rate: int = 32) -> ReplicatedFM:
    """Build the FM-index of the device-resident text on rank 0 and replicate it on every rank
    with one NCCL broadcast of the index image over NVLink."""
    import torch
    import torch.distributed as dist
    ws, rk = world()
    h = C.c_void_p(None)
    if ws == 1:
        ctx.call("tc_fm_build_dev", C.c_void_p(d_text.data_ptr()), n, sa_sample_rate, C.byref(h))
        return ReplicatedFM(ctx, h, None)
    size = torch.zeros(1, dtype=torch.int64, device="cuda")
    root = None
    if rk == 0:
        ctx.call("tc_fm_build_dev", C.c_void_p(d_text.data_ptr()), n, sa_sample_rate, C.byref(h))
        root = ReplicatedFM(ctx, h, None)
        size[0] = int(root.info.blob_bytes)
    dist.broadcast(size, 0)
    nbytes = int(size.item())
    if rk == 0:
        img = torch.as_tensor(_CudaView(ctx.L.tc_fm_blob(h), nbytes), device="cuda")
        dist.broadcast(img, 0)
        torch.cuda.synchronize()
        return root
    img = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dist.broadcast(img, 0)
    torch.cuda.synchronize()
    ctx.call("tc_fm_from_blob_dev", C.c_void_p(img.data_ptr()), nbytes, 0, C.byref(h))
    return ReplicatedFM(ctx, h, img)
