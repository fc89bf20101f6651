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
    """Pin this process to CPU cores next to its GPU (NVML's ideal affinity), so that pinned host buffers are
    allocated on, and copied over, the GPU's own PCIe root / NUMA node.  When several GPUs report the SAME
    affinity set (one NUMA node for all eight GPUs on the measured box: every rank and its three lane threads
    landed on the same 32 cores), the set is cut into disjoint slices, one per GPU that shares it, so the
    ranks' copy-issuing threads do not fight for the same cores.  One process per GPU; call before allocating
    pinned memory.  Returns False if NVML is unavailable or nothing useful can be done."""
    import os
    mode = os.environ.get("TC_B200_BIND", "slice")   # slice (default) | shared (NVML's set as is) | none
    if mode == "none":
        return False
    try:
        import pynvml
        pynvml.nvmlInit()
        ncpu = os.cpu_count() or 1
        nw = (ncpu + 63) // 64
        allowed = set(os.sched_getaffinity(0))

        def affinity(i):
            words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(i), nw)
            return frozenset({64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1} & allowed)

        mine = affinity(gpu_index)
        if not mine:
            return False
        sharers = [i for i in range(pynvml.nvmlDeviceGetCount()) if affinity(i) == mine]
        cpus = sorted(mine)
        if len(sharers) > 1 and mode == "slice":
            per = len(cpus) // len(sharers)
            if per >= 4:                       # fewer than 4 cores per rank: leave the scheduler alone
                j = sharers.index(gpu_index)
                cpus = cpus[j * per:(j + 1) * per]
        os.sched_setaffinity(0, set(cpus))
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


# ---- single-process multi-GPU through the C ABI (tc_mgpu_*, tc_fm_replicate) -------------------------------
# What a caller without torch.distributed gets (the Haskell shim's ...P functions): ONE library call fans out over
# the GPUs of the box with a host thread and a pooled context per device.
def device_count() -> int:
    from ._lib import load
    return int(load().tc_device_count())


def _devs(devices):
    devices = list(range(device_count())) if devices is None else list(devices)
    return devices, (C.c_int * len(devices))(*devices)


def compress_blocks_packed_multi(texts, with_mtf: bool = True, devices=None) -> list:
    """tc_mgpu_blocks_encode_packed: block b runs on devices[b % len(devices)]; one container per block."""
    from ._lib import BlockInfo, _raise, load, pinned_empty, TC_OK
    from .seq import to_bytes
    L = load()
    devices, darr = _devs(devices)
    ts = [np.ascontiguousarray(t if isinstance(t, np.ndarray) else np.frombuffer(to_bytes(t), dtype=np.uint8),
                               dtype=np.uint8) for t in texts]
    nb = len(ts)
    if nb == 0:
        return []
    bound = [int(L.tc_packed_bound(t.size)) for t in ts]
    outs = [pinned_empty(b, np.uint8) for b in bound]
    ns = (C.c_uint64 * nb)(*[t.size for t in ts])
    caps = (C.c_uint64 * nb)(*bound)
    nbytes = (C.c_uint64 * nb)()
    tp = (C.c_void_p * nb)(*[t.ctypes.data for t in ts])
    op = (C.c_void_p * nb)(*[o.ctypes.data for o in outs])
    infos = (BlockInfo * nb)()
    rc = L.tc_mgpu_blocks_encode_packed(len(devices), darr, nb, tp, ns, 1 if with_mtf else 0, op, caps, nbytes, infos)
    if rc != TC_OK:
        _raise(None, rc)
    return [outs[b][: int(nbytes[b])].copy() for b in range(nb)]


class FMReplicas:
    """One replica of an FM-index per device (tc_fm_replicate: peer copies of the image), queried with the
    patterns split in contiguous chunks (tc_mgpu_fm_count / tc_mgpu_fm_locate); results in input order."""

    def __init__(self, fm, devices=None):
        from ._lib import _raise, load, TC_OK
        self.L = load()
        self.devices, self.darr = _devs(devices)
        nd = len(self.devices)
        self.reps = (C.c_void_p * nd)()
        rc = self.L.tc_fm_replicate(fm.h, nd, self.darr, self.reps)
        if rc != TC_OK:
            _raise(None, rc)

    def count_many(self, pats) -> np.ndarray:
        from ._lib import _raise, ptr, TC_OK
        from .fmindex import pack_patterns
        flat, off = pack_patterns(pats)
        q = off.size - 1
        out = np.full(q, -1, dtype=np.int64)
        if q:
            rc = self.L.tc_mgpu_fm_count(len(self.devices), self.darr, self.reps, ptr(flat), ptr(off), q, ptr(out))
            if rc != TC_OK:
                _raise(None, rc)
        return out

    def locate_many(self, pats):
        from ._lib import _raise, ptr, TC_E_CAP, TC_OK
        from .fmindex import pack_patterns
        flat, off = pack_patterns(pats)
        q = off.size - 1
        hit_off = np.zeros(q + 1, dtype=np.uint64)
        total = C.c_uint64(0)
        if q == 0:
            return hit_off, np.empty(0, dtype=np.uint64)
        nd = len(self.devices)
        rc = self.L.tc_mgpu_fm_locate(nd, self.darr, self.reps, ptr(flat), ptr(off), q, ptr(hit_off), None, 0, C.byref(total))
        if rc not in (TC_OK, TC_E_CAP):
            _raise(None, rc)
        pos = np.empty(int(total.value), dtype=np.uint64)
        if pos.size:
            rc = self.L.tc_mgpu_fm_locate(nd, self.darr, self.reps, ptr(flat), ptr(off), q, ptr(hit_off), ptr(pos), pos.size,
                                          C.byref(total))
            if rc != TC_OK:
                _raise(None, rc)
        return hit_off, pos

    def close(self):
        for i in range(len(self.devices)):
            if self.reps[i]:
                self.L.tc_fm_free(C.c_void_p(self.reps[i]))
                self.reps[i] = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
