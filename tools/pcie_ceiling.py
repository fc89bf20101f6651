"""Concurrent pinned-copy ceiling of the box: every rank copies what the end-to-end compression path copies per block
(16 MiB of text host -> device, a 27.2 MB container device -> host), with no kernels in between, for a fixed time.
The aggregate over ranks is what the host side (PCIe uplinks, host memory system) can move; the end-to-end bench
cannot exceed it.  Run alone or under torchrun:

    python tools/pcie_ceiling.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_ceiling.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from text_compression_b200 import multi

IN_BYTES = 16 << 20
OUT_BYTES = 27_160_000      # packed container of a 16 MiB random-byte block (1.625 B per run)
SECONDS = 2.0


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    bound = multi.bind_to_gpu_cpus(local) if os.environ.get("TC_NO_BIND") != "1" else False
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    NB = 4
    h_in = [torch.empty(IN_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(NB)]
    h_out = [torch.empty(OUT_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(NB)]
    for t in h_in:
        t.fill_(7)
    d_in = [torch.empty(IN_BYTES, dtype=torch.uint8, device="cuda") for _ in range(NB)]
    d_out = [torch.zeros(OUT_BYTES, dtype=torch.uint8, device="cuda") for _ in range(NB)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def run(do_in, do_out):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        blocks = 0
        while time.perf_counter() - t0 < SECONDS:
            for k in range(NB):
                if do_in:
                    with torch.cuda.stream(s_in):
                        d_in[k].copy_(h_in[k], non_blocking=True)
                if do_out:
                    with torch.cuda.stream(s_out):
                        h_out[k].copy_(d_out[k], non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()
            blocks += NB
        dt = time.perf_counter() - t0
        v = torch.tensor([blocks / dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v)
        return float(v.item())      # blocks per second, all ranks

    run(True, True)
    res = {}
    for name, a, b in (("h2d_only", True, False), ("d2h_only", False, True), ("both", True, True)):
        bps = run(a, b)
        res[name] = {"blocks_per_s": bps,
                     "h2d_GBps": bps * IN_BYTES / 1e9 if a else 0.0,
                     "d2h_GBps": bps * OUT_BYTES / 1e9 if b else 0.0}
    if rank == 0:
        both = res["both"]
        print(json.dumps({"tool": "pcie_ceiling", "n_gpus": world, "cpu_affinity_bound": bool(bound),
                          "per_block": {"h2d_bytes": IN_BYTES, "d2h_bytes": OUT_BYTES}, "results": res,
                          "e2e_ceiling_text_MBps": both["blocks_per_s"] * IN_BYTES / 1e6,
                          "aggregate_GBps_both_directions": both["h2d_GBps"] + both["d2h_GBps"]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
