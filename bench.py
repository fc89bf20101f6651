#!/usr/bin/env python3
"""bench.py -- BWT+MTF+RLE throughput (and FM-index count queries/s) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference

A step = one pass of the hot path over one 16 MiB block of synthetic random bytes per GPU
(BASELINE.json configs[1]; the multi-block config 5 partitions such blocks over the GPUs, so
N GPUs process N blocks per step with no data-path collective: weak scaling).

Timed region of `value`: inputs already resident in HBM, K steps = K blocks through one
tc_blocks_encode_dev call (texts and runs stay in HBM, three blocks in flight per GPU; the time of
one block at a time is reported in `config`), CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.  `e2e` is the same metric through the
host-buffer C-ABI call (tc_blocks_encode_packed: text in pinned host memory in, packed block
containers in pinned host memory out) with every block's H2D and D2H inside the timed region;
the record-output call (tc_blocks_encode) and the single-block call are timed beside it.  Inputs rotate over 12 distinct blocks (192 MiB > the 126 MB L2) and every
step streams ~4 GB through HBM, so nothing is L2-resident between steps.

The reference is Haskell and no GHC exists in the image (probed on the GPU box too), so the
reference arm and `cpu_baseline` time oracle/tc_oracle.c, the C restatement of the reference
algorithm ("port"), on the box's host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK = 16 << 20
NBLOCKS = 12
METRIC = "bwt_mtf_rle_MB_per_s"
WORKLOAD = "C2: BWT+MTF+RLE of one 16 MiB synthetic random-byte block per GPU per step"


ALPHABET = "bytes"   # "acgtn" with --workload c5


def gen_block(seed: int, n: int) -> np.ndarray:
    from tests.util import gen_acgtn, gen_bytes
    return gen_acgtn(seed, n) if ALPHABET == "acgtn" else gen_bytes(seed, n)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_row(kernel: str):
    """{traffic, warp_instructions, issue_active_pct} per launch of `kernel` from the committed
    ncu --set full capture of this workload (profiles/, produced by tools/make_profiles.sh)."""
    import csv
    path = os.path.join(ROOT, "profiles", f"r2c_ncu_full_summary_{ALPHABET}.csv")
    short = kernel.split("<")[0].strip()
    try:
        rows = list(csv.reader(open(path)))
        ix = {h: i for i, h in enumerate(rows[0])}
        unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        for r in rows[2:]:
            if short in r[ix["Kernel Name"]]:
                ur, uw = rows[1][ix["dram__bytes_read.sum"]], rows[1][ix["dram__bytes_write.sum"]]
                return {"traffic": float(r[ix["dram__bytes_read.sum"]]) * unit.get(ur, 1.0)
                        + float(r[ix["dram__bytes_write.sum"]]) * unit.get(uw, 1.0),
                        "warp_instructions": float(r[ix["smsp__inst_executed.sum"]]),
                        "issue_active_pct": float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]])}
    except Exception:
        pass
    return {"traffic": None}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled during the timed region.  NVML in-process (no fork:
    forking `nvidia-smi` from a process with gigabytes of pinned mappings stalls the timed loop
    for milliseconds); `nvidia-smi` only if the NVML binding is unavailable."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.idx = gpu_index
        self.stop_flag = threading.Event()
        self.sm, self.mx, self.reasons = [], None, set()
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _sample_nvml(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
                          ("sw_power_cap", 0x4)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [x.strip() for x in out.split(",")]
        self.sm.append(float(r[0]))
        self.mx = float(r[1])
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if r[2 + i].lower().startswith("active"):
                self.reasons.add(name)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.h is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.002 if self.h is not None else 0.2)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": ["unsampled"]}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "nvml" if self.h is not None else "nvidia-smi"}


def cpu_port_time(text: np.ndarray):
    """One pass of the oracle (C restatement of the reference algorithm) over `text`."""
    from oracle import oracle as orc
    t0 = time.perf_counter()
    bwt = orc.bwt_encode(text)
    idx, fin = orc.mtf_encode(bwt)
    cnt, sym = orc.rle_encode(idx.astype(np.int16))
    return time.perf_counter() - t0, int(cnt.size)


def workload_config(n, world, sigma, runs):
    """The keys that define the workload -- identical in both arms (the GPU arm's implementation notes live
    under `impl`, outside `config`)."""
    return {"workload": WORKLOAD, "block_bytes": n, "blocks_per_step": world, "sigma": sigma, "runs_per_block": runs,
            "seed": "0xC2 + 1000 * rank + block" if ALPHABET == "bytes" else "0xC2 + 1000 * rank + block (ACGTN)"}


def run_reference(args):
    """Reference arm: the reference's own (sequential) CPU algorithm, restated in C (oracle/tc_oracle.c; the
    Haskell original cannot be built: no GHC in the image).  Like for like with the GPU arm: every step is one
    FULL 16 MiB block of the same generator and seed (block 0 of rank 0), all three passes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    block = gen_block(0xC2, BLOCK)
    for _ in range(args.warmup):          # warm-up on an eighth of the block (caches, page faults)
        cpu_port_time(block[: BLOCK // 8])
    t, runs = 0.0, 0
    for _ in range(args.steps):
        dt, runs = cpu_port_time(block)
        t += dt
    val = args.steps * block.size / 1e6 / t
    sigma = int(np.unique(block).size) + 1
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "MB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(BLOCK, world, sigma, runs),
        "cpu_baseline": {"value": val, "unit": "MB/s", "cores": 1, "kind": "port",
                         "sample": f"one full {block.size >> 20} MiB block per step (block 0 of rank 0, the block the GPU arm "
                                   "checks itself against); C restatement of the reference (comparison suffix sort, list "
                                   "MTF, sequential RLE), single thread like the reference's toBWT / seqToMTF / seqToRLE, "
                                   "which have no parallel form; Haskell original not buildable (no GHC)"},
        "e2e": {"value": val, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_b200(args):
    import torch
    import torch.distributed as dist
    from text_compression_b200 import _lib
    from text_compression_b200._lib import BlockInfo, ptr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: text_compression_b200 has no CPU path")
    torch.cuda.set_device(local)
    from text_compression_b200 import multi as _multi
    numa_bound = _multi.bind_to_gpu_cpus(local)   # pinned buffers on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    ctx = _lib.Context(local, stream.cuda_stream)
    n = BLOCK

    # ---- inputs: NBLOCKS distinct blocks per rank, resident in HBM, plus pinned host copies
    host_blocks = [gen_block(0xC2 + 1000 * rank + b, n) for b in range(NBLOCKS)]
    d_text = torch.empty((NBLOCKS, n), dtype=torch.uint8, device="cuda")
    for b in range(NBLOCKS):
        d_text[b].copy_(torch.from_numpy(host_blocks[b]))
    cap = n + 3
    NOUT = 8   # run buffers: blocks in flight (at most 4 lanes, claimed in order) never share one
    d_count = [torch.empty(cap, dtype=torch.int32, device="cuda") for _ in range(NOUT)]
    d_rsym = [torch.empty(cap, dtype=torch.int16, device="cuda") for _ in range(NOUT)]
    torch.cuda.synchronize()
    info = BlockInfo()

    def step_dev(i):   # one block, one call (the profile pass and the single-block figure)
        b = i % NBLOCKS
        ctx.call("tc_bwt_mtf_rle_encode_dev", C.c_void_p(d_text[b].data_ptr()), n, C.c_void_p(d_count[0].data_ptr()),
                 C.c_void_p(d_rsym[0].data_ptr()), cap, C.byref(info))

    def steps_dev(first, k):
        """k steps = k blocks through tc_blocks_encode_dev: texts and runs stay in HBM, three blocks in flight
        (block i writes run buffer i % NOUT)."""
        tp = (C.c_void_p * k)(*[d_text[(first + i) % NBLOCKS].data_ptr() for i in range(k)])
        cp = (C.c_void_p * k)(*[d_count[i % NOUT].data_ptr() for i in range(k)])
        sp = (C.c_void_p * k)(*[d_rsym[i % NOUT].data_ptr() for i in range(k)])
        ns = (C.c_uint64 * k)(*([n] * k))
        caps = (C.c_uint64 * k)(*([cap] * k))
        infos = (BlockInfo * k)()
        ctx.call("tc_blocks_encode_dev", k, tp, ns, 1, cp, sp, caps, infos)
        return infos

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        # warm-up: at least W steps and at least a quarter of a second of work, so that every GPU of a
        # multi-GPU run has left its idle clocks (210 MHz) before the timed region starts
        t_w = time.perf_counter()
        i = 0
        while i < args.warmup or time.perf_counter() - t_w < 0.25:
            steps_dev(i, max(args.warmup, 4))   # >= the number of lanes: every lane's context and arena exist
            i += max(args.warmup, 4)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:   # the line reports rank 0's clocks; NVML polling from every rank contends on the driver
            sampler.start()
        launches0 = ctx.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        infos = steps_dev(args.warmup, args.steps)   # returns with both lanes drained
        ev1.record(stream)
        barrier()
        launches = ctx.launches - launches0
        sampler.stop_flag.set()
        if rank == 0:
            sampler.join()
        ms = ev0.elapsed_time(ev1)
        if os.environ.get("TC_BENCH_DEBUG"):
            print(f"[rank {rank}] {ms / args.steps:.4f} ms/step", file=sys.stderr)
        info = infos[args.steps - 1]
        R_last = int(info.R)
        sigma = int(info.sigma)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        value = world * args.steps * n / 1e6 / (ms / 1e3)
        # one block at a time (no second block to fill the serial phases of the kernel chain)
        info = BlockInfo()
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev2.record(stream)
        for i in range(args.steps):
            step_dev(args.warmup + i)
        ev3.record(stream)
        torch.cuda.synchronize()
        single_ms = ev2.elapsed_time(ev3) / args.steps

        # ---- e2e: host buffers through the C-ABI, H2D + D2H inside the timed region.  The call is
        # tc_blocks_encode_packed, the multi-block entry point with container output: one call
        # compresses `steps` blocks from pinned host memory into pinned host memory, overlapping the
        # copies of neighbouring blocks with the kernels of the current one (every block's H2D and
        # D2H happen inside the call).  The container carries the same runs at 2 bytes + 1 bit each
        # (tc_packed_unpack gives the 6-byte records back); the record-output call tc_blocks_encode
        # is timed beside it.
        NH = 8   # >= 2 * lanes: up to 2 * lanes blocks have their D2H pending, each into its own buffer
        pcap = int(ctx.L.tc_packed_bound(n))
        h_in = [_lib.pinned_empty(n, np.uint8) for _ in range(NH)]
        h_out = [_lib.pinned_empty(pcap, np.uint8) for _ in range(NH)]
        h_cnt = [_lib.pinned_empty(cap, np.uint32) for _ in range(NH)]
        h_sym = [_lib.pinned_empty(cap, np.int16) for _ in range(NH)]
        for j in range(NH):
            h_in[j][:] = host_blocks[j]

        def batch(nb, packed=True):
            tp = (C.c_void_p * nb)(*[h_in[b % NH].ctypes.data for b in range(nb)])
            ns = (C.c_uint64 * nb)(*([n] * nb))
            infos = (BlockInfo * nb)()
            if packed:
                op = (C.c_void_p * nb)(*[h_out[b % NH].ctypes.data for b in range(nb)])
                caps = (C.c_uint64 * nb)(*([pcap] * nb))
                nbytes = (C.c_uint64 * nb)()
                ctx.call("tc_blocks_encode_packed", nb, tp, ns, 1, op, caps, nbytes, infos)
                return infos, int(nbytes[nb - 1])
            cp = (C.c_void_p * nb)(*[h_cnt[b % NH].ctypes.data for b in range(nb)])
            sp = (C.c_void_p * nb)(*[h_sym[b % NH].ctypes.data for b in range(nb)])
            caps = (C.c_uint64 * nb)(*([cap] * nb))
            ctx.call("tc_blocks_encode", nb, tp, ns, 1, cp, sp, caps, infos)
            return infos, int(infos[nb - 1].R) * 6

        def timed_batch(packed):
            batch(max(8, args.warmup), packed)   # lanes claim two blocks each: every lane (context, arena, streams) is warm
            barrier()
            t0 = time.perf_counter()
            _, nbytes = batch(args.steps, packed)
            torch.cuda.synchronize()
            e_s = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([e_s], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e_s = float(t.item())
            return world * args.steps * n / 1e6 / e_s, nbytes

        e_steps = args.steps
        e2e_records, d2h_records = timed_batch(False)
        e2e_val, d2h = timed_batch(True)
        # the same call over 4 x steps blocks: the first H2D and the last D2H of a call are not overlapped with
        # anything, which a 10-block call shows as ~15 % and a long one does not
        LONG = 4 * args.steps
        barrier()
        t0 = time.perf_counter()
        batch(LONG, True)
        torch.cuda.synchronize()
        e2e_long = LONG * n / 1e6 / (time.perf_counter() - t0)
        # the container of the last block must unpack to runs that cover its BWT
        last = h_out[(e_steps - 1) % NH]
        uinfo = BlockInfo()
        assert ctx.L.tc_packed_unpack(ptr(last), d2h, ptr(h_cnt[0]), ptr(h_sym[0]), cap, C.byref(uinfo)) == 0
        assert int(h_cnt[0][: int(uinfo.R)].sum(dtype=np.uint64)) == n + 1 and d2h_records == 6 * int(uinfo.R)
        # the single-block call (tc_bwt_mtf_rle_encode), copies not overlapped, for comparison
        sinfo = BlockInfo()
        ctx.call("tc_bwt_mtf_rle_encode", ptr(h_in[0]), n, ptr(h_cnt[0]), ptr(h_sym[0]), cap, C.byref(sinfo))
        t0 = time.perf_counter()
        for i in range(e_steps):
            ctx.call("tc_bwt_mtf_rle_encode", ptr(h_in[i % NH]), n, ptr(h_cnt[i % NH]), ptr(h_sym[i % NH]), cap,
                     C.byref(sinfo))
        torch.cuda.synchronize()
        e2e_single = e_steps * n / 1e6 / (time.perf_counter() - t0)

        # ---- per-kernel timing (CUDA events around every launch) for the roofline
        barrier()
        ctx.profile(True)
        PSTEPS = 2
        for i in range(PSTEPS):
            step_dev(i)
        prof = ctx.profile_report()
        ctx.profile(False)

    peak, peak_src = peaks()
    tot_ms = sum(v[1] for v in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1][1])
    tname, (tn, tms, tbytes) = top
    roofline = {"bound": "hbm", "kernel": tname, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                "traffic": None, "launches_per_step": tn / PSTEPS, "avg_launch_us": 1e3 * tms / tn,
                "share_of_kernel_time": tms / tot_ms, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": tbytes / tn if tn else None}
    if tbytes:
        roofline["achieved"] = tbytes / 1e9 / (tms / 1e3)
        roofline["frac"] = roofline["achieved"] / peak
    roofline.update(ncu_row(tname))   # traffic + the issue-side figures of the same capture (the kernel is issue-bound)
    N = n + 1

    def pass_ms(prefixes):
        return sum(v[1] for k, v in prof.items() if k.split("<")[0].strip().startswith(prefixes)) / PSTEPS

    mtf_ms = pass_ms(("mtf",))
    rle_ms = pass_ms(("rle_",))
    bwt_ms = tot_ms / PSTEPS - mtf_ms - rle_ms
    mtf_bytes = N * 3                     # u8 symbol in, u16 index out
    rle_bytes = N * 2 + 6 * R_last        # u16 index in, (u32 count, i16 symbol) per run out
    passes = {
        "bwt": {"ms": bwt_ms},
        "mtf": {"ms": mtf_ms, "GBps": mtf_bytes / 1e9 / (mtf_ms / 1e3), "frac": mtf_bytes / 1e9 / (mtf_ms / 1e3) / peak},
        "rle": {"ms": rle_ms, "GBps": rle_bytes / 1e9 / (rle_ms / 1e3), "frac": rle_bytes / 1e9 / (rle_ms / 1e3) / peak},
        "kernel_ms_per_step": {k: v[1] / PSTEPS for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
    }

    line = None
    if rank == 0:
        # CPU port on the full block 0 -- the first timed block -- and identity of the GPU's output with it
        from oracle import oracle as orc
        t0 = time.perf_counter()
        o_bwt = orc.bwt_encode(host_blocks[0])
        o_idx, o_fin = orc.mtf_encode(o_bwt)
        o_cnt, o_sym = orc.rle_encode(o_idx.astype(np.int16))
        cpu_s = time.perf_counter() - t0
        chk = BlockInfo()
        g_cnt, g_sym = np.empty(cap, np.uint32), np.empty(cap, np.int16)
        ctx.call("tc_bwt_mtf_rle_encode", ptr(host_blocks[0]), n, ptr(g_cnt), ptr(g_sym), cap, C.byref(chk))
        identical = (int(chk.R) == o_cnt.size and np.array_equal(g_cnt[: int(chk.R)], o_cnt.astype(np.uint32))
                     and np.array_equal(g_sym[: int(chk.R)], o_sym) and list(chk.final_list[: chk.sigma]) == o_fin.tolist()
                     and int(chk.primary) == int(np.nonzero(o_bwt < 0)[0][0]))
        assert identical, "GPU output of the timed block differs from the CPU port"
        line = {
            "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(n, world, sigma, R_last),
            "impl_notes": {"api": "tc_blocks_encode_dev: one call over `steps` HBM-resident blocks, three blocks in flight per GPU (lanes)",
                           "one_block_at_a_time_ms_per_step": single_ms,
                           "warmup_note": "W untimed steps, extended to >= 0.25 s so all GPUs leave idle clocks",
                           "l2": f"inputs rotate over {NBLOCKS} distinct blocks ({NBLOCKS * n >> 20} MiB > L2); "
                                 "each step streams > 1 GB of sort traffic",
                           "identical_to_cpu_port_on_block_0": bool(identical)},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_val, "unit": "MB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": d2h,
                    "api": "tc_blocks_encode_packed: one call over `steps` blocks, pinned host buffers in and out, "
                           "copies of neighbouring blocks overlapped with compute, three blocks in flight (lanes: contexts + "
                           "host threads inside the call); output = packed block container "
                           "(header + runs at 1.625 B each, lossless: tc_packed_unpack returns the records)",
                    "long_call_MBps_this_rank": e2e_long, "long_call_blocks": LONG,
                    "record_output_MBps": e2e_records, "record_output_d2h_bytes_per_step": d2h_records,
                    "single_block_call_MBps": e2e_single, "cpu_affinity_bound_to_gpu": bool(numa_bound)},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "passes": passes,
            "cpu_baseline": {"value": n / 1e6 / cpu_s, "unit": "MB/s", "cores": 1, "kind": "port",
                             "sample": f"the full {n >> 20} MiB block 0, one pass; C restatement of the "
                                       "reference algorithm, single thread (the reference path is sequential)"},
        }
    if args.decode:
        if rank == 0:   # a single-GPU figure: the other ranks would only contend for the host's copy path
            line["decode"] = run_decode(args, ctx, host_blocks[0], peak)
        barrier()
    if line is not None and args.c1:
        line["c1_roundtrip"] = run_c1(ctx)
    if args.fm and world >= 1:
        fm = run_fm(args, ctx, stream, world, rank, local, peak)
        if line is not None:
            line["fm_count"] = fm
    if args.locate:
        torch.cuda.empty_cache()
        loc = run_locate(args, ctx, stream, world, rank, local, peak)
        if line is not None:
            line["fm_locate"] = loc
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_decode(args, ctx, block0, peak):
    """fromRLE -> fromMTF -> fromBWT on 16 MiB blocks (the inverse chain of the headline), both alphabets:
    tc_packed_decode, container in pinned host memory in, text in pinned host memory out (e2e), and the sum of
    the kernels' device times of one call (CUDA events around every launch)."""
    import torch
    from tests.util import gen_acgtn, gen_bytes
    from text_compression_b200 import _lib, block
    from text_compression_b200._lib import ptr
    out = {}
    n = BLOCK
    reps = max(3, min(args.steps, 6))
    for name, text in (("bytes", gen_bytes(0xC2, n)), ("acgtn", gen_acgtn(0xC5, n))):
        blob = block.compress_blocks_packed([text], True, ctx)[0]
        h_blob = _lib.pinned_empty(blob.size, np.uint8)
        h_blob[:] = blob
        h_text = _lib.pinned_empty(n + 2, np.uint8)
        n_out = C.c_uint64(0)

        def once():
            ctx.call("tc_packed_decode", ptr(h_blob), blob.size, ptr(h_text), n + 2, C.byref(n_out))

        for _ in range(3):
            once()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / reps
        assert n_out.value == n and np.array_equal(h_text[:n], text), "decode round trip failed"
        ctx.profile(True)
        once()
        prof = ctx.profile_report()
        ctx.profile(False)
        dev_ms = sum(v[1] for v in prof.values())
        top = sorted(prof.items(), key=lambda kv: -kv[1][1])[:6]
        # multi-block decompression: NB containers through one tc_blocks_decode_packed call (lanes), distinct buffers
        NB = 12
        hb = [_lib.pinned_empty(blob.size, np.uint8) for _ in range(NB)]
        ht = [_lib.pinned_empty(n + 2, np.uint8) for _ in range(NB)]
        for x in hb:
            x[:] = blob
        bp = (C.c_void_p * NB)(*[x.ctypes.data for x in hb])
        by = (C.c_uint64 * NB)(*([blob.size] * NB))
        tp = (C.c_void_p * NB)(*[x.ctypes.data for x in ht])
        cp = (C.c_uint64 * NB)(*([n + 2] * NB))
        no = (C.c_uint64 * NB)()
        ctx.call("tc_blocks_decode_packed", NB, bp, by, tp, cp, no)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.call("tc_blocks_decode_packed", NB, bp, by, tp, cp, no)
        torch.cuda.synchronize()
        batch_s = (time.perf_counter() - t0) / NB
        assert all(int(no[b]) == n for b in range(NB)) and np.array_equal(ht[NB - 1][:n], text), "batch decode failed"
        out[name] = {"device_MBps": n / 1e6 / (dev_ms / 1e3), "device_ms": dev_ms, "e2e_MBps": n / 1e6 / e2e_s,
                     "e2e_batch_MBps": n / 1e6 / batch_s, "e2e_batch_blocks": NB,
                     "h2d_bytes": int(blob.size), "d2h_bytes": n, "launches": int(sum(v[0] for v in prof.values())),
                     "frac_of_hbm": (2 * n + blob.size) / 1e9 / (dev_ms / 1e3) / peak,
                     "top_kernels_ms": {k.split("<")[0]: v[1] for k, v in top}}
    out["workload"] = ("inverse chain on 16 MiB blocks: packed container -> runs -> MTF indices -> BWT -> text; device_* and "
                       "e2e_MBps: one block per tc_packed_decode call; e2e_batch_MBps: 12 containers through one "
                       "tc_blocks_decode_packed call (lanes), pinned host in and out")
    return out


def _fm_inputs(n, q, m, seed, mut_frac):
    """Text and reads of the FM workloads from the splitmix64 generators of tests/util.py, so that
    the CPU side can regenerate them; built in slices to bound host memory."""
    from tests.util import gen_acgtn, gen_reads
    step = 100_000_000
    text = np.concatenate([gen_acgtn(seed + 1000 * i, min(step, n - o)) for i, o in enumerate(range(0, n, step))])
    rstep = 1_000_000
    reads = np.concatenate([gen_reads(seed + 1 + 7 * i, text, min(rstep, q - o), m, mut_frac)
                            for i, o in enumerate(range(0, q, rstep))])
    return text, reads


def _broadcast_reads(reads_host, q, m, world, rank):
    """Rank 0 made the batch; every rank gets the whole of it (one NCCL broadcast) and answers its contiguous chunk."""
    import torch
    import torch.distributed as dist
    d = torch.empty((q, m), dtype=torch.uint8, device="cuda")
    if rank == 0:
        d.copy_(torch.from_numpy(reads_host))
    if world > 1:
        dist.broadcast(d, 0)
    return d


def run_fm(args, ctx, stream, world, rank, local, peak):
    """Config 3: toFMIndex on synthetic ACGTN + countFMIndex for 100-bp reads.  Index built on rank 0 and
    replicated with one NCCL broadcast (timed); ONE batch of reads split into contiguous chunks, one per rank;
    the counts gathered back in input order (timed)."""
    import torch
    import torch.distributed as dist
    from text_compression_b200 import multi
    from text_compression_b200._lib import ptr
    n, q, m = args.fm_n, args.fm_q, 100
    q -= q % world
    text_h = reads_h = None
    if rank == 0:
        text_h, reads_h = _fm_inputs(n, q, m, 0xC3, 0.10)
    with torch.cuda.stream(stream):
        text = torch.empty(n, dtype=torch.uint8, device="cuda")
        if rank == 0:
            text.copy_(torch.from_numpy(text_h))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fm = multi.build_replicated(ctx, text, n, args.fm_rate)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        bcast_ms = None
        if world > 1:      # the broadcast alone, again, on the image that now exists everywhere
            img = torch.as_tensor(multi._CudaView(ctx.L.tc_fm_blob(fm.h), int(fm.info.blob_bytes)), device="cuda")
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            dist.broadcast(img, 0)
            e1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            bcast_ms = float(t.item())
        del text
        all_reads = _broadcast_reads(reads_h, q, m, world, rank)
        lo, hi = multi.query_slice(q, world, rank)
        q_local = hi - lo
        reads = all_reads[lo:hi].contiguous()
        del all_reads
        off = (torch.arange(q_local + 1, device="cuda", dtype=torch.int64) * m).contiguous()
        counts = torch.empty(q_local, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()

        def once():
            ctx.call("tc_fm_count_dev", fm.h, C.c_void_p(reads.data_ptr()), C.c_void_p(off.data_ptr()), q_local,
                     C.c_void_p(counts.data_ptr()))

        for _ in range(max(3, args.warmup)):
            once()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            once()
        ev1.record(stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        # gather in input order (what the caller of ...CountP gets back), timed on its own
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if world > 1:
            outs = [torch.empty(q_local, dtype=torch.int64, device="cuda") for _ in range(world)]
            dist.all_gather(outs, counts)
            all_counts = torch.cat(outs)
        else:
            all_counts = counts
        torch.cuda.synchronize()
        gather_ms = 1e3 * (time.perf_counter() - t0)
        found = int((all_counts >= 0).sum().item())
        # e2e: host patterns in, host counts out, through tc_fm_count (copies inside the call), this rank's chunk
        from text_compression_b200 import _lib
        h_reads = _lib.pinned_empty(q_local * m, np.uint8)
        h_reads[:] = reads.cpu().numpy().reshape(-1)
        h_off = np.arange(q_local + 1, dtype=np.uint64) * m
        h_cnt = _lib.pinned_empty(q_local, np.int64)
        ctx.call("tc_fm_count", fm.h, ptr(h_reads), ptr(h_off), q_local, ptr(h_cnt))
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        reps = max(2, min(args.steps, 5))
        for _ in range(reps):
            ctx.call("tc_fm_count", fm.h, ptr(h_reads), ptr(h_off), q_local, ptr(h_cnt))
        e_s = (time.perf_counter() - t0) / reps
        if world > 1:
            t = torch.tensor([e_s], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_s = float(t.item())
        assert np.array_equal(h_cnt, counts.cpu().numpy())
    qps = world * q_local * args.steps / (ms / 1e3)
    per_query = m + 8 + 2 * (m - 1) * 32
    out = {"metric": "fm_count_queries_per_s", "value": qps, "unit": "queries/s", "n_gpus": world,
           "config": {"workload": f"C3: toFMIndex on {n} bp synthetic ACGTN (splitmix64 seed 0xC3) + countFMIndex, "
                                  f"{q_local * world} reads x {m} bp (10 % with one substitution), one batch split into "
                                  "contiguous chunks per GPU, index replicated by NCCL broadcast",
                      "sa_sample_rate": args.fm_rate},
           "build_s": build_s, "ms_per_batch": ms / args.steps, "found_frac": found / max(q_local * world, 1),
           "broadcast_ms": bcast_ms,
           "broadcast_GBps": (int(fm.info.blob_bytes) / 1e9 / (bcast_ms / 1e3)) if bcast_ms else None,
           "broadcast_frac_of_nvlink_770": (int(fm.info.blob_bytes) / 1e9 / (bcast_ms / 1e3) / 770.0) if bcast_ms else None,
           "gather_ms": gather_ms,
           "e2e": {"value": world * q_local / e_s, "unit": "queries/s", "h2d_bytes_per_batch": int(q_local * m + 8 * (q_local + 1)),
                   "d2h_bytes_per_batch": int(8 * q_local), "api": "tc_fm_count: pinned host patterns in, host counts out"},
           "roofline": {"bound": "hbm", "achieved": qps / world * per_query / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": qps / world * per_query / 1e9 / peak, "bytes_per_query": per_query,
                        "note": "algorithmic 32-byte sectors per query; the 100 Mbp index (198 MB) is served largely from the "
                                "126 MB L2, so this is a fraction of the HBM figure, not HBM traffic (profiles/: dram__bytes "
                                "and lts__t_bytes of fm_count_kernel); fm_locate runs on the 1 Gbp index, which does not fit L2"},
           "index_bytes": int(fm.info.blob_bytes)}
    if rank == 0:
        # CPU baseline on a bounded sample: checkpointed-Occ restatement, all host threads
        from oracle import oracle as orc
        ns, qs = min(n, 4_000_000), min(q_local, 200_000)
        ofm = orc.FMIndexSampled(text_h[:ns])
        from tests.util import gen_reads
        rd = gen_reads(0xC3 + 99, text_h[:ns], qs, m)
        flat = np.ascontiguousarray(rd.reshape(-1))
        offh = (np.arange(qs + 1, dtype=np.uint64) * m)
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        ofm.count_batch(flat, offh, cores)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": qs / dt, "unit": "queries/s", "cores": cores, "kind": "port",
                               "sample": f"{qs} reads against the first {ns} bp (index build not timed); pthreads over "
                                         "contiguous chunks like parListChunk; checkpointed Occ instead of the "
                                         "reference's dense sigma x N table"}
    fm.close()
    return out


def run_locate(args, ctx, stream, world, rank, local, peak):
    """Config 4: locateFMIndex with a sampled suffix array on a synthetic ACGTN reference (splitmix64 seed 0xC4),
    32-bp patterns: one batch split into contiguous chunks per GPU, index replicated with one NCCL broadcast."""
    import torch
    import torch.distributed as dist
    from text_compression_b200 import multi
    n, q, m, rate = args.loc_n, args.loc_q, 32, args.fm_rate
    q -= q % world
    text_h = pats_h = None
    if rank == 0:
        text_h, pats_h = _fm_inputs(n, q, m, 0xC4, 0.0)
    with torch.cuda.stream(stream):
        text = torch.empty(n, dtype=torch.uint8, device="cuda")
        if rank == 0:
            text.copy_(torch.from_numpy(text_h))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fm = multi.build_replicated(ctx, text, n, rate)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        del text
        all_pats = _broadcast_reads(pats_h, q, m, world, rank)
        lo, hi = multi.query_slice(q, world, rank)
        q_local = hi - lo
        pats = all_pats[lo:hi].contiguous()
        del all_pats
        off = (torch.arange(q_local + 1, device="cuda", dtype=torch.int64) * m).contiguous()
        cap = 4 * q_local + 1024
        hit_off = torch.empty(q_local + 1, dtype=torch.int64, device="cuda")
        pos = torch.empty(cap, dtype=torch.int64, device="cuda")
        total = C.c_uint64(0)
        torch.cuda.synchronize()

        def once():
            ctx.call("tc_fm_locate_dev", fm.h, C.c_void_p(pats.data_ptr()), C.c_void_p(off.data_ptr()), q_local,
                     C.c_void_p(hit_off.data_ptr()), C.c_void_p(pos.data_ptr()), cap, C.byref(total))

        for _ in range(3):
            once()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            once()
        ev1.record(stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        hits = int(total.value)
        ok = None
        if rank == 0:
            # every pattern was cut from the text: a sample of them against the brute-force occurrence scan
            from oracle import oracle as orc
            ho = hit_off.cpu().numpy()
            ps = pos[:hits].cpu().numpy()
            sel = np.arange(0, q_local, max(1, q_local // 500))[:500]
            cnt, oho, opos = orc.naive_search(text_h, pats_h[lo:hi][sel], want_pos=True)
            ok = all(np.array_equal(np.sort(ps[ho[i]:ho[i + 1]]), opos[oho[k]:oho[k + 1]].astype(np.int64))
                     for k, i in enumerate(sel.tolist()))
    pps = world * q_local * args.steps / (ms / 1e3)
    per_pat = m + 2 * (m - 1) * 32
    per_hit = 8 + 4 + 32 * (rate - 1) // 2
    out = {"metric": "fm_locate_patterns_per_s", "value": pps, "unit": "patterns/s", "n_gpus": world,
           "config": {"workload": f"C4: locateFMIndex, sampled SA (rate {rate}), {n} bp synthetic ACGTN (splitmix64 seed 0xC4), "
                                  f"{q_local * world} patterns x {m} bp, one batch split into contiguous chunks per GPU",
                      "sa_sample_rate": rate},
           "build_s": build_s, "ms_per_batch": ms / args.steps, "hits_per_pattern": hits / max(q_local, 1),
           "positions_equal_brute_force_scan_on_500_patterns": ok,
           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak,
                        "achieved": (q_local * per_pat + hits * per_hit) * args.steps / 1e9 / (ms / 1e3),
                        "bytes_per_pattern": per_pat, "bytes_per_hit": per_hit,
                        "note": "the 1 Gbp index image is 1.98 GB: every rank block and SA sample comes from HBM, not L2"},
           "index_bytes": int(fm.info.blob_bytes)}
    out["roofline"]["frac"] = out["roofline"]["achieved"] / peak
    fm.close()
    return out


def run_c1(ctx):
    """Config 1: toBWT -> toMTF -> toRLE and the inverse on a 64 KiB ACGT ByteString, host buffers
    in and out (a latency case: one small block), against the CPU restatement."""
    from tests.util import gen_acgt
    from text_compression_b200 import block
    from oracle import oracle as orc
    text = gen_acgt(0xC1, 65536)
    blk = block.compress_bwt_mtf_rle(text, ctx)
    assert block.decompress(blk, ctx) == text.tobytes()
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        blk = block.compress_bwt_mtf_rle(text, ctx)
    t_enc = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        block.decompress(blk, ctx)
    t_dec = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    bwt = orc.bwt_encode(text)
    idx, fin = orc.mtf_encode(bwt)
    cnt, sym = orc.rle_encode(idx.astype(np.int16))
    t_cpu_enc = time.perf_counter() - t0
    same = blk.counts.tolist() == cnt.tolist() and blk.syms.tolist() == sym.tolist() and blk.final_list.tolist() == fin.tolist()
    return {"workload": "C1: 64 KiB ACGT, BWT->MTF->RLE and inverse, host buffers, one call each",
            "encode_us": 1e6 * t_enc, "decode_us": 1e6 * t_dec, "cpu_port_encode_us": 1e6 * t_cpu_enc,
            "identical_to_cpu_port": bool(same), "runs": int(blk.R)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--fm", type=int, default=1, help="also run the FM-index count workload (config 3)")
    ap.add_argument("--fm-n", type=int, default=100_000_000)
    ap.add_argument("--fm-q", type=int, default=10_000_000)
    ap.add_argument("--fm-rate", type=int, default=32)
    ap.add_argument("--locate", type=int, default=1, help="also run the locate workload (config 4)")
    ap.add_argument("--c1", type=int, default=1, help="also time the 64 KiB round trip (config 1)")
    ap.add_argument("--decode", type=int, default=1, help="also time the inverse chain on 16 MiB blocks")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2: random-byte blocks (the headline); c5: ACGTN blocks (multi-block genome text)")
    ap.add_argument("--loc-n", type=int, default=1_000_000_000)
    ap.add_argument("--loc-q", type=int, default=1_000_000)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    global ALPHABET, WORKLOAD
    if args.workload == "c5":
        ALPHABET = "acgtn"
        WORKLOAD = "C5: BWT+MTF+RLE of 16 MiB synthetic ACGTN blocks, one block per GPU per step (blocks round-robin over GPUs)"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
