"""FM-index build + one count batch + one locate batch on synthetic ACGTN (for ncu captures and quick timing).
usage: python tools/fm_step.py [n] [q_count] [q_locate] [rate]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from tests.util import gen_acgtn, gen_reads  # noqa: E402
from text_compression_b200 import _lib, fmindex  # noqa: E402
from text_compression_b200._lib import ptr  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
qc = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
ql = int(sys.argv[3]) if len(sys.argv) > 3 else 200_000
rate = int(sys.argv[4]) if len(sys.argv) > 4 else 32
step = 100_000_000
text = np.concatenate([gen_acgtn(0xC3 + 1000 * i, min(step, n - o)) for i, o in enumerate(range(0, n, step))])
ctx = _lib.Context(0)
t0 = time.perf_counter()
fm = fmindex.FMIndex(text, "B", rate, ctx)
print(f"build {time.perf_counter() - t0:.3f} s, image {int(fm.info.blob_bytes) / 1e6:.0f} MB")
reads = gen_reads(0xC3 + 1, text, qc, 100)
pats = gen_reads(0xC3 + 2, text, ql, 32, mut_frac=0.0)
for rep in range(2):
    t0 = time.perf_counter()
    # flat arrays straight through the C ABI (no Python list of a million byte strings)
    off = np.arange(qc + 1, dtype=np.uint64) * 100
    cnt = np.empty(qc, dtype=np.int64)
    ctx.call("tc_fm_count", fm.h, ptr(np.ascontiguousarray(reads.reshape(-1))), ptr(off), qc, ptr(cnt))
    t1 = time.perf_counter()
    off = np.arange(ql + 1, dtype=np.uint64) * 32
    ho = np.empty(ql + 1, dtype=np.uint64)
    cap = 8 * ql + 1024
    pos = np.empty(cap, dtype=np.uint64)
    tot = C.c_uint64(0)
    ctx.call("tc_fm_locate", fm.h, ptr(np.ascontiguousarray(pats.reshape(-1))), ptr(off), ql, ptr(ho), ptr(pos), cap, C.byref(tot))
    t2 = time.perf_counter()
print(f"count {qc} reads {1e3 * (t1 - t0):.1f} ms (host in/out), found {(cnt >= 0).mean():.3f}; locate {ql} patterns {1e3 * (t2 - t1):.1f} ms, hits {tot.value}")
