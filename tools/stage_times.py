"""Per-kernel CUDA-event times of one composite step: python tools/stage_times.py bytes|acgtn [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import gen_bytes, gen_acgtn, gen_ascii
from text_compression_b200 import _lib, block
kind = sys.argv[1] if len(sys.argv) > 1 else "bytes"
n = int(sys.argv[2]) if len(sys.argv) > 2 else (16 << 20)
gen = {"bytes": gen_bytes, "acgtn": gen_acgtn, "ascii": gen_ascii}[kind]
text = gen(0xC5, n)
ctx = _lib.Context(0)
for _ in range(3):
    blk = block.compress_bwt_mtf_rle(text, ctx)
ctx.profile(True)
R = 3
for _ in range(R):
    blk = block.compress_bwt_mtf_rle(text, ctx)
rep = ctx.profile_report()
tot = sum(v[1] for v in rep.values()) / R
print(f"{kind} n={n} sigma={blk.sigma} R={blk.R} kernel_ms/step={tot:.3f} -> {n/1e6/(tot/1e3):.0f} MB/s (kernel time only)")
for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k[:46]:46s} x{v[0]/R:5.1f} {v[1]/R*1e3:9.1f} us")
