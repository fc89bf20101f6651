"""Per-kernel CUDA-event times of one decompression: python tools/decode_times.py bytes|acgtn [n]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import gen_bytes, gen_acgtn
from text_compression_b200 import _lib, block
kind = sys.argv[1] if len(sys.argv) > 1 else "bytes"
n = int(sys.argv[2]) if len(sys.argv) > 2 else (16 << 20)
text = (gen_bytes if kind == "bytes" else gen_acgtn)(0xC5, n)
ctx = _lib.Context(0)
blk = block.compress_bwt_mtf_rle(text, ctx)
for _ in range(2):
    out = block.decompress(blk, ctx)
assert out == text.tobytes()
ctx.profile(True)
t0 = time.perf_counter()
out = block.decompress(blk, ctx)
wall = time.perf_counter() - t0
rep = ctx.profile_report()
tot = sum(v[1] for v in rep.values())
print(f"{kind} n={n} decode kernel_ms={tot:.3f} wall_ms={1e3*wall:.1f} -> {n/1e6/(tot/1e3):.0f} MB/s (kernel time only)")
for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {k[:46]:46s} x{v[0]:5.1f} {v[1]*1e3:9.1f} us")
