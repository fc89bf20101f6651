"""Throughput on non-synthetic text: a corpus of Python sources from site-packages (correlated,
repetitive; tests/util.python_corpus).  python tools/real_text.py [MiB]"""
import sys, os, time, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from text_compression_b200 import _lib, block
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 16
from tests.util import python_corpus
text = python_corpus(mib << 20)
print("corpus bytes", text.size, "distinct", len(set(text.tolist()[:1000000])))
ctx = _lib.Context(0)
for _ in range(2):
    blk = block.compress_bwt_mtf_rle(text, ctx)
ctx.profile(True)
t0 = time.perf_counter()
blk = block.compress_bwt_mtf_rle(text, ctx)
wall = time.perf_counter() - t0
rep = ctx.profile_report()
tot = sum(v[1] for v in rep.values())
print(f"n={text.size} R={blk.R} ({blk.R/text.size:.3f} runs/byte) kernel_ms={tot:.2f} wall_ms={1e3*wall:.1f} -> {text.size/1e6/(tot/1e3):.0f} MB/s (kernel time)")
for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  {k[:46]:46s} x{v[0]:5.1f} {v[1]*1e3:9.1f} us")
assert block.decompress(blk, ctx) == text.tobytes()
print("round trip ok")
