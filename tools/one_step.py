"""One BWT+MTF+RLE step on a 16 MiB random-byte block (for ncu captures)."""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import gen_bytes, gen_acgtn
from text_compression_b200 import _lib, block

kind = sys.argv[1] if len(sys.argv) > 1 else "bytes"
n = int(sys.argv[2]) if len(sys.argv) > 2 else (16 << 20)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
text = gen_bytes(0xC2, n) if kind == "bytes" else gen_acgtn(0xC5, n)
ctx = _lib.Context(0)
for _ in range(reps):
    blk = block.compress_bwt_mtf_rle(text, ctx)
print("R", blk.R, "sigma", blk.sigma, "launches", ctx.launches)
