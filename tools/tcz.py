"""File compressor over the block container: python tools/tcz.py c|d IN OUT [block MiB]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_compression_b200 import stream

mode, src, dst = sys.argv[1:4]
mib = int(sys.argv[4]) if len(sys.argv) > 4 else 16
data = open(src, "rb").read()
out = stream.compress_stream(data, mib << 20) if mode == "c" else stream.decompress_stream(data)
open(dst, "wb").write(out)
print(f"{len(data)} -> {len(out)} bytes")
