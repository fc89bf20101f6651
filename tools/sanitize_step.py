"""Small end-to-end pass for compute-sanitizer: both alphabets, both sort paths, encode + decode, FM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import gen_bytes, gen_acgtn, gen_ascii
from text_compression_b200 import _lib, block, fmindex
ctx = _lib.Context(0)
for gen, n in ((gen_bytes, 300_001), (gen_acgtn, 200_003), (gen_ascii, 70_000), (gen_acgtn, 5000)):
    t = gen(7, n)
    for f in (block.compress_bwt_mtf_rle, block.compress_bwt_rle):
        b = f(t, ctx)
        assert block.decompress(b, ctx) == t.tobytes()
rep = np.tile(gen_acgtn(3, 3000), 8)          # repeats: ties, deep compare, doubling rounds
b = block.compress_bwt_mtf_rle(rep, ctx)
assert block.decompress(b, ctx) == rep.tobytes()
bl = block.compress_blocks([gen_bytes(1, 50000), gen_acgtn(2, 80000), np.empty(0, np.uint8)], True, ctx)
texts = [gen_bytes(1, 50000), gen_acgtn(2, 80000), np.empty(0, np.uint8), np.frombuffer(b"a" * 9000 + b"bc" * 500, np.uint8),
         np.tile(np.arange(256, dtype=np.uint8), 30), gen_ascii(4, 4099)]
for with_mtf in (True, False):
    for blob, t in zip(block.compress_blocks_packed(texts, with_mtf, ctx), texts):
        u = block.unpack_block(blob)
        assert int(u.counts.sum()) == (t.size + 1 if t.size else 0) or not with_mtf
        if with_mtf:
            assert block.decompress_packed(blob, ctx) == t.tobytes()
fm = fmindex.FMIndex(gen_acgtn(0xC3, 100000), "B", 32, ctx)
pats = [gen_acgtn(0xC3, 100000)[o:o + 20].tobytes() for o in range(0, 50000, 501)]
c = fm.count_many(pats)
ho, pos = fm.locate_many(pats)
assert (c > 0).all()
print("sanitize step ok")
