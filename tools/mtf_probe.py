"""Per-kernel device times of the MTF encoders on 16 MiB streams (tc_ctx_profile), both kernels.
usage: python tools/mtf_probe.py [n]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import gen_acgtn, gen_ascii, gen_bytes  # noqa: E402
from text_compression_b200 import _lib  # noqa: E402
from text_compression_b200._lib import ptr  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16 << 20
only = sys.argv[2] if len(sys.argv) > 2 else None      # one stream only
kernels = sys.argv[3] if len(sys.argv) > 3 else "01"   # "0": thread-per-chunk, "1": warp-per-chunk
rng = np.random.default_rng(5)
streams = {
    "bytes": gen_bytes(0xC2, n),
    "acgtn": gen_acgtn(0xC5, n),
    "ascii": gen_ascii(0xC2B, n),
    "skew96": (32 + np.minimum(rng.geometric(0.35, size=n) - 1, 94)).astype(np.uint8),
}
if only:
    streams = {only: streams[only]}
for v2 in kernels:
    os.environ["TC_B200_MTF_V2"] = v2
    ctx = _lib.Context(0)
    for name, b in streams.items():
        idx = np.empty(n, dtype=np.uint16)
        fin = np.empty(257, dtype=np.int16)
        sg = C.c_uint32(0)
        for rep in range(3):
            if rep == 2:
                ctx.profile(True)
            ctx.call("tc_mtf_encode_u8", ptr(b), n, n + 5, ptr(idx), ptr(fin), C.byref(sg))
        prof = ctx.profile_report()
        ctx.profile(False)
        tot = sum(v[1] for v in prof.values())
        print(f"v2={v2} {name} sigma={sg.value} total {tot * 1e3:.1f} us  " +
              "  ".join(f"{k.split('<')[0].replace('_kernel', '')}={v[1] * 1e3:.1f}" for k, v in prof.items()))
    ctx.close()
