"""Per-call wall time of tc_blocks_encode_packed (C2 blocks), for TC_B200_LANES=1/2."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import gen_bytes, gen_acgtn
from text_compression_b200 import _lib
from text_compression_b200._lib import BlockInfo
n = 16 << 20
gen = gen_acgtn if len(sys.argv) > 1 and sys.argv[1] == "acgtn" else gen_bytes
ctx = _lib.Context(0)
NH = 4
pcap = int(ctx.L.tc_packed_bound(n))
h_in = [_lib.pinned_empty(n, np.uint8) for _ in range(NH)]
h_out = [_lib.pinned_empty(pcap, np.uint8) for _ in range(NH)]
for j in range(NH):
    h_in[j][:] = gen(0xC2 + j, n)
def batch(nb):
    tp = (C.c_void_p * nb)(*[h_in[b % NH].ctypes.data for b in range(nb)])
    ns = (C.c_uint64 * nb)(*([n] * nb))
    infos = (BlockInfo * nb)()
    op = (C.c_void_p * nb)(*[h_out[b % NH].ctypes.data for b in range(nb)])
    caps = (C.c_uint64 * nb)(*([pcap] * nb))
    nbytes = (C.c_uint64 * nb)()
    t0 = time.perf_counter()
    ctx.call("tc_blocks_encode_packed", nb, tp, ns, 1, op, caps, nbytes, infos)
    return time.perf_counter() - t0
for nb in (3, 10, 10, 10, 20, 40):
    t = batch(nb)
    print(f"lanes={os.environ.get('TC_B200_LANES','2')} nb={nb}: {1e3*t:.2f} ms = {1e3*t/nb:.3f} ms/block = {nb*n/1e6/t:.0f} MB/s", flush=True)
ctx.profile(True)      # profiling keeps the batch on one lane: per-kernel times of the packed path
batch(3)
rep = ctx.profile_report()
ctx.profile(False)
for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    if k.startswith("rle") or k.startswith("(rle"):
        print(f"  {k[:46]:46s} x{v[0]/3:5.1f} {v[1]/3*1e3:9.1f} us")
print(f"  all kernels: {sum(v[1] for v in rep.values())/3:.3f} ms/block")
