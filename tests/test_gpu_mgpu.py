"""The multi-GPU entry points of the C ABI (tc_mgpu_*, tc_fm_replicate) against the oracle: one process, one
library call per batch, every visible GPU.  With a single GPU the same calls run with one device (and twice
the same device, which exercises the chunking and the peer copy path)."""
import numpy as np
import pytest

from tests.util import gen_acgtn, gen_bytes, gen_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _device_sets():
    from text_compression_b200 import multi
    n = multi.device_count()
    sets = [[0], [0, 0]]
    if n >= 2:
        sets.append(list(range(n)))
    return sets


def test_mgpu_blocks_encode_packed_vs_oracle(orc):
    from text_compression_b200 import block, multi
    texts = [gen_acgtn(0xC5 + b, 150_001 + 977 * b) for b in range(5)] + [gen_bytes(9, 70_000), gen_bytes(10, 0), gen_bytes(11, 1)]
    want = []
    for t in texts:
        if t.size == 0:
            want.append(None)
            continue
        idx, fin = orc.mtf_encode(orc.bwt_encode(t))
        cnt, sym = orc.rle_encode(idx.astype(np.int16))
        want.append((cnt, sym, fin))
    for devs in _device_sets():
        blobs = multi.compress_blocks_packed_multi(texts, True, devs)
        assert len(blobs) == len(texts)
        for b, blob in enumerate(blobs):
            u = block.unpack_block(blob)
            if want[b] is None:
                assert u.R == 0
                continue
            cnt, sym, fin = want[b]
            assert u.counts.tolist() == cnt.tolist() and u.syms.tolist() == sym.tolist(), (devs, b)
            assert u.final_list.tolist() == fin.tolist()


def test_fm_replicate_and_sharded_queries_vs_oracle(orc):
    from text_compression_b200 import fmindex, multi
    text = gen_acgtn(0xC3, 300_000)
    fm = fmindex.FMIndex(text, "B", 16)
    ofm = orc.FMIndex(text)
    reads = gen_reads(0xC3 + 1, text, 2_001, 20)
    pats = [r.tobytes() for r in reads] + [b"", b"ACGTX", b"A"]
    want_c = [ofm.count(p) for p in pats]
    for devs in _device_sets():
        reps = multi.FMReplicas(fm, devs)
        assert reps.count_many(pats).tolist() == want_c, devs
        ho, pos = reps.locate_many(pats[:300] + pats[-3:-1])
        sel = pats[:300] + pats[-3:-1]
        for i, p in enumerate(sel):
            assert pos[int(ho[i]):int(ho[i + 1])].tolist() == ofm.locate(p).tolist(), (devs, i)
        reps.close()
    fm.close()


def _nccl_worker(rank, world, port, out):
    """One process per GPU: index built on rank 0, image broadcast over NCCL (multi.build_replicated), each rank
    counts its contiguous chunk with tc_fm_count_dev, counts gathered in input order."""
    import ctypes as C
    import os
    import sys
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    from tests.util import gen_acgtn, gen_reads
    from text_compression_b200 import _lib, multi
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = _lib.Context(rank)
        text = gen_acgtn(0xC3, 400_000)
        d_text = torch.from_numpy(text).cuda()
        fm = multi.build_replicated(ctx, d_text, text.size, 16)
        assert int(fm.info.n) == text.size
        reads = gen_reads(0xC3 + 7, text, 5_003, 24)
        q = reads.shape[0]
        lo, hi = multi.query_slice(q, world, rank)
        mine = np.ascontiguousarray(reads[lo:hi])
        d_pats = torch.from_numpy(mine.reshape(-1)).cuda()
        d_off = (torch.arange(hi - lo + 1, dtype=torch.int64, device="cuda") * reads.shape[1])
        d_cnt = torch.empty(hi - lo, dtype=torch.int64, device="cuda")
        ctx.call("tc_fm_count_dev", fm.h, C.c_void_p(d_pats.data_ptr()), C.c_void_p(d_off.data_ptr()), hi - lo,
                 C.c_void_p(d_cnt.data_ptr()))
        torch.cuda.synchronize()
        full = multi.gather_in_order(d_cnt.cpu().numpy(), q)
        if rank == 0:
            out.put(full.tolist())
        fm.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_nccl_build_replicated_and_sharded_count_vs_oracle(orc):
    """World size 2 over NCCL (needs two GPUs: `gpurun --gpus 2`): the replicated index answers like the oracle."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mctx = mp.get_context("spawn")
    out = mctx.Queue()
    procs = [mctx.Process(target=_nccl_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    text = gen_acgtn(0xC3, 400_000)
    reads = gen_reads(0xC3 + 7, text, 5_003, 24)
    ofm = orc.FMIndex(text)
    want = [ofm.count(r.tobytes()) for r in reads]
    assert got == want
