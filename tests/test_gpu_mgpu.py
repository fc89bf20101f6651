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
