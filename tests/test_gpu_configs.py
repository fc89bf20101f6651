"""GPU parity at the BASELINE.json config sizes: the CUDA path (through the C ABI) against the CPU
oracle on the exact seeded blocks SURVEY.md 8d names -- every output array compared element for
element (primary, BWT, MTF indices, the FINAL list in order, run counts and run symbols), not through
a round trip.  The oracle needs ~11 s per 16 MiB block.  Run on the B200 box with `-m gpu`."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from tests.util import gen_acgtn, gen_ascii, gen_bytes, gen_reads, gen_words

pytestmark = pytest.mark.gpu

BLOCK = 16 << 20


@pytest.fixture(scope="module")
def ctx():
    from text_compression_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _same(name, got, want):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    if not np.array_equal(got, want):
        bad = np.nonzero(got != want)[0]
        raise AssertionError((name, "first mismatches at", bad[:5].tolist(), got[bad[:5]].tolist(),
                              want[bad[:5]].tolist(), "of", int(bad.size)))
    assert _digest(got.astype(np.int64)) == _digest(want.astype(np.int64))


# The oracle is single-threaded C and needs 6-25 s per config-size block; ctypes releases the GIL, so the oracle side
# of all config-size tests is computed by a few worker threads that start when the first of these tests runs, while
# the GPU side of the tests proceeds.  Every test still compares against exactly the oracle's arrays.
_JOBS = {}


def _oracle_block(text, both_chains=True):
    from oracle import oracle as orc
    o = {"text": text}
    o["bwt"] = orc.bwt_encode(text)            # createSuffixArray/saToBWT, seqToMTF, seqToRLE (oracle/tc_oracle.c)
    o["primary"] = int(np.nonzero(o["bwt"] < 0)[0][0])
    o["idx"], o["fin"] = orc.mtf_encode(o["bwt"])
    o["cnt"], o["sym"] = orc.rle_encode(o["idx"].astype(np.int16))
    if both_chains:
        o["r_cnt"], o["r_sym"] = orc.rle_encode(o["bwt"])
    return o


def _job_fm(n, q, m, seed):
    from oracle import oracle as orc
    step = 100_000_000
    text = np.concatenate([gen_acgtn(seed + 1000 * i, min(step, n - o)) for i, o in enumerate(range(0, n, step))])
    reads = gen_reads(seed + 1, text, q, m)
    # short patterns too: 12-mers have ~n / 4^12 occurrences each, which exercises multi-hit locate
    shorts = gen_reads(seed + 2, text, 64, 12, mut_frac=0.0)
    return text, [(p,) + tuple(orc.naive_search(text, p, want_pos=True)) for p in (reads, shorts)]


def _job_lsd32():
    from oracle import oracle as orc
    text = gen_acgtn(0xC5 + 7, 32 << 20)
    return (text,) + tuple(orc.bwt_encode(text, want_sa=True))


def _job(name):
    """Result of the named oracle job; the first call starts all of them."""
    if not _JOBS:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=5)
        _JOBS["c2"] = pool.submit(lambda: _oracle_block(gen_bytes(0xC2, BLOCK)))
        _JOBS["c2t"] = pool.submit(lambda: _oracle_block(gen_ascii(0xC2B, BLOCK)))
        _JOBS["c5"] = pool.submit(lambda: _oracle_block(gen_acgtn(0xC5, BLOCK)))
        _JOBS["lsd32"] = pool.submit(_job_lsd32)
        _JOBS["c3"] = pool.submit(lambda: _job_fm(100_000_000, 10_000, 100, 0xC3))
    return _JOBS[name].result()


def _check_block(ctx, orc, text, both_chains=True, oracle=None):
    """One block through every stage of both chains, each stage compared with the oracle."""
    from text_compression_b200 import block
    from text_compression_b200._lib import ptr
    n = text.size
    N = n + 1
    o = oracle if oracle is not None else _oracle_block(text, both_chains)
    o_bwt, o_primary, o_idx, o_fin, o_cnt, o_sym = o["bwt"], o["primary"], o["idx"], o["fin"], o["cnt"], o["sym"]
    # stage by stage through the C ABI
    bwt = np.empty(N, dtype=np.uint8)
    primary = C.c_uint64(0)
    ctx.call("tc_bwt_encode", ptr(text), n, ptr(bwt), C.byref(primary), None)
    assert primary.value == o_primary
    g = bwt.astype(np.int16)
    g[o_primary] = -1
    _same("bwt", g, o_bwt)
    idx = np.empty(N, dtype=np.uint16)
    fin = np.empty(257, dtype=np.int16)
    sigma = C.c_uint32(0)
    ctx.call("tc_mtf_encode_u8", ptr(bwt), N, o_primary, ptr(idx), ptr(fin), C.byref(sigma))
    assert sigma.value == o_fin.size
    _same("mtf final list (in order)", fin[: sigma.value], o_fin)
    _same("mtf indices", idx, o_idx)
    # the composite the bench times (device-resident chaining), records out
    blk = block.compress_bwt_mtf_rle(text, ctx)
    assert blk.primary == o_primary and blk.sigma == o_fin.size
    _same("composite final list", blk.final_list, o_fin)
    _same("composite run counts", blk.counts, o_cnt)
    _same("composite run symbols", blk.syms, o_sym)
    # run maximality, stated directly: no two neighbouring runs carry the same symbol
    assert not np.any(blk.syms[1:] == blk.syms[:-1])
    assert int(blk.counts.sum(dtype=np.uint64)) == N
    # packed container of the same block (the e2e output of the bench) unpacks to the same records
    blob = block.compress_blocks_packed([text], True, ctx)[0]
    u = block.unpack_block(blob)
    _same("container run counts", u.counts, o_cnt)
    _same("container run symbols", u.syms, o_sym)
    _same("container final list", u.final_list, o_fin)
    if both_chains:   # bytestringToBWTToRLEB: runs over the BWT symbols incl. Nothing
        r_cnt, r_sym = o["r_cnt"], o["r_sym"]
        blk2 = block.compress_bwt_rle(text, ctx)
        _same("bwt->rle counts", blk2.counts, r_cnt)
        _same("bwt->rle symbols", blk2.syms, r_sym)
    # and back
    assert block.decompress(blk, ctx) == text.tobytes()


def test_c2_random_byte_block_vs_oracle(ctx, orc):
    """BASELINE config 2, the block bench.py times: gen_bytes(0xC2, 16 MiB), sigma = 257."""
    o = _job("c2")
    _check_block(ctx, orc, o["text"], oracle=o)


def test_c2_text_block_vs_oracle(ctx, orc):
    """BASELINE config 2, Text variant: 16 MiB printable ASCII (seed 0xC2B), through the Text API too."""
    o = _job("c2t")
    text = o["text"]
    _check_block(ctx, orc, text, oracle=o)
    from text_compression_b200 import mtf as M
    s = text[: 1 << 20].tobytes().decode("ascii")
    m = M.textToBWTToMTFT(s, ctx)
    o_idx, o_fin = orc.mtf_encode(orc.bwt_encode(text[: 1 << 20]))
    _same("textToBWTToMTFT indices", m.indices, o_idx)
    _same("textToBWTToMTFT final list", m.final_list.codes, o_fin)


def test_c5_acgtn_block_vs_oracle(ctx, orc):
    """BASELINE config 5: one 16 MiB ACGTN block (seed 0xC5), sigma = 6 (the register-list MTF path)."""
    o = _job("c5")
    _check_block(ctx, orc, o["text"], oracle=o)


def test_lsd_path_32mi_vs_oracle(ctx, orc):
    """> 25 Mi symbols takes the LSD + prefix-doubling suffix sort -- the path that builds the C3 / C4
    indices.  32 Mi ACGTN: SA and BWT against the oracle."""
    from text_compression_b200._lib import ptr
    text, o_bwt, o_sa = _job("lsd32")
    n = text.size
    bwt = np.empty(n + 1, dtype=np.uint8)
    sa = np.empty(n + 1, dtype=np.uint32)
    primary = C.c_uint64(0)
    ctx.call("tc_bwt_encode", ptr(text), n, ptr(bwt), C.byref(primary), ptr(sa))
    _same("suffix array", sa, o_sa)
    g = bwt.astype(np.int16)
    g[primary.value] = -1
    _same("bwt", g, o_bwt)


def _check_fm_full(ctx, text, cases, rate):
    """Index over synthetic ACGTN; sampled reads (10 % with a substitution) against an independent occurrence scan of
    the text (oracle.naive_search, done by _job_fm): counts, the located positions as sets, and the SA-rank order of
    every multi-hit pattern by comparing the located suffixes."""
    from text_compression_b200 import fmindex
    n = text.size
    fm = fmindex.FMIndex(text, "B", rate, ctx)
    assert int(fm.info.N) == n + 1 and int(fm.info.sigma) == 6   # $ACGNT
    for pats, cnt, ho, pos in cases:
        got = fm.count_many([p.tobytes() for p in pats])
        want = np.where(cnt > 0, cnt, -1)   # countFMIndex: Nothing when there is no occurrence
        _same("count", got, want)
        gho, gpos = fm.locate_many([p.tobytes() for p in pats])
        _same("hit offsets", gho, ho)
        mm = pats.shape[1]
        for i in range(pats.shape[0]):
            a = gpos[int(gho[i]):int(gho[i + 1])].astype(np.int64)
            b = pos[int(ho[i]):int(ho[i + 1])].astype(np.int64)
            assert np.array_equal(np.sort(a), b), (i, a[:8], b[:8])
            # SA-rank order (src/Data/FMIndex.hs:473-474): located suffixes ascend lexicographically
            for x, y in zip(a[:-1].tolist()[:64], a[1:].tolist()[:64]):
                sx, sy = text[x - 1: x - 1 + mm + 256].tobytes(), text[y - 1: y - 1 + mm + 256].tobytes()
                assert sx < sy, (i, x, y)
    fm.close()


def test_c3_full_size_count_and_locate(ctx, orc):
    """BASELINE config 3 at full size: 100 Mbp reference, 10,000 sampled 100-bp reads, SA rate 32."""
    text, cases = _job("c3")
    _check_fm_full(ctx, text, cases, 32)


def test_c4_locate_lsd_built_index(ctx, orc):
    """BASELINE config 4: sampled SA (rate 32), 2,000 32-bp patterns, on an index that is built by the LSD + doubling
    path and lives in HBM.  128 Mbp by default so the GPU suite stays within a few minutes; TC_TEST_C4_FULL=1 runs the
    full 1 Gbp reference (40-80 s: generating and scanning 1 GB on the host).  bench.py checks 500 located patterns
    against a brute-force scan at the full 1 Gbp size on every run."""
    import os
    n = 1_000_000_000 if os.environ.get("TC_TEST_C4_FULL") == "1" else 128_000_000
    text, cases = _job_fm(n, 2_000, 32, 0xC4)
    _check_fm_full(ctx, text, cases, 32)


def test_mtf_kernels_agree(ctx, orc):
    """The thread-per-chunk MTF encoder (default) and the warp-per-chunk one (TC_B200_MTF_V2=1) against the
    oracle on skewed, run-heavy and uniform streams over alphabets of 9..257 symbols."""
    import os
    from text_compression_b200 import _lib
    from text_compression_b200._lib import ptr
    os.environ["TC_B200_MTF_V2"] = "1"
    try:
        ctx2 = _lib.Context(0)
    finally:
        del os.environ["TC_B200_MTF_V2"]
    rng = np.random.default_rng(77)
    for n in (1, 31, 32, 33, 735, 736, 737, 4097, 70001, 1_000_003, 2_000_000):
        for sigma, mode in ((9, "uniform"), (40, "skew"), (96, "runs"), (200, "skew"), (256, "uniform"), (256, "runs")):
            alpha = rng.choice(256, size=sigma, replace=False).astype(np.uint8)
            if mode == "uniform":
                b = alpha[rng.integers(0, sigma, size=n)]
            elif mode == "skew":   # what the BWT of real text looks like: a few symbols dominate locally
                b = alpha[np.minimum(rng.geometric(0.3, size=n) - 1, sigma - 1)]
            else:
                reps = rng.integers(1, 40, size=n // 8 + 1)
                b = np.repeat(alpha[rng.integers(0, sigma, size=reps.size)], reps)[:n]
            b = np.ascontiguousarray(b)
            for primary in sorted({0, n // 3, n + 9} if n > 100_000 else {0, n // 3, n - 1, n + 9}):
                x = b.astype(np.int16)
                if primary < n:
                    x[primary] = -1
                o_idx, o_fin = orc.mtf_encode(x)
                for c in (ctx, ctx2):
                    idx = np.empty(n, dtype=np.uint16)
                    fin = np.empty(257, dtype=np.int16)
                    sg = C.c_uint32(0)
                    c.call("tc_mtf_encode_u8", ptr(b), n, primary, ptr(idx), ptr(fin), C.byref(sg))
                    _same(f"indices n={n} sigma={sigma} {mode} primary={primary}", idx, o_idx)
                    _same("final list", fin[: sg.value], o_fin)
    ctx2.close()


def test_raw_byte_keys_vs_oracle(ctx, orc):
    """Equiprobable bytes take the leading-bytes uniform key (sufsort.cu uk_keys_raw_kernel): 4 MiB + 3 of random bytes
    (odd length: the last suffixes read past the end), against the oracle and against the arithmetic-code keys
    (TC_B200_NO_RAWKEY=1); and a text whose byte histogram is flat but which repeats a 4 KiB block (every key
    ties: the general path must take over)."""
    import os
    from text_compression_b200 import _lib
    from text_compression_b200._lib import ptr
    os.environ["TC_B200_NO_RAWKEY"] = "1"
    try:
        ctx2 = _lib.Context(0)
    finally:
        del os.environ["TC_B200_NO_RAWKEY"]
    rng = np.random.default_rng(5)
    blk = np.repeat(np.arange(256, dtype=np.uint8), 16)
    rng.shuffle(blk)
    texts = [gen_bytes(77, (4 << 20) + 3), np.tile(blk, 300)[: 300 * 4096 - 5]]
    for t in texts:
        o_bwt, o_sa = orc.bwt_encode(t, want_sa=True)
        for c in (ctx, ctx2):
            bwt = np.empty(t.size + 1, dtype=np.uint8)
            sa = np.empty(t.size + 1, dtype=np.uint32)
            primary = C.c_uint64(0)
            c.call("tc_bwt_encode", ptr(t), t.size, ptr(bwt), C.byref(primary), ptr(sa))
            _same("suffix array", sa, o_sa)
            g = bwt.astype(np.int16)
            g[primary.value] = -1
            _same("bwt", g, o_bwt)
    ctx2.close()


def test_correlated_text_vs_oracle(ctx, orc):
    """Correlated text leaves the uniform-key suffix sort (oversized buckets, deep ties) for the LSD + prefix-doubling
    path with many rounds; MTF indices are mostly 0 and runs are long.  Synthetic word text with verbatim repeats, and
    1 MiB / 4 MiB slices of real Python sources, every stage against the oracle."""
    _check_block(ctx, orc, gen_words(0x9C, 1 << 20))
    _check_block(ctx, orc, gen_words(0x9D, (3 << 20) + 12345, dup_every=1 << 16, dup_len=20000), both_chains=False)
    from tests.util import python_corpus
    corpus = python_corpus(6 << 20)
    if corpus.size < (6 << 20):
        pytest.skip("site-packages holds less than 6 MiB of Python sources")
    _check_block(ctx, orc, corpus[: 1 << 20])
    _check_block(ctx, orc, corpus[5 << 20:], both_chains=False)
    _check_block(ctx, orc, corpus[1 << 20: 5 << 20], both_chains=False)


def test_mtf_decode_kernels_agree(ctx, orc):
    """seqFromMTF for alphabets of 9..257 symbols: the select-based kernels (default) and the list-shifting ones
    (TC_B200_MTFD_V1=1) against the oracle -- uniform indices (the worst case of list shifting), mostly-zero indices,
    always-the-back, with and without Nothing in the alphabet, sizes around the chunk and tile boundaries; an index
    beyond the list is the reference's `index out of bounds` in both."""
    import os
    from text_compression_b200 import _lib
    from text_compression_b200._lib import TC_E_INDEX, ptr
    os.environ["TC_B200_MTFD_V1"] = "1"
    try:
        ctx1 = _lib.Context(0)
    finally:
        del os.environ["TC_B200_MTFD_V1"]
    rng = np.random.default_rng(1234)
    for n in (1, 7, 8, 9, 31, 223, 224, 225, 4097, 35_841, 160 * 224 + 5, 1_000_003, 2_500_001):
        for sigma, mode, nothing in ((9, "uniform", True), (33, "zeros", False), (64, "back", True), (200, "uniform", False),
                                     (256, "small", True), (257, "uniform", True), (257, "back", True)):
            syms = np.sort(rng.choice(256, size=sigma - (1 if nothing else 0), replace=False)).astype(np.int16)
            fin = np.concatenate([np.array([-1], np.int16), syms]) if nothing else syms
            fin = fin[rng.permutation(fin.size)]          # any order: only the set matters (the decoder re-sorts it)
            if mode == "uniform":
                idx = rng.integers(0, sigma, size=n)
            elif mode == "zeros":
                idx = np.where(rng.random(n) < 0.9, 0, rng.integers(0, sigma, size=n))
            elif mode == "small":
                idx = np.minimum(rng.geometric(0.4, size=n) - 1, sigma - 1)
            else:
                idx = np.full(n, sigma - 1)
            idx = np.ascontiguousarray(idx, dtype=np.uint16)
            want = orc.mtf_decode(idx, fin)
            for c in (ctx, ctx1):
                out = np.empty(n, dtype=np.int16)
                c.call("tc_mtf_decode", ptr(idx), n, ptr(fin), fin.size, ptr(out))
                _same(f"mtf decode n={n} sigma={sigma} {mode}", out, want)
    idx = np.zeros(70_000, dtype=np.uint16)
    idx[54_321] = 40
    fin = np.arange(40, dtype=np.int16)
    for c in (ctx, ctx1):
        out = np.empty(idx.size, dtype=np.int16)
        assert c.L.tc_mtf_decode(c.h, ptr(idx), idx.size, ptr(fin), fin.size, ptr(out)) == TC_E_INDEX
    ctx1.close()


def test_blocks_decode_packed_vs_texts(ctx, orc):
    """tc_blocks_decode_packed (multi-block decompression, lanes) gives back every block of a ragged batch, both
    chains, an empty block included; equals the single-container call; a stream round trip goes through it; a
    malformed container stops the batch with an error."""
    from text_compression_b200 import block, stream
    from text_compression_b200._lib import TcError
    texts = [gen_bytes(3, 1_000_003), gen_acgtn(4, 2_500_000), gen_words(5, 700_001), gen_bytes(6, 0), gen_ascii(7, 1),
             gen_acgtn(8, 4097), gen_bytes(9, 3 << 20), gen_acgtn(10, 65_536), gen_bytes(11, 4095)]
    from text_compression_b200._lib import FromJustError
    for with_mtf in (True, False):
        blobs = block.compress_blocks_packed(texts, with_mtf, ctx)
        keep = list(range(len(texts)))
        if not with_mtf:
            # reference quirk: seqToRLE writes a trailing Nothing twice, so a BWT that ENDS with
            # its Nothing does not survive BWT -> RLE -> BWT in the reference either: fromBWT meets a second Nothing
            quirk = [b for b in keep if texts[b].size and int(np.nonzero(orc.bwt_encode(texts[b]) < 0)[0][0]) == texts[b].size]
            assert 4 in quirk                                   # the one-symbol text: BWT = [x, Nothing]
            for b in quirk:
                with pytest.raises(FromJustError):
                    block.decompress_packed(blobs[b], ctx)
                with pytest.raises(FromJustError):              # and it stops a batch
                    block.decompress_blocks_packed([blobs[0], blobs[b], blobs[1]], ctx)
            keep = [b for b in keep if b not in quirk]
        got = block.decompress_blocks_packed([blobs[b] for b in keep], ctx)
        assert len(got) == len(keep)
        for g, b in zip(got, keep):
            assert g == texts[b].tobytes(), (with_mtf, b)
            assert block.decompress_packed(blobs[b], ctx) == g
    # the pointer-jumping rounds as separate launches (devices without cooperative launch) give the same text
    import os
    from text_compression_b200 import _lib
    os.environ["TC_B200_NO_COOP"] = "1"
    try:
        ctx_nc = _lib.Context(0)
    finally:
        del os.environ["TC_B200_NO_COOP"]
    blobs = block.compress_blocks_packed(texts[:3], True, ctx)
    for b in range(3):
        assert block.decompress_packed(blobs[b], ctx_nc) == texts[b].tobytes()
    ctx_nc.close()
    data = np.concatenate(texts).tobytes()
    assert stream.decompress_stream(stream.compress_stream(data, 1 << 20, ctx=ctx), ctx) == data
    bad = [np.array(b, copy=True) for b in block.compress_blocks_packed(texts[:4], True, ctx)]
    bad[1][0] ^= 0xff          # magic
    with pytest.raises((TcError, ValueError)):
        block.decompress_blocks_packed(bad, ctx)
    assert block.decompress_blocks_packed([], ctx) == []
