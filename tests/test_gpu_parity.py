"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs, bit-exact.  Run on the B200 box with `-m gpu`."""
import numpy as np
import pytest

from tests.util import gen_acgt, gen_acgtn, gen_ascii, gen_bytes

pytestmark = pytest.mark.gpu


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


def _maybe_stream(rng, n, alpha, p_nothing):
    x = rng.choice(np.asarray(alpha, dtype=np.int16), size=n).astype(np.int16)
    if p_nothing > 0 and n:
        x[rng.random(n) < p_nothing] = -1
    return x


SIZES = [0, 1, 2, 3, 5, 17, 255, 256, 257, 2047, 2048, 2049, 4095, 4096, 4097, 8193, 70001, 300007]


# ---------------------------------------------------------------- RLE
@pytest.mark.parametrize("n", SIZES)
def test_rle_encode_i16(ctx, orc, n):
    from text_compression_b200.rle import seqToRLE, seqFromRLE
    from text_compression_b200.seq import MaybeSeq
    rng = np.random.default_rng(n + 1)
    for alpha, pn in (([65, 67], 0.0), ([65, 67, 71, 84], 0.02), ([65], 0.3), (list(range(256)), 0.01), ([7], 1.0)):
        x = _maybe_stream(rng, n, alpha, pn)
        got = seqToRLE(MaybeSeq(x, "B"), ctx)
        cnt, sym = orc.rle_encode(x)
        assert got.counts.tolist() == cnt.tolist() and got.syms.tolist() == sym.tolist(), (n, alpha[:4], pn)
        dec = seqFromRLE(got, ctx)
        assert dec.codes.tolist() == orc.rle_decode(cnt, sym).tolist()


def test_rle_quirks(ctx, orc):
    from text_compression_b200.rle import seqToRLE
    from text_compression_b200.seq import MaybeSeq
    cases = [[-1], [-1, -1], [97], [97, -1], [-1, 97, 97, 98], [97, 97, -1, -1, 98], [97, -1, 98, -1],
             [97, 97, 97, 97], [-1, -1, -1, 97, -1, -1]]
    for c in cases:
        x = np.array(c, dtype=np.int16)
        got = seqToRLE(MaybeSeq(x, "B"), ctx)
        cnt, sym = orc.rle_encode(x)
        assert list(zip(got.counts.tolist(), got.syms.tolist())) == list(zip(cnt.tolist(), sym.tolist())), c


@pytest.mark.parametrize("n", [1, 16, 4096, 4097, 100003])
def test_rle_u8_u16(ctx, orc, n):
    import ctypes as C
    from text_compression_b200._lib import ptr
    rng = np.random.default_rng(n)
    for alpha in ([65, 67, 71, 84], list(range(256))):
        b = rng.choice(np.asarray(alpha, dtype=np.uint8), size=n).astype(np.uint8)
        for primary in sorted({0, n // 2, n - 1, n + 5}):
            x = b.astype(np.int16)
            if primary < n:
                x[primary] = -1
            cnt, sym = orc.rle_encode(x)
            cap = n + 3
            oc, os_ = np.empty(cap, np.uint32), np.empty(cap, np.int16)
            R = C.c_uint64(0)
            ctx.call("tc_rle_encode_u8", ptr(b), n, primary, ptr(oc), ptr(os_), cap, C.byref(R))
            assert oc[:R.value].tolist() == cnt.tolist() and os_[:R.value].tolist() == sym.tolist(), (n, primary)
        idx = rng.integers(0, 257, size=n).astype(np.uint16)
        idx[rng.random(n) < 0.5] = 0
        cnt, sym = orc.rle_encode(idx.astype(np.int16))
        oc, os_ = np.empty(n + 1, np.uint32), np.empty(n + 1, np.int16)
        R = C.c_uint64(0)
        ctx.call("tc_rle_encode_u16", ptr(idx), n, ptr(oc), ptr(os_), n + 1, C.byref(R))
        assert oc[:R.value].tolist() == cnt.tolist() and os_[:R.value].tolist() == sym.tolist()


def test_rle_cap(ctx):
    import ctypes as C
    from text_compression_b200._lib import ptr, TC_E_CAP
    x = np.arange(100, dtype=np.int16) % 7
    oc, os_ = np.empty(10, np.uint32), np.empty(10, np.int16)
    R = C.c_uint64(0)
    rc = ctx.call("tc_rle_encode", ptr(x), 100, ptr(oc), ptr(os_), 10, C.byref(R), allow=(TC_E_CAP,))
    assert rc == TC_E_CAP and R.value == 100


def test_rle_decode_long_runs(ctx, orc):
    from text_compression_b200.rle import seqFromRLE
    from text_compression_b200.seq import RLE
    cnt = np.array([100000, 1, 0, 3, 50000, 7, 1], dtype=np.uint32)
    sym = np.array([65, -1, 66, 67, 68, -1, 69], dtype=np.int16)
    got = seqFromRLE(RLE(cnt, sym, "B"), ctx)
    assert got.codes.tolist() == orc.rle_decode(cnt.astype(np.int64), sym).tolist()


# ---------------------------------------------------------------- MTF
@pytest.mark.parametrize("n", SIZES)
def test_mtf_roundtrip_and_parity(ctx, orc, n):
    from text_compression_b200.mtf import seqToMTF, seqFromMTF
    from text_compression_b200.seq import MaybeSeq
    rng = np.random.default_rng(n + 7)
    for alpha, pn in (([65, 67, 71, 84], 0.001), (list(range(256)), 0.001), ([120], 0.0), (list(range(32, 127)), 0.0),
                      ([3, 200], 0.4)):
        x = _maybe_stream(rng, n, alpha, pn)
        got = seqToMTF(MaybeSeq(x, "B"), ctx)
        idx, fin = orc.mtf_encode(x)
        assert got.final_list.codes.tolist() == fin.tolist(), (n, len(alpha))
        bad = np.nonzero(got.indices != idx)[0]
        assert bad.size == 0, (n, len(alpha), bad[:5], got.indices[bad[:5]], idx[bad[:5]])
        dec = seqFromMTF(got, ctx)
        assert dec.codes.tolist() == x.tolist(), (n, len(alpha))


def test_mtf_skewed_large(ctx, orc):
    from text_compression_b200.mtf import seqToMTF, seqFromMTF
    from text_compression_b200.seq import MaybeSeq
    rng = np.random.default_rng(5)
    n = 1_000_003
    x = rng.choice(np.arange(256, dtype=np.int16), size=n, p=np.r_[[0.9], np.full(255, 0.1 / 255)]).astype(np.int16)
    x[123456] = -1
    got = seqToMTF(MaybeSeq(x, "B"), ctx)
    idx, fin = orc.mtf_encode(x)
    assert got.final_list.codes.tolist() == fin.tolist()
    assert np.array_equal(got.indices, idx)
    assert np.array_equal(seqFromMTF(got, ctx).codes, x)


def test_mtf_decode_bad_index(ctx):
    from text_compression_b200 import SeqIndexError
    from text_compression_b200.mtf import seqFromMTF
    from text_compression_b200.seq import MTF, MaybeSeq
    m = MTF(np.array([0, 1, 5], dtype=np.int64), MaybeSeq(np.array([97, 98, -1], dtype=np.int16), "B"))
    with pytest.raises(SeqIndexError):
        seqFromMTF(m, ctx)


# ---------------------------------------------------------------- BWT
def _bwt_cases():
    rng = np.random.default_rng(11)
    yield "empty", np.empty(0, np.uint8)
    for n in (1, 2, 3, 4, 31, 1000, 4097, 65536):
        yield f"acgt{n}", gen_acgt(0xC1, n)
        yield f"bytes{n}", gen_bytes(0xC2, n)
    yield "a", np.frombuffer(b"a", np.uint8)
    yield "ba", np.frombuffer(b"ba", np.uint8)
    yield "aaaa", np.full(5000, 97, np.uint8)
    yield "abab", np.tile(np.frombuffer(b"ab", np.uint8), 3000)
    yield "period7", np.tile(gen_acgt(3, 7), 2000)
    yield "fib", _fib(16)
    yield "twobytes", rng.integers(0, 2, 20000).astype(np.uint8) * 255
    yield "zeros_in_text", rng.integers(0, 3, 5000).astype(np.uint8)
    yield "ascii", gen_ascii(0xC2B, 50000)
    yield "acgtn", gen_acgtn(0xC3, 200000)


def _fib(k):
    a, b = b"a", b"ab"
    for _ in range(k):
        a, b = b, b + a
    return np.frombuffer(b, np.uint8)


@pytest.mark.parametrize("name,text", list(_bwt_cases()), ids=[c[0] for c in _bwt_cases()])
def test_bwt_encode_decode(ctx, orc, name, text):
    from text_compression_b200.bwt import bwt_u8, toBWT, fromBWT, createSuffixArray
    bwt, primary, sa = bwt_u8(text, want_sa=True, ctx=ctx)
    want_bwt, want_sa = orc.bwt_encode(text, want_sa=True)
    got = toBWT(text, ctx).seq.codes
    assert got.size == want_bwt.size
    if text.size:
        assert sa.tolist() == want_sa.tolist(), name
        assert got.tolist() == want_bwt.tolist(), name
        assert primary == int(np.nonzero(want_bwt < 0)[0][0])
    from text_compression_b200.seq import BWT, MaybeSeq
    back = fromBWT(BWT(MaybeSeq(got, "W")), ctx)
    assert bytes(back) == text.tobytes(), name


def test_bwt_decode_malformed(ctx, orc):
    from text_compression_b200 import FromJustError
    from text_compression_b200.bwt import fromBWT
    from text_compression_b200.seq import BWT, MaybeSeq
    rng = np.random.default_rng(3)
    # no Nothing -> empty (src/Data/BWT/Internal.hs:174-175)
    assert fromBWT(BWT(MaybeSeq(np.array([97, 98, 99], np.int16), "W")), ctx) == []
    n_err = n_ok = 0
    for trial in range(60):
        n = int(rng.integers(1, 400))
        x = rng.integers(97, 100, size=n).astype(np.int16)
        k = 1 if trial % 2 == 0 else int(rng.integers(1, 4))
        x[rng.choice(n, size=min(k, n), replace=False)] = -1
        try:
            want = orc.bwt_decode(x).tolist()
            werr = False
        except orc.OracleError as e:
            assert e.rc == orc.ORC_E_FROMJUST
            werr = True
        if werr:
            with pytest.raises(FromJustError):
                fromBWT(BWT(MaybeSeq(x, "W")), ctx)
            n_err += 1
        else:
            assert fromBWT(BWT(MaybeSeq(x, "W")), ctx) == want, x.tolist()
            n_ok += 1
    assert n_ok > 0


# ---------------------------------------------------------------- suffix-sort paths
def _sort_path_cases():
    rng = np.random.default_rng(23)
    a = gen_acgt(5, 6000)
    yield "repeat_block", np.concatenate([a, gen_acgt(6, 3000), a, a[:2500]])          # ties far beyond the key
    yield "low_entropy", np.where(rng.random(40000) < 0.97, 65, rng.integers(66, 70, 40000)).astype(np.uint8)
    yield "two_symbols", rng.integers(0, 2, 30000).astype(np.uint8) + 97
    yield "runs", np.repeat(gen_acgt(7, 3000), rng.integers(1, 30, 3000))
    yield "bytes_8191", gen_bytes(1, 8191)
    yield "bytes_4096", gen_bytes(2, 4096)
    yield "acgtn_300k", gen_acgtn(3, 300_000)
    yield "ascii_131k", gen_ascii(4, 131_072)
    yield "zero_bytes", np.where(rng.random(50000) < 0.5, 0, rng.integers(0, 4, 50000)).astype(np.uint8)


@pytest.mark.parametrize("name,text", list(_sort_path_cases()), ids=[c[0] for c in _sort_path_cases()])
def test_suffix_sort_msd_and_fallback(ctx, orc, name, text, monkeypatch):
    """The same inputs through the MSD path (default) and the LSD + prefix-doubling path
    (TC_B200_NO_MSD=1, read at context creation): both must equal the oracle's suffix array."""
    from text_compression_b200 import _lib
    from text_compression_b200.bwt import bwt_u8
    want_bwt, want_sa = orc.bwt_encode(text, want_sa=True)
    bwt, primary, sa = bwt_u8(text, want_sa=True, ctx=ctx)
    assert np.array_equal(sa, want_sa), name
    assert primary == int(np.nonzero(want_bwt < 0)[0][0])
    monkeypatch.setenv("TC_B200_NO_MSD", "1")
    c2 = _lib.Context(0)
    try:
        bwt2, primary2, sa2 = bwt_u8(text, want_sa=True, ctx=c2)
    finally:
        c2.close()
    assert np.array_equal(sa2, want_sa), name
    assert primary2 == primary and np.array_equal(np.delete(bwt2, primary2), np.delete(bwt, primary))


def test_suffix_sort_fuzz_msd_vs_lsd(ctx, orc, monkeypatch):
    """Random sizes, alphabet sizes (every code width b = 1..9) and symbol skews: the MSD path
    must give the same suffix array as the LSD path, and both the oracle's on the smaller ones."""
    from text_compression_b200 import _lib
    from text_compression_b200.bwt import bwt_u8
    rng = np.random.default_rng(2024)
    monkeypatch.setenv("TC_B200_NO_MSD", "1")
    c2 = _lib.Context(0)
    try:
        sigmas = [1, 2, 3, 4, 7, 8, 15, 16, 31, 33, 64, 100, 128, 200, 255, 256]
        for trial, sigma in enumerate(sigmas * 2):
            n = int(rng.integers(4096, 3_000_000 if trial % 4 == 0 else 300_000))
            alpha = rng.choice(256, size=sigma, replace=False).astype(np.uint8)
            if trial % 3 == 0:
                p = rng.random(sigma) ** 4 + 1e-6     # heavily skewed frequencies
                p /= p.sum()
            else:
                p = np.full(sigma, 1.0 / sigma)
            text = alpha[rng.choice(sigma, size=n, p=p)]
            bwt, primary, sa = bwt_u8(text, want_sa=True, ctx=ctx)
            bwt2, primary2, sa2 = bwt_u8(text, want_sa=True, ctx=c2)
            assert primary == primary2 and np.array_equal(sa, sa2), (trial, sigma, n)
            assert np.array_equal(np.delete(bwt, primary), np.delete(bwt2, primary2)), (trial, sigma, n)
            if n <= 60_000 and sigma > 1:
                _, want_sa = orc.bwt_encode(text, want_sa=True)
                assert np.array_equal(sa, want_sa), (trial, sigma, n)
    finally:
        c2.close()


def test_block_over_16mib_roundtrip(ctx):
    """n > 2^24: the record no longer has room for the preceding byte, so the last sort level
    gathers the BWT symbols from the text instead."""
    from text_compression_b200 import block
    text = gen_acgtn(0xC5, (16 << 20) + 12345)
    blk = block.compress_bwt_mtf_rle(text, ctx)
    assert blk.N == text.size + 1 and int(blk.counts.sum()) == blk.N
    assert block.decompress(blk, ctx) == text.tobytes()
    blob = block.compress_blocks_packed([text], True, ctx)[0]
    u = block.unpack_block(blob)
    assert np.array_equal(u.counts, blk.counts) and np.array_equal(u.syms, blk.syms)
    assert u.final_list.tolist() == blk.final_list.tolist() and blob.size < 2.2 * text.size
    assert block.decompress_packed(blob, ctx) == text.tobytes()
    blk2 = block.compress_bwt_rle(gen_acgtn(0xC6, 16 << 20), ctx)
    assert block.decompress(blk2, ctx) == gen_acgtn(0xC6, 16 << 20).tobytes()


# ---------------------------------------------------------------- reference's own tests, through the API
def test_reference_hunit_vectors(ctx, golden):
    from text_compression_b200 import rle as R, mtf as M
    from text_compression_b200.seq import MTF, RLE, MaybeSeq
    r1, r2 = golden["rle"]
    # Data/RLE.hs:316-319
    assert R.textToBWTToRLET(r1["text"], ctx).to_list() == r1["rle"]
    assert R.textToBWTToRLEB(r2["text"], ctx).to_list() == [None if x is None else x.encode() for x in r2["rle"]]
    assert R.textFromBWTFromRLET(RLE.from_list(r1["rle"], "T"), ctx) == r1["text"]
    assert R.textFromBWTFromRLET(RLE.from_list(r2["rle"], "T"), ctx) == r2["text"]
    # Data/MTF.hs:290-298
    m = golden["mtf"][0]
    got = M.textToBWTToMTFB(m["text"], ctx)
    assert got.to_tuple() == (m["indices"], [None if x is None else x.encode() for x in m["final_list"]])
    mm = MTF(np.array(m["indices"], np.int64), MaybeSeq.from_list([None if x is None else x.encode() for x in m["final_list"]], "B"))
    assert M.textFromBWTFromMTFB(mm, ctx) == m["text"]


def test_fm_doc_tables(ctx, golden):
    from text_compression_b200 import fmindex as F
    d = golden["fmindex_doc"]
    fm = F.textToBWTToFMIndexT(d["text"], ctx=ctx)
    assert "".join("$" if c is None else c for c in F.textFromFMIndexT(fm).to_list()) == d["bwt"]
    assert fm.Cc == list(zip(d["C_vals"], [None if c == "$" else c for c in d["C_syms"]]))
    occ = fm.OccCK
    for (sym, row), c in zip(occ, d["C_syms"]):
        assert [o for (_, o, _) in row] == d["occ"][c]
    assert F.countFMIndex("abra", fm) == 2
    assert F.locateFMIndex("abra", fm) == [8, 1]
    assert F.locateFMIndex("a", fm) == [11, 8, 1, 4, 6]
    assert F.countFMIndex("xa", fm) == 5 and F.countFMIndex("ax", fm) is None and F.countFMIndex("", fm) is None
    assert F.textFMIndexCountS(["abra", "zzz", "a"], d["text"], ctx=ctx) == [("abra", 2), ("zzz", None), ("a", 5)]
    assert F.bytestringFMIndexLocateS([b"abra"], d["text"].encode(), ctx=ctx) == [(b"abra", [8, 1])]
    assert F.bytestringFMIndexCountS([], b"abc", ctx=ctx) == [] and F.bytestringFMIndexCountS([b"a"], b"", ctx=ctx) == []


# ---------------------------------------------------------------- FM-index vs oracle
@pytest.mark.parametrize("rate", [1, 4, 32])
def test_fm_count_locate(ctx, orc, rate):
    from text_compression_b200.fmindex import FMIndex
    rng = np.random.default_rng(rate)
    text = gen_acgtn(0xC3, 60000)
    fm = FMIndex(text, "B", rate, ctx)
    ofm = orc.FMIndex(text)
    assert fm.Cc[0][0] == 0 and [c for c, _ in fm.Cc] == ofm.Cc.tolist()
    pats = []
    for _ in range(3000):
        m = int(rng.integers(1, 40))
        o = int(rng.integers(0, text.size - m))
        p = text[o:o + m].copy()
        u = rng.random()
        if u < 0.15:
            p[rng.integers(0, m)] = rng.choice(np.frombuffer(b"ACGTNXZ", np.uint8))
        pats.append(p.tobytes())
    pats += [b"", b"X", b"AX", b"XA", b"A", b"N", b"NN"]
    got = fm.count_many(pats)
    want = np.array([ofm.count(p) for p in pats])
    assert np.array_equal(got, want), np.nonzero(got != want)[0][:10]
    sel = [p for p in pats if len(p) >= 3][:600] + [b"AC", b"XA", b""]
    ho, pos = fm.locate_many(sel)
    for i, p in enumerate(sel):
        assert pos[ho[i]:ho[i + 1]].tolist() == ofm.locate(p).tolist(), (i, p)
    if rate == 1:
        assert [s for _, s in fm.SA] == ofm.sa.tolist()
    fm.close()


def test_fm_bytes_alphabet(ctx, orc):
    from text_compression_b200.fmindex import FMIndex
    text = gen_bytes(9, 30000)
    fm = FMIndex(text, "B", 16, ctx)
    ofm = orc.FMIndex(text[:30000])
    pats = [text[o:o + 3].tobytes() for o in range(0, 3000, 7)] + [bytes([1, 2, 3, 4, 5])]
    assert fm.count_many(pats).tolist() == [ofm.count(p) for p in pats]
    ho, pos = fm.locate_many(pats[:50])
    for i, p in enumerate(pats[:50]):
        assert pos[ho[i]:ho[i + 1]].tolist() == ofm.locate(p).tolist()


# ---------------------------------------------------------------- composed helpers
@pytest.mark.parametrize("gen,n", [(gen_acgt, 65536), (gen_bytes, 100000), (gen_acgtn, 30011), (gen_ascii, 4096)])
def test_composites(ctx, orc, gen, n):
    from text_compression_b200 import block
    text = gen(0xC1, n)
    bwt = orc.bwt_encode(text)
    b1 = block.compress_bwt_rle(text, ctx)
    cnt, sym = orc.rle_encode(bwt)
    assert b1.counts.tolist() == cnt.tolist() and b1.syms.tolist() == sym.tolist()
    assert block.decompress(b1, ctx) == text.tobytes()
    b2 = block.compress_bwt_mtf_rle(text, ctx)
    idx, fin = orc.mtf_encode(bwt)
    cnt, sym = orc.rle_encode(idx.astype(np.int16))
    assert b2.final_list.tolist() == fin.tolist()
    assert b2.counts.tolist() == cnt.tolist() and b2.syms.tolist() == sym.tolist()
    assert block.decompress(b2, ctx) == text.tobytes()


def test_compress_blocks_pipeline(ctx, orc):
    """tc_blocks_encode (copies overlapped with compute) gives, block by block, exactly what the
    single-block call gives -- ragged sizes, an empty block, both chains."""
    from text_compression_b200 import block
    texts = [gen_acgtn(1, 70001), gen_bytes(2, 4096), np.empty(0, np.uint8), gen_ascii(3, 33333), gen_acgt(4, 5),
             gen_bytes(5, 200000), gen_acgtn(6, 1)]
    for with_mtf in (True, False):
        got = block.compress_blocks(texts, with_mtf, ctx)
        assert len(got) == len(texts)
        for t, g in zip(texts, got):
            one = (block.compress_bwt_mtf_rle if with_mtf else block.compress_bwt_rle)(t, ctx)
            assert (g.n, g.N, g.primary, g.sigma) == (one.n, one.N, one.primary, one.sigma)
            assert g.final_list.tolist() == one.final_list.tolist()
            assert np.array_equal(g.counts, one.counts) and np.array_equal(g.syms, one.syms)
            if with_mtf or g.primary != g.N - 1:   # Q1: the reference's own round trip breaks there
                assert block.decompress(g, ctx) == t.tobytes()
    bwt = orc.bwt_encode(texts[0])
    cnt, sym = orc.rle_encode(bwt)
    g0 = block.compress_blocks(texts[:1], False, ctx, pinned=False)[0]
    assert g0.counts.tolist() == cnt.tolist() and g0.syms.tolist() == sym.tolist()


def test_packed_container(ctx, orc):
    """tc_blocks_encode_packed writes, byte for byte, the container an independent numpy writer lays
    out from the ORACLE's runs (long runs -> count exceptions in run order, Nothing symbols, MTF
    index 256, ragged and empty blocks); host unpack and device decode invert it."""
    from tests.util import pack_container
    from text_compression_b200 import block
    texts = [gen_acgtn(1, 70001), gen_bytes(2, 4096), np.empty(0, np.uint8), gen_ascii(3, 33333), gen_acgt(4, 5),
             np.frombuffer(b"a" * 300000 + b"b" * 255 + b"c" * 254 + b"ab" * 4000 + b"z" * 70000, np.uint8),
             np.tile(np.arange(256, dtype=np.uint8), 40), np.repeat(gen_bytes(8, 600), 700), gen_bytes(5, 200000),
             gen_acgtn(6, 1)]
    for with_mtf in (True, False):
        got = block.compress_blocks_packed(texts, with_mtf, ctx)
        assert len(got) == len(texts)
        for t, blob in zip(texts, got):
            if t.size == 0:
                want = pack_container(0, 0, 0, 0, [], [], [], with_mtf)
            else:
                bwt = orc.bwt_encode(t)
                primary = int(np.nonzero(bwt < 0)[0][0])
                if with_mtf:
                    idx, fin = orc.mtf_encode(bwt)
                    cnt, sym = orc.rle_encode(idx.astype(np.int16))
                else:
                    fin = np.empty(0, np.int16)
                    cnt, sym = orc.rle_encode(bwt)
                want = pack_container(t.size, t.size + 1, primary, len(fin), fin, cnt, sym, with_mtf)
            assert blob.size == want.size and np.array_equal(blob, want), (t.size, with_mtf)
            u = block.unpack_block(blob)
            if t.size:
                assert np.array_equal(u.counts, cnt) and np.array_equal(u.syms, sym)
            if with_mtf or u.primary != u.N - 1 or t.size == 0:   # Q1: the reference's own round trip breaks there
                assert block.decompress_packed(blob, ctx) == t.tobytes()


def test_device_resident_batch(ctx):
    """tc_blocks_encode_dev (texts and runs stay in HBM, two blocks in flight) and the single-block
    tc_bwt_mtf_rle_encode_dev give, block by block, the records of the host-buffer call."""
    import ctypes as C
    import torch
    from text_compression_b200 import block
    from text_compression_b200._lib import BlockInfo
    texts = [gen_bytes(11, 150001), gen_acgtn(12, 90000), gen_ascii(13, 4097), np.empty(0, np.uint8), gen_bytes(14, 70000),
             gen_acgt(15, 33), gen_acgtn(16, 250000)]
    nb = len(texts)
    d_text = [torch.from_numpy(t.copy()).cuda() if t.size else torch.empty(1, dtype=torch.uint8, device="cuda") for t in texts]
    d_cnt = [torch.empty(t.size + 3, dtype=torch.int32, device="cuda") for t in texts]
    d_sym = [torch.empty(t.size + 3, dtype=torch.int16, device="cuda") for t in texts]
    torch.cuda.synchronize()
    for with_mtf in (1, 0):
        tp = (C.c_void_p * nb)(*[x.data_ptr() for x in d_text])
        cp = (C.c_void_p * nb)(*[x.data_ptr() for x in d_cnt])
        sp = (C.c_void_p * nb)(*[x.data_ptr() for x in d_sym])
        ns = (C.c_uint64 * nb)(*[t.size for t in texts])
        caps = (C.c_uint64 * nb)(*[t.size + 3 for t in texts])
        infos = (BlockInfo * nb)()
        ctx.call("tc_blocks_encode_dev", nb, tp, ns, with_mtf, cp, sp, caps, infos)
        for b, t in enumerate(texts):
            one = (block.compress_bwt_mtf_rle if with_mtf else block.compress_bwt_rle)(t, ctx)
            R = int(infos[b].R)
            assert (R, int(infos[b].primary), int(infos[b].sigma)) == (one.R, one.primary, one.sigma)
            assert np.array_equal(d_cnt[b][:R].cpu().numpy().view(np.uint32), one.counts)
            assert np.array_equal(d_sym[b][:R].cpu().numpy(), one.syms)
            assert list(infos[b].final_list[: one.sigma]) == one.final_list.tolist()
    info = BlockInfo()
    ctx.call("tc_bwt_mtf_rle_encode_dev", C.c_void_p(d_text[0].data_ptr()), texts[0].size, C.c_void_p(d_cnt[1].data_ptr()),
             C.c_void_p(d_sym[1].data_ptr()), texts[1].size + 3, C.byref(info), allow=(-2,))
    one = block.compress_bwt_mtf_rle(texts[0], ctx)
    assert int(info.R) == one.R       # capacity of block 1's buffers is too small: TC_E_CAP, R still reported


def test_stream_roundtrip(ctx):
    """TCZ1 stream (text_compression_b200/stream.py): blocks of a byte string as packed containers --
    ragged last block, empty input, both chains, a block size that gives many tiny blocks."""
    from text_compression_b200 import stream
    data = gen_ascii(21, 100000).tobytes() + gen_acgtn(22, 50001).tobytes() + b"a" * 3000
    for bs, with_mtf in ((1 << 16, True), (40000, False), (1 << 20, True), (777, True)):
        z = stream.compress_stream(data, bs, with_mtf, ctx)
        block_bytes, total, parts = stream.split_stream(z)
        assert (block_bytes, total, len(parts)) == (bs, len(data), -(-len(data) // bs))
        if with_mtf:   # the BWT -> RLE chain inherits the reference's Q1 round-trip break on some inputs
            assert stream.decompress_stream(z, ctx) == data
    assert stream.decompress_stream(stream.compress_stream(b"", 1024, True, ctx), ctx) == b""
    with pytest.raises(ValueError):
        stream.split_stream(z[:-5])


def test_q1_trailing_nothing_stream(ctx, orc):
    """Texts that are their own greatest suffix: the reference's RLE re-emits a stale pair (Q1)
    and its own round trip breaks; the GPU stream must equal the oracle's, not round-trip."""
    from text_compression_b200 import block
    for t in (b"a", b"ba", b"aaaa", b"cba"):
        b1 = block.compress_bwt_rle(t, ctx)
        cnt, sym = orc.rle_encode(orc.bwt_encode(t))
        assert list(zip(b1.counts.tolist(), b1.syms.tolist())) == list(zip(cnt.tolist(), sym.tolist())), t


# ---------------------------------------------------------------- full-size properties
def test_block_16mib_roundtrip(ctx):
    from text_compression_b200 import block
    text = gen_bytes(0xC2, 16 << 20)
    blk = block.compress_bwt_mtf_rle(text, ctx)
    assert blk.N == text.size + 1 and int(blk.counts.sum()) == blk.N      # run lengths cover the BWT
    assert sorted(blk.final_list.tolist()) == [-1] + list(range(256))
    assert block.decompress(blk, ctx) == text.tobytes()
    blob = block.compress_blocks_packed([text], True, ctx)[0]
    u = block.unpack_block(blob)
    assert np.array_equal(u.counts, blk.counts) and np.array_equal(u.syms, blk.syms)
    assert u.final_list.tolist() == blk.final_list.tolist() and blob.size < 2.2 * text.size
    assert block.decompress_packed(blob, ctx) == text.tobytes()
