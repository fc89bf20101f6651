"""The polymorphic corners of the reference API through the byte kernels: `toBWT :: Ord a`
on arbitrary ordered elements, MTF / RLE on multi-byte Pack items, createBWTMatrix as a view, sortTB and
magicInverseBWT by name -- each against the oracle (run on independently assigned ranks)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _ranked(items):
    """Independent rank assignment for the oracle: sorted distinct non-None elements -> 0..k-1, None -> -1."""
    syms = sorted({x for x in items if x is not None})
    rk = {s: i for i, s in enumerate(syms)}
    return np.array([-1 if x is None else rk[x] for x in items], dtype=np.int16), syms


def _unrank(codes, syms):
    return [None if c < 0 else syms[c] for c in np.asarray(codes).tolist()]


WORDS = "the quick brown fox jumps over the lazy dog and the quick dog naps while the fox runs".split()


@pytest.mark.parametrize("xs", [
    WORDS * 40,
    [(i * 7919) % 13 for i in range(5000)],                       # ints
    [(w, len(w)) for w in WORDS * 9],                             # tuples
    [3.5, -1.25, 3.5, 0.0, 7.75, -1.25] * 50,                     # floats
    ["été", "naïve", "über", "naïve"] * 30,   # multi-byte Text items
    ["x"],
])
def test_generic_bwt_vs_oracle(orc, xs):
    from text_compression_b200 import generic
    codes, syms = _ranked(xs)
    o_bwt, o_sa = orc.bwt_encode(codes.astype(np.uint8), want_sa=True)
    assert generic.toBWT(xs) == _unrank(o_bwt, syms)
    assert [p for _, p in generic.createSuffixArray(xs)] == o_sa.tolist()
    assert generic.fromBWT(generic.toBWT(xs)) == list(xs)


def test_generic_bwt_limits(orc):
    from text_compression_b200 import generic
    assert generic.toBWT([]) == [] and generic.fromBWT([]) == []
    with pytest.raises(generic.TooManySymbols):
        generic.toBWT(list(range(300)))
    assert generic.toBWT(list(range(256))) == [255, None] + list(range(255))   # 256 distinct elements still fit
    # malformed columns: no Nothing -> [], like the reference
    assert generic.fromBWT(["b", "a"]) == []


@pytest.mark.parametrize("items", [
    [b"ab", b"ab", None, b"c", b"c", b"c", b"ab"],
    [None, b"xyz", b"xyz", b"q"],                                  # Q2: a leading Nothing is dropped
    [b"aa", b"aa", None, None, b"bb"],                             # Q3
    [b"long-item"] * 300 + [None],                                 # Q1 + a three-digit count
    ["é", "é", "z", None, "z"],
    [WORDS[i % len(WORDS)].encode() for i in range(2000)],
])
def test_generic_rle_mtf_vs_oracle(orc, items):
    from text_compression_b200 import generic
    codes, syms = _ranked(items)
    cnt, rs = orc.rle_encode(codes)
    like_bytes = isinstance(next(x for x in items if x is not None), bytes)
    want = []
    for c, s in zip(cnt.tolist(), _unrank(rs, syms)):
        want += [str(c).encode() if like_bytes else str(c), s]
    got = generic.seqToRLE(items)
    assert got == want
    assert generic.seqFromRLE(got) == _unrank(orc.rle_decode(cnt, rs), syms)
    idx, fin = orc.mtf_encode(codes)
    g_idx, g_fin = generic.seqToMTF(items)
    assert g_idx == idx.tolist() and g_fin == _unrank(fin, syms)
    assert generic.seqFromMTF(g_idx, g_fin) == list(items)


def test_bwt_matrix_view(orc, golden):
    """createBWTMatrix on the reference's worked example (src/Data/FMIndex/Internal.hs:49-113): first column = the
    sorted symbols (what C[c] is derived from), last column = the documented BWT, rows = sorted rotations."""
    from text_compression_b200 import generic
    doc = golden["fmindex_doc"]
    t = list(doc["text"])
    m = generic.createBWTMatrix(t)
    assert len(m) == len(t) + 1
    dollar = lambda col: "".join("$" if c is None else c for c in col)
    assert dollar(m.last_column()) == doc["bwt"]
    assert dollar(m.first_column()) == "$" + "".join(sorted(t))
    rows = m.to_list()
    rot = t + [None]
    key = lambda r: [(-1,) if c is None else (0, c) for c in r]
    assert rows == sorted((rot[k:] + rot[:k] for k in range(len(rot))), key=key)
    # C[c] from the first column (seqToCc): first index of every symbol
    f = m.first_column()
    assert [f.index(None if s == "$" else s) for s in doc["C_syms"]] == doc["C_vals"]
    assert len(generic.createBWTMatrix([])) == 0
    # a bigger one, against the oracle's suffix array
    xs = [(i * 31) % 7 for i in range(3000)]
    big = generic.createBWTMatrix(xs)
    sa = orc.bwt_encode(np.array(xs, dtype=np.uint8), want_sa=True)[1].tolist()
    assert big.sa.tolist() == sa and big.row(5) == (xs + [None])[sa[5] - 1:] + (xs + [None])[:sa[5] - 1]


def test_sorttb_and_magic_inverse(orc):
    from text_compression_b200 import generic
    from text_compression_b200._lib import FromJustError
    assert generic.sortTB((None, 5), ("a", 0)) < 0 and generic.sortTB(("a", 3), ("a", 1)) > 0
    assert generic.sortTB(("a", 1), ("b", 0)) < 0 and generic.sortTB((None, 2), (None, 2)) == 0
    for text in ("abracadabra", "mississippi", "aaaa", "ba"):
        col = generic.toBWT(list(text))
        pairs = generic.sort_pairs(col)
        want = orc.bwt_decode(_ranked(col)[0])          # what the reference's fromBWT gives on this column
        syms = sorted(set(text))
        assert generic.magicInverseBWT(pairs) == [syms[c] for c in want.tolist()]
    # pairs that are NOT a sorted column: walked on the host like the reference
    assert generic.magicInverseBWT([("x", 1), (None, 0)]) == ["x"]   # e = 1, f = 0: emits pairs[0], then f = 1 = e
    assert generic.magicInverseBWT([("a", 0), ("b", 1)]) == []
    with pytest.raises(FromJustError):
        generic.magicInverseBWT([(None, 1), (None, 0)])


def test_generic_fm_index_vs_oracle(orc):
    """FM-index over words, ints and tuples (rank-compressed on the host): counts and located positions equal the
    oracle's on the same codes; a pattern with an element the text lacks, and the empty pattern, give Nothing."""
    from text_compression_b200 import generic
    rng = np.random.default_rng(11)
    vocab = ["the", "quick", "brown", "fox", "jumps", "over", "lazy", "dog", "and", "cat"]
    texts = [[vocab[i] for i in rng.integers(0, len(vocab), size=5000)],
             rng.integers(-50, 50, size=20000).tolist(),
             [(int(a), "xy"[int(b)]) for a, b in zip(rng.integers(0, 7, size=3000), rng.integers(0, 2, size=3000))]]
    for xs in texts:
        fm = generic.FMIndexG(xs, 4)
        al = generic.Alphabet(xs)
        ofm = orc.FMIndex(al.encode(xs))
        pats = [xs[o:o + m] for o, m in zip(rng.integers(0, len(xs) - 8, size=200).tolist(), rng.integers(1, 6, size=200).tolist())]
        want_c = [ofm.count(bytes(al.rank[x] for x in p)) for p in pats]
        got_c = fm.count_many(pats)
        assert got_c == [None if c < 0 else c for c in want_c]
        got_l = fm.locate_many(pats[:50])
        for p, g in zip(pats[:50], got_l):
            assert g == ofm.locate(bytes(al.rank[x] for x in p)).tolist()
            for pos in g:
                assert xs[pos - 1: pos - 1 + len(p)] == p
        assert fm.count([]) is None and fm.count([xs[0], object]) is None and fm.locate([object]) == []
        fm.close()
