"""The polymorphic corners of the reference API, through the same kernels.

  * `toBWT :: Ord a => [a] -> BWT a` / `fromBWT` on ANY ordered elements (src/Data/BWT.hs:55,93) and the
    MTF / RLE kernels on multi-byte `Pack` items (arbitrary ByteString / Text elements,
    src/Data/RLE/Internal.hs:66-90, src/Data/MTF/Internal.hs:128-232): only the order (BWT, MTF alphabet) and
    the equality (RLE) of the elements are ever used, so an input with at most 256 distinct elements is
    rank-compressed on the host -- dense ranks in alphabet order, an order isomorphism -- and runs through the
    byte kernels.  More than 256 distinct elements raise `TooManySymbols` (the kernels are 8-bit; the
    Haskell shim falls back to a host sort there).
  * `createBWTMatrix` / `BWTMatrix` (src/Data/BWT/Internal.hs:88,209-241): every rotation of text$ in sorted order,
    as a lazily materialised VIEW over the GPU suffix array: row k is the rotation that starts at SA[k]; nothing
    O(n^2) is stored.  The FM-index builders only read its first column.
  * `sortTB` and `magicInverseBWT` by name (src/Data/BWT/Internal.hs:144-200).

Values use the Haskell shapes directly: a `Seq (Maybe a)` is a Python list with None for Nothing.
"""
from __future__ import annotations

import ctypes as C
from functools import cmp_to_key

import numpy as np

from ._lib import FromJustError, TC_E_FROMJUST, default_context, ptr

__all__ = ["TooManySymbols", "Alphabet", "toBWT", "fromBWT", "createSuffixArray", "seqToMTF", "seqFromMTF", "seqToRLE",
           "seqFromRLE", "BWTMatrix", "createBWTMatrix", "sortTB", "magicInverseBWT", "FMIndexG"]


class TooManySymbols(ValueError):
    """More than 256 distinct elements: outside the byte kernels."""


class Alphabet:
    """Sorted distinct elements of an input and their dense ranks."""

    def __init__(self, items):
        self.symbols = sorted(set(items))
        if len(self.symbols) > 256:
            raise TooManySymbols(f"{len(self.symbols)} distinct elements (at most 256 reach the GPU kernels)")
        self.rank = {s: r for r, s in enumerate(self.symbols)}

    def encode(self, items) -> np.ndarray:
        return np.fromiter((self.rank[x] for x in items), dtype=np.uint8, count=len(items))

    def encode_maybe(self, items) -> np.ndarray:
        return np.fromiter((-1 if x is None else self.rank[x] for x in items), dtype=np.int16, count=len(items))

    def decode_maybe(self, codes) -> list:
        return [None if c < 0 else self.symbols[c] for c in np.asarray(codes).tolist()]


def _alphabet_of_maybe(items) -> Alphabet:
    return Alphabet([x for x in items if x is not None])


# ---- Data.BWT on any ordered elements ----------------------------------------------------------------------
def createSuffixArray(xs, ctx=None) -> list:
    """createSuffixArray :: Ord a => Seq a -> SuffixArray a: [(suffixindex, suffixstartpos)] in rank order."""
    from .bwt import createSuffixArray as sa_bytes
    xs = list(xs)
    sa = sa_bytes(Alphabet(xs).encode(xs), ctx) if xs else np.array([1], dtype=np.uint32)
    return [(k + 1, int(p)) for k, p in enumerate(sa.tolist())]


def toBWT(xs, ctx=None) -> list:
    """toBWT :: Ord a => [a] -> BWT a, as a list with None for the Nothing."""
    from .bwt import bwt_u8
    xs = list(xs)
    if not xs:
        return []
    al = Alphabet(xs)
    bwt, primary, _ = bwt_u8(al.encode(xs), ctx=ctx)
    out = [al.symbols[c] for c in bwt.tolist()]
    out[primary] = None
    return out


def fromBWT(bwt, ctx=None) -> list:
    """fromBWT :: Ord a => BWT a -> [a]; malformed columns behave like the reference (no Nothing -> [],
    a second Nothing on the walk -> fromJust)."""
    ctx = ctx or default_context()
    bwt = list(bwt)
    if not bwt:
        return []
    al = _alphabet_of_maybe(bwt)
    codes = al.encode_maybe(bwt)
    out = np.empty(len(bwt), dtype=np.uint8)
    n_out = C.c_uint64(0)
    ctx.call("tc_bwt_decode", ptr(codes), codes.size, ptr(out), out.size, C.byref(n_out))
    return [al.symbols[c] for c in out[: n_out.value].tolist()]


# ---- Data.MTF / Data.RLE on multi-byte Pack items ------------------------------------------------------------
def seqToMTF(xs, ctx=None):
    """seqToMTF on any `Seq (Maybe b)`: (indices, FINAL list)."""
    from .mtf import seqToMTF as mtf_codes
    from .seq import MaybeSeq
    xs = list(xs)
    if not xs:
        return [], []
    al = _alphabet_of_maybe(xs)
    m = mtf_codes(MaybeSeq(al.encode_maybe(xs), "W"), ctx)
    return m.indices.tolist(), al.decode_maybe(m.final_list.codes)


def seqFromMTF(indices, final_list, ctx=None) -> list:
    from .mtf import seqFromMTF as unmtf_codes
    from .seq import MTF, MaybeSeq
    final_list = list(final_list)
    if not len(indices) or not final_list:
        return []
    al = _alphabet_of_maybe(final_list)
    s = unmtf_codes(MTF(np.asarray(indices, dtype=np.int64), MaybeSeq(al.encode_maybe(final_list), "W")), ctx)
    return al.decode_maybe(s.codes)


def _render(count: int, like):
    return str(count).encode() if isinstance(like, (bytes, bytearray)) else str(count)


def seqToRLE(xs, ctx=None) -> list:
    """seqToRLE on any `Seq (Maybe b)`, flat like the reference: [Just (show count), symbol, ...]; counts are
    rendered in the item type (bytes items -> bytes counts, str items -> str counts), Q1-Q3 included."""
    from .rle import seqToRLE as rle_codes
    from .seq import MaybeSeq
    xs = list(xs)
    if not xs:
        return []
    al = _alphabet_of_maybe(xs)
    like = next((x for x in xs if x is not None), b"")
    r = rle_codes(MaybeSeq(al.encode_maybe(xs), "W"), ctx)
    out = []
    for c, s in zip(r.counts.tolist(), al.decode_maybe(r.syms)):
        out.append(_render(c, like))
        out.append(s)
    return out


def seqFromRLE(flat, ctx=None) -> list:
    """seqFromRLE on the flat form; an odd trailing element is ignored; Nothing in a count slot is the
    reference's fromJust."""
    from .rle import seqFromRLE as unrle_codes
    from .seq import RLE
    flat = list(flat)
    pairs = [(flat[k], flat[k + 1]) for k in range(0, len(flat) - 1, 2)]
    if not pairs:
        return []
    if any(y1 is None for y1, _ in pairs):
        raise FromJustError(TC_E_FROMJUST, "Nothing in a count slot of an RLE value")
    al = _alphabet_of_maybe([s for _, s in pairs])
    cnt = np.array([max(int(y1.decode() if isinstance(y1, (bytes, bytearray)) else y1), 0) for y1, _ in pairs], dtype=np.uint32)
    s = unrle_codes(RLE(cnt, al.encode_maybe([s for _, s in pairs]), "W"), ctx)
    return al.decode_maybe(s.codes)


# ---- Data.FMIndex on any ordered elements ----------------------------------------------------------------------
class FMIndexG:
    """FM-index over a sequence of ANY ordered elements with at most 256 distinct values: the text is
    rank-compressed into bytes and indexed on the device once; the handle persists, so a batch of queries does not
    rebuild the index as the reference's wrappers do (src/Data/FMIndex.hs:368,419,481,546).  A pattern is a sequence
    of elements; one that contains an element the text does not have, or no element at all, has no occurrence
    (countFMIndex gives Nothing: src/Data/FMIndex/Internal.hs:347-438)."""

    def __init__(self, xs, sa_sample_rate: int = 1, ctx=None):
        from .fmindex import FMIndex
        xs = list(xs)
        self.alphabet = Alphabet(xs)
        self.fm = FMIndex(self.alphabet.encode(xs) if xs else None, "B", sa_sample_rate, ctx)

    def _codes(self, pat):
        pat = list(pat)
        if not pat or any(x not in self.alphabet.rank for x in pat):
            return None
        return bytes(self.alphabet.rank[x] for x in pat)

    def count_many(self, pats) -> list:
        """countFMIndex per pattern: the number of occurrences, None for Nothing."""
        coded = [self._codes(p) for p in pats]
        live = [k for k, c in enumerate(coded) if c is not None]
        out = [None] * len(coded)
        if live:
            got = self.fm.count_many([coded[k] for k in live])
            for k, v in zip(live, got.tolist()):
                out[k] = None if v < 0 else int(v)
        return out

    def count(self, pat):
        return self.count_many([pat])[0]

    def locate_many(self, pats) -> list:
        """locateFMIndex per pattern: the 1-based text positions of its occurrences in suffix-array order
        (src/Data/FMIndex.hs:473-474), [] when there is none."""
        coded = [self._codes(p) for p in pats]
        live = [k for k, c in enumerate(coded) if c is not None]
        out = [[] for _ in coded]
        if live:
            ho, pos = self.fm.locate_many([coded[k] for k in live])
            for i, k in enumerate(live):
                out[k] = pos[int(ho[i]):int(ho[i + 1])].astype(np.int64).tolist()
        return out

    def locate(self, pat) -> list:
        return self.locate_many([pat])[0]

    def close(self):
        self.fm.close()


# ---- BWT matrix as a view, sortTB, magicInverseBWT -----------------------------------------------------------
class BWTMatrix:
    """newtype BWTMatrix a = BWTMatrix (Seq (Seq (Maybe a))): the n+1 rotations of text$ sorted by their
    suffix.  Rows are produced on demand from the suffix array (row k starts at text position SA[k])."""

    def __init__(self, text, sa_1based):
        self.text = list(text)
        self.sa = np.asarray(sa_1based, dtype=np.int64)

    def __len__(self):
        return int(self.sa.size) if self.text else 0

    def row(self, k: int) -> list:
        p = int(self.sa[k]) - 1                       # 0-based start; p == n is the rotation that starts with $
        return self.text[p:] + [None] + self.text[:p]

    def first_column(self) -> list:
        """F column: what the FM-index builders read (src/Data/FMIndex.hs:150-155)."""
        n = len(self.text)
        return [None if int(p) == n + 1 else self.text[int(p) - 1] for p in self.sa.tolist()] if self.text else []

    def last_column(self) -> list:
        """L column = the BWT."""
        return [None if int(p) == 1 else self.text[int(p) - 2] for p in self.sa.tolist()] if self.text else []

    def to_list(self) -> list:
        return [self.row(k) for k in range(len(self))]


def createBWTMatrix(t, ctx=None) -> BWTMatrix:
    """createBWTMatrix :: Ord a => [a] -> BWTMatrix a (empty input: the empty matrix; the reference's value for
    it cannot be evaluated)."""
    t = list(t)
    if not t:
        return BWTMatrix([], np.empty(0, dtype=np.int64))
    return BWTMatrix(t, [p for _, p in createSuffixArray(t, ctx)])


def sortTB(a, b) -> int:
    """sortTB (c1,i1) (c2,i2) = compare c1 c2 <> compare i1 i2, Nothing first: -1 / 0 / 1."""
    (c1, i1), (c2, i2) = a, b
    k1, k2 = (c1 is not None, c1), (c2 is not None, c2)
    if k1 != k2:
        if k1[0] != k2[0]:
            return -1 if k2[0] else 1
        return -1 if c1 < c2 else 1
    return (i1 > i2) - (i1 < i2)


def magicInverseBWT(sorted_pairs, ctx=None) -> list:
    """magicInverseBWT :: Seq (Maybe a, Int) -> Seq a on the pairs `fromBWT` sorts with sortTB.  When the pairs
    are what sorting a column by sortTB gives, the column is rebuilt and inverted on the device; any other input
    is walked on the host exactly as the reference walks it."""
    pairs = list(sorted_pairs)
    if not pairs:
        return []
    n = len(pairs)
    idx = [i for _, i in pairs]
    is_sorted = all(sortTB(pairs[k], pairs[k + 1]) < 0 for k in range(n - 1))
    if is_sorted and sorted(idx) == list(range(n)):
        col = [None] * n
        seen = [False] * n
        for c, i in pairs:
            col[i], seen[i] = c, True
        # the sentinel is the only None; a column with several is still what the device inverts faithfully
        return fromBWT(col, ctx)
    e = next((k for k, (c, _) in enumerate(pairs) if c is None), None)
    if e is None:
        return []
    out, f = [], pairs[e][1]
    while f != e:
        c, nxt = pairs[f]
        if c is None:
            raise FromJustError(TC_E_FROMJUST, "Maybe.fromJust: Nothing")
        out.append(c)
        f = nxt
    return out


def sort_pairs(col) -> list:
    """The pairs `fromBWT` builds and sorts: zip column [0..], sorted with sortTB."""
    return sorted(zip(col, range(len(col))), key=cmp_to_key(sortTB))
