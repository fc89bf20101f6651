"""Data.MTF on the B200 (mirror of src/Data/MTF.hs + src/Data/MTF/Internal.hs)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import default_context, ptr
from .bwt import bytestringFromByteStringBWT, bytestringToBWT, textToBWT
from .seq import BWT, MTF, MaybeSeq, TextBWT

__all__ = [
    "nubSeq", "seqToMTF", "seqFromMTF",
    "bytestringToBWTToMTFB", "bytestringToBWTToMTFT", "textToBWTToMTFB", "textToBWTToMTFT", "textBWTToMTFB",
    "bytestringBWTToMTFB", "textBWTToMTFT", "bytestringBWTToMTFT", "textToMTFB", "bytestringToMTFB", "textToMTFT",
    "bytestringToMTFT",
    "bytestringFromBWTFromMTFB", "bytestringFromBWTFromMTFT", "textFromBWTFromMTFB", "textFromBWTFromMTFT",
    "textBWTFromMTFT", "bytestringBWTFromMTFT", "textBWTFromMTFB", "bytestringBWTFromMTFB", "textFromMTFB",
    "bytestringFromMTFB", "textFromMTFT", "bytestringFromMTFT",
]


def nubSeq(xs: MaybeSeq) -> MaybeSeq:
    """nubSeq' (src/Data/MTF/Internal.hs:79-99): sorted distinct elements, Nothing first."""
    return MaybeSeq(np.unique(xs.codes), xs.kind)


def seqToMTF(xs: MaybeSeq, ctx=None) -> MTF:
    """seqToMTF (src/Data/MTF/Internal.hs:128-175): (indices, FINAL list)."""
    ctx = ctx or default_context()
    N = len(xs)
    idx = np.empty(N, dtype=np.uint16)
    fin = np.empty(257, dtype=np.int16)
    sigma = C.c_uint32(0)
    ctx.call("tc_mtf_encode", ptr(xs.codes), N, ptr(idx), ptr(fin), C.byref(sigma))
    return MTF(idx.astype(np.int64), MaybeSeq(fin[: sigma.value].copy(), xs.kind))


def seqFromMTF(m: MTF, ctx=None) -> MaybeSeq:
    """seqFromMTF (src/Data/MTF/Internal.hs:201-232)."""
    ctx = ctx or default_context()
    kind = m.final_list.kind
    N = int(m.indices.size)
    if N == 0 or len(m.final_list) == 0:
        return MaybeSeq(np.empty(0, dtype=np.int16), kind)
    ind = np.asarray(m.indices)
    if ind.min() < 0 or ind.max() > 0xffff:
        from ._lib import SeqIndexError, TC_E_INDEX
        raise SeqIndexError(TC_E_INDEX, "MTF index out of range")
    idx = np.ascontiguousarray(ind, dtype=np.uint16)
    out = np.empty(N, dtype=np.int16)
    ctx.call("tc_mtf_decode", ptr(idx), N, ptr(m.final_list.codes), len(m.final_list), ptr(out))
    return MaybeSeq(out, kind)


def _bwt_seq(x) -> MaybeSeq:
    if isinstance(x, TextBWT):
        x = x.bwt
    return x.seq


# ---- to MTF (src/Data/MTF.hs:82-175) ---------------------------------------------------
def bytestringToBWTToMTFB(bs, ctx=None): return bytestringBWTToMTFB(bytestringToBWT(bs, ctx), ctx)
def bytestringToBWTToMTFT(bs, ctx=None): return bytestringBWTToMTFT(bytestringToBWT(bs, ctx), ctx)
def textToBWTToMTFB(t, ctx=None): return textBWTToMTFB(textToBWT(t, ctx), ctx)
def textToBWTToMTFT(t, ctx=None): return textBWTToMTFT(textToBWT(t, ctx), ctx)
def textBWTToMTFB(xs: TextBWT, ctx=None): return seqToMTF(_bwt_seq(xs).as_kind("B"), ctx)
def bytestringBWTToMTFB(xs: BWT, ctx=None): return seqToMTF(_bwt_seq(xs).as_kind("B"), ctx)
def textBWTToMTFT(xs: TextBWT, ctx=None): return seqToMTF(_bwt_seq(xs).as_kind("T"), ctx)
def bytestringBWTToMTFT(xs: BWT, ctx=None): return seqToMTF(_bwt_seq(xs).as_kind("T"), ctx)


def _empty_mtf(kind):
    return MTF(np.empty(0, dtype=np.int64), MaybeSeq(np.empty(0, dtype=np.int16), kind))


def _seq_to_mtf(xs, kind, ctx):
    xs = xs if isinstance(xs, MaybeSeq) else MaybeSeq.from_list(xs, kind)
    if len(xs) == 0:
        return _empty_mtf(kind)
    return seqToMTF(xs.as_kind(kind), ctx)


def textToMTFB(xs, ctx=None): return _seq_to_mtf(xs, "B", ctx)
def bytestringToMTFB(xs, ctx=None): return _seq_to_mtf(xs, "B", ctx)
def textToMTFT(xs, ctx=None): return _seq_to_mtf(xs, "T", ctx)
def bytestringToMTFT(xs, ctx=None): return _seq_to_mtf(xs, "T", ctx)


# ---- from MTF (src/Data/MTF.hs:184-283) -------------------------------------------------
def _bwt_from_mtf(m: MTF, kind, ctx) -> BWT:
    return BWT(seqFromMTF(m, ctx).as_kind(kind))


def textBWTFromMTFT(m, ctx=None): return _bwt_from_mtf(m, "T", ctx)
def bytestringBWTFromMTFT(m, ctx=None): return _bwt_from_mtf(m, "B", ctx)
def textBWTFromMTFB(m, ctx=None): return _bwt_from_mtf(m, "T", ctx)
def bytestringBWTFromMTFB(m, ctx=None): return _bwt_from_mtf(m, "B", ctx)
def bytestringFromBWTFromMTFB(m, ctx=None): return bytestringFromByteStringBWT(bytestringBWTFromMTFB(m, ctx), ctx)
def bytestringFromBWTFromMTFT(m, ctx=None): return bytestringFromByteStringBWT(textBWTFromMTFT(m, ctx), ctx)
def textFromBWTFromMTFB(m, ctx=None): return bytestringFromBWTFromMTFB(m, ctx).decode("utf-8")
def textFromBWTFromMTFT(m, ctx=None): return bytestringFromByteStringBWT(bytestringBWTFromMTFT(m, ctx), ctx).decode("utf-8")
def textFromMTFB(m, ctx=None): return seqFromMTF(m, ctx).as_kind("T")
def bytestringFromMTFB(m, ctx=None): return seqFromMTF(m, ctx).as_kind("B")
def textFromMTFT(m, ctx=None): return seqFromMTF(m, ctx).as_kind("T")
def bytestringFromMTFT(m, ctx=None): return seqFromMTF(m, ctx).as_kind("B")
