"""Data.RLE on the B200 (mirror of src/Data/RLE.hs + src/Data/RLE/Internal.hs)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import TC_E_CAP, default_context, ptr
from .bwt import bytestringFromByteStringBWT, bytestringToBWT, textToBWT
from .seq import BWT, RLE, MaybeSeq, TextBWT

__all__ = [
    "seqToRLE", "seqFromRLE",
    "bytestringToBWTToRLEB", "bytestringToBWTToRLET", "textToBWTToRLEB", "textToBWTToRLET", "textBWTToRLEB",
    "bytestringBWTToRLEB", "textBWTToRLET", "bytestringBWTToRLET", "textToRLEB", "bytestringToRLEB", "textToRLET",
    "bytestringToRLET",
    "bytestringFromBWTFromRLEB", "bytestringFromBWTFromRLET", "textFromBWTFromRLEB", "textFromBWTFromRLET",
    "textBWTFromRLET", "bytestringBWTFromRLET", "textBWTFromRLEB", "bytestringBWTFromRLEB", "textFromRLEB",
    "bytestringFromRLEB", "textFromRLET", "bytestringFromRLET",
]


def seqToRLE(xs: MaybeSeq, ctx=None) -> RLE:
    """seqToRLE (src/Data/RLE/Internal.hs:104-153), Nothing quirks included."""
    ctx = ctx or default_context()
    N = len(xs)
    cap = 2 * N + 1
    cnt = np.empty(cap, dtype=np.uint32)
    sym = np.empty(cap, dtype=np.int16)
    R = C.c_uint64(0)
    ctx.call("tc_rle_encode", ptr(xs.codes), N, ptr(cnt), ptr(sym), cap, C.byref(R))
    return RLE(cnt[: R.value].copy(), sym[: R.value].copy(), xs.kind)


def seqFromRLE(r: RLE, ctx=None) -> MaybeSeq:
    """seqFromRLE (src/Data/RLE/Internal.hs:155-189)."""
    ctx = ctx or default_context()
    Rn = int(r.counts.size)
    if Rn == 0:
        return MaybeSeq(np.empty(0, dtype=np.int16), r.kind)
    N = C.c_uint64(0)
    rc = ctx.call("tc_rle_decode", ptr(r.counts), ptr(r.syms), Rn, None, 0, C.byref(N), allow=(TC_E_CAP,))
    out = np.empty(int(N.value), dtype=np.int16)
    if out.size:
        ctx.call("tc_rle_decode", ptr(r.counts), ptr(r.syms), Rn, ptr(out), out.size, C.byref(N))
    return MaybeSeq(out, r.kind)


def _bwt_seq(x) -> MaybeSeq:
    if isinstance(x, TextBWT):
        x = x.bwt
    return x.seq


def _empty_rle(kind):
    return RLE(np.empty(0, dtype=np.uint32), np.empty(0, dtype=np.int16), kind)


def _bwt_to_rle(x, kind, ctx):
    s = _bwt_seq(x)
    if len(s) == 0:
        return _empty_rle(kind)
    return seqToRLE(s.as_kind(kind), ctx)


# ---- to RLE (src/Data/RLE.hs:83-175) ----------------------------------------------------
def bytestringToBWTToRLEB(bs, ctx=None): return bytestringBWTToRLEB(bytestringToBWT(bs, ctx), ctx)
def bytestringToBWTToRLET(bs, ctx=None): return bytestringBWTToRLET(bytestringToBWT(bs, ctx), ctx)
def textToBWTToRLEB(t, ctx=None): return textBWTToRLEB(textToBWT(t, ctx), ctx)
def textToBWTToRLET(t, ctx=None): return textBWTToRLET(textToBWT(t, ctx), ctx)
def textBWTToRLEB(xs, ctx=None): return _bwt_to_rle(xs, "B", ctx)
def bytestringBWTToRLEB(xs, ctx=None): return _bwt_to_rle(xs, "B", ctx)
def textBWTToRLET(xs, ctx=None): return _bwt_to_rle(xs, "T", ctx)
def bytestringBWTToRLET(xs, ctx=None): return _bwt_to_rle(xs, "T", ctx)


def _seq_to_rle(xs, kind, ctx):
    xs = xs if isinstance(xs, MaybeSeq) else MaybeSeq.from_list(xs, kind)
    if len(xs) == 0:
        return _empty_rle(kind)
    return seqToRLE(xs.as_kind(kind), ctx)


def textToRLEB(xs, ctx=None): return _seq_to_rle(xs, "B", ctx)
def bytestringToRLEB(xs, ctx=None): return _seq_to_rle(xs, "B", ctx)
def textToRLET(xs, ctx=None): return _seq_to_rle(xs, "T", ctx)
def bytestringToRLET(xs, ctx=None): return _seq_to_rle(xs, "T", ctx)


# ---- from RLE (src/Data/RLE.hs:184-275) --------------------------------------------------
def _as_rle(r, kind="B") -> RLE:
    return r if isinstance(r, RLE) else RLE.from_list(r, kind)


def _bwt_from_rle(r, kind, ctx) -> BWT:
    return BWT(seqFromRLE(_as_rle(r), ctx).as_kind(kind))


def textBWTFromRLET(r, ctx=None): return _bwt_from_rle(r, "T", ctx)
def bytestringBWTFromRLET(r, ctx=None): return _bwt_from_rle(r, "B", ctx)
def textBWTFromRLEB(r, ctx=None): return _bwt_from_rle(r, "T", ctx)
def bytestringBWTFromRLEB(r, ctx=None): return _bwt_from_rle(r, "B", ctx)
def bytestringFromBWTFromRLEB(r, ctx=None): return bytestringFromByteStringBWT(bytestringBWTFromRLEB(r, ctx), ctx)
def bytestringFromBWTFromRLET(r, ctx=None): return bytestringFromByteStringBWT(textBWTFromRLET(r, ctx), ctx)
def textFromBWTFromRLEB(r, ctx=None): return bytestringFromBWTFromRLEB(r, ctx).decode("utf-8")
def textFromBWTFromRLET(r, ctx=None): return bytestringFromByteStringBWT(bytestringBWTFromRLET(r, ctx), ctx).decode("utf-8")
def textFromRLEB(r, ctx=None): return seqFromRLE(_as_rle(r), ctx).as_kind("T")
def bytestringFromRLEB(r, ctx=None): return seqFromRLE(_as_rle(r), ctx).as_kind("B")
def textFromRLET(r, ctx=None): return seqFromRLE(_as_rle(r, "T"), ctx).as_kind("T")
def bytestringFromRLET(r, ctx=None): return seqFromRLE(_as_rle(r, "T"), ctx).as_kind("B")
