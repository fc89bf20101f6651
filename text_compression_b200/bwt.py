"""Data.BWT on the B200 (mirror of src/Data/BWT.hs + src/Data/BWT/Internal.hs).

Function names, argument meaning and error behaviour follow the reference; the work is
done by libtc_b200.so (tc_bwt_encode / tc_bwt_decode).  No CPU path exists here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import default_context, ptr
from .seq import BWT, MaybeSeq, TextBWT, to_bytes

__all__ = ["toBWT", "bytestringToBWT", "TextBWT", "textToBWT", "fromBWT", "bytestringFromWord8BWT",
           "bytestringFromByteStringBWT", "textFromBWT", "createSuffixArray", "saToBWT", "bwt_u8"]


def _text_array(xs) -> np.ndarray:
    if isinstance(xs, np.ndarray):
        return np.ascontiguousarray(xs, dtype=np.uint8)
    return np.frombuffer(to_bytes(xs), dtype=np.uint8)


def bwt_u8(text, want_sa: bool = False, ctx=None):
    """Raw form: (bwt uint8[n+1], primary, sa_1based or None)."""
    ctx = ctx or default_context()
    t = _text_array(text)
    n = t.size
    bwt = np.empty(n + 1 if n else 0, dtype=np.uint8)
    sa = np.empty(n + 1, dtype=np.uint32) if (want_sa and n) else None
    primary = C.c_uint64(0)
    ctx.call("tc_bwt_encode", ptr(t), n, ptr(bwt), C.byref(primary), ptr(sa))
    return bwt, int(primary.value), sa


def createSuffixArray(xs, ctx=None) -> np.ndarray:
    """createSuffixArray (src/Data/BWT/Internal.hs:110-134): suffixstartpos (1-based) in rank
    order, n+1 entries including the empty suffix; suffixindex is the position + 1."""
    t = _text_array(xs)
    if t.size == 0:
        return np.array([1], dtype=np.uint32)   # DS.tails of the empty Seq: just the empty suffix
    return bwt_u8(t, want_sa=True, ctx=ctx)[2]


def saToBWT(sa_1based: np.ndarray, t) -> MaybeSeq:
    """saToBWT (src/Data/BWT/Internal.hs:98-106) -- a host gather, kept for API completeness."""
    t = _text_array(t)
    sa = np.asarray(sa_1based, dtype=np.int64)
    out = np.where(sa != 1, t[np.maximum(sa - 2, 0)].astype(np.int16) if t.size else np.int16(0), np.int16(-1))
    return MaybeSeq(out.astype(np.int16), "W")


def toBWT(xs, ctx=None) -> BWT:
    """toBWT :: Ord a => [a] -> BWT a   (src/Data/BWT.hs:55-64), for Word8 symbols."""
    bwt, primary, _ = bwt_u8(xs, ctx=ctx)
    codes = bwt.astype(np.int16)
    if codes.size:
        codes[primary] = -1
    return BWT(MaybeSeq(codes, "W"))


def bytestringToBWT(bs, ctx=None) -> BWT:
    """bytestringToBWT = toBWT . BS.unpack   (src/Data/BWT.hs:68-70)"""
    return toBWT(bytes(bs), ctx=ctx)


def textToBWT(t: str, ctx=None) -> TextBWT:
    """textToBWT = TextBWT . bytestringToBWT . encodeUtf8   (src/Data/BWT.hs:79-81)"""
    return TextBWT(bytestringToBWT(t.encode("utf-8"), ctx=ctx))


def fromBWT(bwt: BWT, ctx=None) -> list:
    """fromBWT :: Ord a => BWT a -> [a]   (src/Data/BWT.hs:93-104)."""
    return list(_from_bwt_bytes(bwt, ctx))


def _from_bwt_bytes(bwt: BWT, ctx=None) -> bytes:
    ctx = ctx or default_context()
    codes = bwt.seq.codes
    N = codes.size
    out = np.empty(max(N, 1), dtype=np.uint8)
    n_out = C.c_uint64(0)
    ctx.call("tc_bwt_decode", ptr(codes), N, ptr(out), out.size, C.byref(n_out))
    return out[: n_out.value].tobytes()


def bytestringFromWord8BWT(bwt: BWT, ctx=None) -> bytes:
    """bytestringFromWord8BWT = BS.pack . fromBWT   (src/Data/BWT.hs:108-110)"""
    return _from_bwt_bytes(bwt, ctx)


def bytestringFromByteStringBWT(bwt: BWT, ctx=None) -> bytes:
    """bytestringFromByteStringBWT = BS.concat . fromBWT   (src/Data/BWT.hs:114-116)"""
    return _from_bwt_bytes(bwt, ctx)


def textFromBWT(tb: TextBWT, ctx=None) -> str:
    """textFromBWT = decodeUtf8 . bytestringFromWord8BWT   (src/Data/BWT.hs:120-123)"""
    return _from_bwt_bytes(tb.bwt, ctx).decode("utf-8")
