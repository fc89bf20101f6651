"""Data.FMIndex on the B200 (mirror of src/Data/FMIndex.hs + src/Data/FMIndex/Internal.hs).

The reference stores a dense sigma x N Occ table and the full suffix array; here the index
lives in HBM as rank-block bit-planes + a sampled SA (tc_fm_build).  The public value shape
(Cc, OccCK, SA) can still be materialised for small inputs (`FMIndex.Cc/.OccCK/.SA`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import TC_E_CAP, FmInfo, default_context, ptr
from .bwt import bytestringFromWord8BWT, textFromBWT
from .seq import BWT, MaybeSeq, TextBWT, to_bytes

__all__ = [
    "FMIndex", "seqToOccCK", "seqToCc", "seqFromFMIndex", "countFMIndex", "locateFMIndex",
    "bytestringToBWTToFMIndexB", "bytestringToBWTToFMIndexT", "textToBWTToFMIndexB", "textToBWTToFMIndexT",
    "textBWTToFMIndexB", "bytestringBWTToFMIndexB", "textBWTToFMIndexT", "bytestringBWTToFMIndexT",
    "bytestringFromBWTFromFMIndexB", "bytestringFromBWTFromFMIndexT", "textFromBWTFromFMIndexB",
    "textFromBWTFromFMIndexT", "textBWTFromFMIndexT", "bytestringBWTFromFMIndexT", "textBWTFromFMIndexB",
    "bytestringBWTFromFMIndexB", "textFromFMIndexB", "bytestringFromFMIndexB", "textFromFMIndexT",
    "bytestringFromFMIndexT",
    "bytestringFMIndexCountS", "textFMIndexCountS", "bytestringFMIndexCountP", "textFMIndexCountP",
    "bytestringFMIndexLocateS", "textFMIndexLocateS", "bytestringFMIndexLocateP", "textFMIndexLocateP",
]


def pack_patterns(pats):
    """[pattern] -> (uint8 bytes, uint64 offsets[q+1])."""
    bs = [to_bytes(p) for p in pats]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    flat = np.frombuffer(b"".join(bs), dtype=np.uint8) if bs else np.empty(0, dtype=np.uint8)
    return np.ascontiguousarray(flat), off


class FMIndex:
    """newtype FMIndex b = FMIndex (Cc b, OccCK b, SA b)  (src/Data/FMIndex/Internal.hs:153).

    `sa_sample_rate=1` keeps the full suffix array (what the reference stores); larger rates
    trade locate time for memory.  An empty index mirrors
    FMIndex (Cc Empty, OccCK Empty, SA Empty) (src/Data/FMIndex.hs:139,165).
    """

    def __init__(self, text=None, kind: str = "B", sa_sample_rate: int = 1, ctx=None, _handle=None):
        self.ctx = ctx or default_context()
        self.kind = kind
        self.h = C.c_void_p(None)
        self.info = None
        if _handle is not None:
            self.h = _handle
        elif text is not None:
            t = np.frombuffer(to_bytes(text), dtype=np.uint8) if not isinstance(text, np.ndarray) else text
            if t.size:
                self.ctx.call("tc_fm_build", ptr(np.ascontiguousarray(t)), t.size, sa_sample_rate, C.byref(self.h))
        if self.h.value:
            self.info = FmInfo()
            self.ctx.L.tc_fm_get_info(self.h, C.byref(self.info))

    # -- lifecycle
    @property
    def empty(self) -> bool:
        return not self.h.value

    def close(self):
        if self.h.value:
            self.ctx.L.tc_fm_free(self.h)
            self.h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- batched queries (device)
    def count_many(self, pats) -> np.ndarray:
        """countFMIndex for each pattern: int64, -1 == Nothing."""
        flat, off = pack_patterns(pats)
        q = off.size - 1
        out = np.full(q, -1, dtype=np.int64)
        if self.empty or q == 0:
            return out   # countFMIndex _ (FMIndex (Cc Empty,_,_)) = Nothing (:349-351)
        self.ctx.call("tc_fm_count", self.h, ptr(flat), ptr(off), q, ptr(out))
        return out

    def locate_many(self, pats):
        """locateFMIndex + rank->position map: (hit_off uint64[q+1], pos_1based uint64[total])."""
        flat, off = pack_patterns(pats)
        q = off.size - 1
        hit_off = np.zeros(q + 1, dtype=np.uint64)
        if self.empty or q == 0:
            return hit_off, np.empty(0, dtype=np.uint64)
        total = C.c_uint64(0)
        self.ctx.call("tc_fm_locate", self.h, ptr(flat), ptr(off), q, ptr(hit_off), None, 0, C.byref(total),
                      allow=(TC_E_CAP,))
        pos = np.empty(int(total.value), dtype=np.uint64)
        if pos.size:
            self.ctx.call("tc_fm_locate", self.h, ptr(flat), ptr(off), q, ptr(hit_off), ptr(pos), pos.size,
                          C.byref(total))
        return hit_off, pos

    # -- the reference's value shape, materialised on the host (small N only)
    def _export(self, want_sa: bool):
        N = int(self.info.N)
        bwt = np.empty(N, dtype=np.int16)
        sa = np.empty(N, dtype=np.uint32) if want_sa else None
        self.ctx.call("tc_fm_export", self.h, ptr(bwt), ptr(sa))
        return bwt, sa

    @property
    def alphabet(self) -> np.ndarray:
        return np.array(self.info.alphabet[: self.info.sigma], dtype=np.int16)

    @property
    def bwt(self) -> MaybeSeq:
        if self.empty:
            return MaybeSeq(np.empty(0, dtype=np.int16), self.kind)
        return MaybeSeq(self._export(False)[0], self.kind)

    @property
    def Cc(self):
        """Cc (Seq (Int, Maybe b))  (src/Data/FMIndex/Internal.hs:161)"""
        if self.empty:
            return []
        syms = MaybeSeq(self.alphabet, self.kind).to_list()
        return [(int(self.info.C[j]), syms[j]) for j in range(self.info.sigma)]

    @property
    def OccCK(self):
        """OccCK (Seq (Maybe b, Seq (Int, Int, Maybe b)))  (:157): dense, k = 1..N inclusive."""
        if self.empty:
            return []
        bwt = self._export(False)[0]
        syms = MaybeSeq(self.alphabet, self.kind).to_list()
        col = MaybeSeq(bwt, self.kind).to_list()
        rows = []
        for j, a in enumerate(self.alphabet.tolist()):
            occ = np.cumsum(bwt == a)
            rows.append((syms[j], [(k + 1, int(occ[k]), col[k]) for k in range(bwt.size)]))
        return rows

    @property
    def SA(self):
        """SA (SuffixArray b)  (:169) as (suffixindex, suffixstartpos) pairs."""
        if self.empty:
            return []
        sa = self._export(True)[1]
        return [(k + 1, int(s)) for k, s in enumerate(sa.tolist())]

    def __eq__(self, o):
        return isinstance(o, FMIndex) and self.kind == o.kind and self.Cc == o.Cc and self.bwt == o.bwt


# ---- Internal-level functions (src/Data/FMIndex/Internal.hs) ---------------------------------
def seqToOccCK(xs: MaybeSeq):
    """seqToOccCK (:195-259) on a BWT sequence -- dense host table (small N; the device index
    answers Occ from rank blocks instead)."""
    alpha = np.unique(xs.codes)
    syms = MaybeSeq(alpha, xs.kind).to_list()
    col = xs.to_list()
    return [(syms[j], [(k + 1, int(o), col[k]) for k, o in enumerate(np.cumsum(xs.codes == a).tolist())])
            for j, a in enumerate(alpha.tolist())]


def seqToCc(xs: MaybeSeq):
    """seqToCc (:275-316) on the F column: first-occurrence index of each alphabet symbol."""
    alpha = np.unique(xs.codes)
    syms = MaybeSeq(alpha, xs.kind).to_list()
    return [(int(np.argmax(xs.codes == a)), syms[j]) for j, a in enumerate(alpha.tolist())]


def seqFromFMIndex(fm: FMIndex) -> MaybeSeq:
    """seqFromFMIndex (:324-336): the BWT column stored in the first Occ row."""
    return fm.bwt


def _pattern_bytes(pat) -> bytes:
    if isinstance(pat, MaybeSeq):
        return bytes(pat.codes.astype(np.uint8).tolist())
    if isinstance(pat, (list, tuple)):
        return b"".join(to_bytes(x) for x in pat)
    return to_bytes(pat)


def countFMIndex(pat, fm: FMIndex):
    """countFMIndex :: Seq b -> FMIndex b -> Maybe Int  (:347-438)"""
    c = int(fm.count_many([_pattern_bytes(pat)])[0])
    return None if c < 0 else c


def locateFMIndex(pat, fm: FMIndex):
    """locateFMIndex (:448-542) composed with the wrappers' rank->position map
    (src/Data/FMIndex.hs:496): 1-based text positions in SA-rank order."""
    _, pos = fm.locate_many([_pattern_bytes(pat)])
    return [int(p) for p in pos.tolist()]


# ---- builders (src/Data/FMIndex.hs:108-235) ------------------------------------------------
def bytestringToBWTToFMIndexB(xs, sa_sample_rate=1, ctx=None): return FMIndex(bytes(xs), "B", sa_sample_rate, ctx)
def bytestringToBWTToFMIndexT(xs, sa_sample_rate=1, ctx=None): return FMIndex(bytes(xs), "T", sa_sample_rate, ctx)
def textToBWTToFMIndexB(xs, sa_sample_rate=1, ctx=None): return FMIndex(xs.encode("utf-8"), "B", sa_sample_rate, ctx)
def textToBWTToFMIndexT(xs, sa_sample_rate=1, ctx=None): return FMIndex(xs.encode("utf-8"), "T", sa_sample_rate, ctx)


def _from_bwt(bwm, xs, kind, sa_sample_rate, ctx):
    # BWTMatrix Empty -> FMIndex (Cc Empty, OccCK Empty, SA Empty) (src/Data/FMIndex.hs:139,165).
    # The matrix is only used for its first column (= sorted BWT symbols); the text the SA is
    # built over is recovered with fromBWT exactly as the reference does (:143-147,169-173).
    if bwm is not None and len(bwm) == 0:
        return FMIndex(None, kind, sa_sample_rate, ctx)
    text = textFromBWT(xs, ctx).encode("utf-8") if isinstance(xs, TextBWT) else bytestringFromWord8BWT(xs, ctx)
    return FMIndex(text, kind, sa_sample_rate, ctx)


def textBWTToFMIndexB(bwm, xs, sa_sample_rate=1, ctx=None): return _from_bwt(bwm, xs, "B", sa_sample_rate, ctx)
def bytestringBWTToFMIndexB(bwm, xs, sa_sample_rate=1, ctx=None): return _from_bwt(bwm, xs, "B", sa_sample_rate, ctx)
def textBWTToFMIndexT(bwm, xs, sa_sample_rate=1, ctx=None): return _from_bwt(bwm, xs, "T", sa_sample_rate, ctx)
def bytestringBWTToFMIndexT(bwm, xs, sa_sample_rate=1, ctx=None): return _from_bwt(bwm, xs, "T", sa_sample_rate, ctx)


# ---- extractors (src/Data/FMIndex.hs:244-351) -----------------------------------------------
def _bwt_of(fm, kind) -> BWT: return BWT(seqFromFMIndex(fm).as_kind(kind))
def textBWTFromFMIndexT(fm): return _bwt_of(fm, "T")
def bytestringBWTFromFMIndexT(fm): return _bwt_of(fm, "B")
def textBWTFromFMIndexB(fm): return _bwt_of(fm, "T")
def bytestringBWTFromFMIndexB(fm): return _bwt_of(fm, "B")


def _text_of(fm, ctx=None) -> bytes:
    from .bwt import bytestringFromByteStringBWT
    return bytestringFromByteStringBWT(_bwt_of(fm, "B"), ctx or fm.ctx)


def bytestringFromBWTFromFMIndexB(fm): return _text_of(fm)
def bytestringFromBWTFromFMIndexT(fm): return _text_of(fm)
def textFromBWTFromFMIndexB(fm): return _text_of(fm).decode("utf-8")
def textFromBWTFromFMIndexT(fm): return _text_of(fm).decode("utf-8")
def textFromFMIndexB(fm): return seqFromFMIndex(fm).as_kind("T")
def bytestringFromFMIndexB(fm): return seqFromFMIndex(fm).as_kind("B")
def textFromFMIndexT(fm): return seqFromFMIndex(fm).as_kind("T")
def bytestringFromFMIndexT(fm): return seqFromFMIndex(fm).as_kind("B")


# ---- batch count / locate (src/Data/FMIndex.hs:362-462,475-599) --------------------------------
# Like the reference, every call rebuilds the index from `input` (:368,391,419,449,...), echoes
# the patterns and keeps input order.  The ...P variants are IO in the reference only to read
# getNumCapabilities for parListChunk; here the query batch is already spread over the whole
# GPU (one query per thread), and over several GPUs by text_compression_b200.multi.
def _count(allpats, input_, kind, rate, ctx):
    allpats = list(allpats)
    data = to_bytes(input_)
    if not allpats or not data:
        return []
    fm = FMIndex(data, kind, rate, ctx)
    try:
        c = fm.count_many(allpats)
    finally:
        fm.close()
    return [(p, None if v < 0 else int(v)) for p, v in zip(allpats, c.tolist())]


def _locate(allpats, input_, kind, rate, ctx):
    allpats = list(allpats)
    data = to_bytes(input_)
    if not allpats or not data:
        return []
    fm = FMIndex(data, kind, rate, ctx)
    try:
        hit_off, pos = fm.locate_many(allpats)
    finally:
        fm.close()
    ho = hit_off.tolist()
    return [(p, pos[ho[i]:ho[i + 1]].astype(np.int64).tolist()) for i, p in enumerate(allpats)]


def bytestringFMIndexCountS(allpats, input_, sa_sample_rate=32, ctx=None): return _count(allpats, input_, "B", sa_sample_rate, ctx)
def textFMIndexCountS(allpats, input_, sa_sample_rate=32, ctx=None): return _count(allpats, input_, "T", sa_sample_rate, ctx)
def bytestringFMIndexCountP(allpats, input_, sa_sample_rate=32, ctx=None): return _count(allpats, input_, "B", sa_sample_rate, ctx)
def textFMIndexCountP(allpats, input_, sa_sample_rate=32, ctx=None): return _count(allpats, input_, "T", sa_sample_rate, ctx)
def bytestringFMIndexLocateS(allpats, input_, sa_sample_rate=32, ctx=None): return _locate(allpats, input_, "B", sa_sample_rate, ctx)
def textFMIndexLocateS(allpats, input_, sa_sample_rate=32, ctx=None): return _locate(allpats, input_, "T", sa_sample_rate, ctx)
def bytestringFMIndexLocateP(allpats, input_, sa_sample_rate=32, ctx=None): return _locate(allpats, input_, "B", sa_sample_rate, ctx)
def textFMIndexLocateP(allpats, input_, sa_sample_rate=32, ctx=None): return _locate(allpats, input_, "T", sa_sample_rate, ctx)
