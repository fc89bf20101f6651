"""ctypes binding of the CPU oracle (oracle/tc_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(text_compression_b200) never imports this module.

Representation: maybe-symbols are numpy int16 with -1 == Nothing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborc.so")

ORC_OK, ORC_E_CAP, ORC_E_FROMJUST, ORC_E_INDEX, ORC_E_NOMEM = 0, -2, -3, -4, -5


class OracleError(RuntimeError):
    """Mirrors a Haskell exception of the reference (fromJust / DS.index)."""

    def __init__(self, rc):
        super().__init__({-3: "fromJust Nothing", -4: "index out of bounds", -2: "capacity", -5: "no memory"}.get(rc, str(rc)))
        self.rc = rc


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, i16p, i32p, i64p, u32p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_int16, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64))
        vp = C.c_void_p
        L.orc_suffix_array.argtypes = [vp, C.c_uint64, vp]
        L.orc_bwt_encode.argtypes = [vp, C.c_uint64, vp, u64p, vp]
        L.orc_bwt_decode.argtypes = [vp, C.c_uint64, vp, C.c_uint64, u64p]
        L.orc_nub_sorted.argtypes = [vp, C.c_uint64, vp]
        L.orc_nub_sorted.restype = C.c_uint32
        L.orc_mtf_encode.argtypes = [vp, C.c_uint64, vp, vp, u32p]
        L.orc_mtf_decode.argtypes = [vp, C.c_uint64, vp, C.c_uint32, vp, u64p]
        L.orc_rle_encode.argtypes = [vp, C.c_uint64, vp, vp, C.c_uint64, u64p]
        L.orc_rle_decode.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_uint64, u64p]
        L.orc_fm_build.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
        L.orc_fm_free.argtypes = [vp]
        L.orc_fm_count.argtypes = [vp, vp, C.c_uint64]
        L.orc_fm_count.restype = C.c_int64
        L.orc_fm_locate.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64]
        L.orc_fm_locate.restype = C.c_int64
        L.orc_fm_count_batch.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_int]
        for name, rt in (("orc_fm_N", C.c_uint64), ("orc_fm_sigma", C.c_uint32), ("orc_fm_alpha", vp), ("orc_fm_C", vp),
                         ("orc_fm_occ", vp), ("orc_fm_bwt", vp), ("orc_fm_sa", vp)):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = rt
        L.orc_fms_build.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
        L.orc_fms_free.argtypes = [vp]
        L.orc_fms_count.argtypes = [vp, vp, C.c_uint64]
        L.orc_fms_count.restype = C.c_int64
        L.orc_fms_count_batch.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_int]
        L.orc_naive_search.argtypes = [vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, vp, vp, vp, C.c_uint64, u64p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u8(x) -> np.ndarray:
    if isinstance(x, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


def _check(rc):
    if rc != 0:
        raise OracleError(rc)


# ---- BWT ------------------------------------------------------------------
def suffix_array(text) -> np.ndarray:
    """createSuffixArray: 1-based start positions in rank order (n+1 entries)."""
    t = _u8(text)
    sa = np.empty(t.size + 1, dtype=np.uint32)
    lib().orc_suffix_array(_p(t), t.size, _p(sa))
    return sa


def bwt_encode(text, want_sa: bool = False):
    """toBWT: int16 array of n+1 maybe-symbols (empty for empty input)."""
    t = _u8(text)
    N = C.c_uint64(0)
    bwt = np.empty(t.size + 1, dtype=np.int16)
    sa = np.empty(t.size + 1, dtype=np.uint32) if want_sa else None
    _check(lib().orc_bwt_encode(_p(t), t.size, _p(bwt), C.byref(N), _p(sa)))
    bwt = bwt[: N.value]
    return (bwt, sa[: N.value]) if want_sa else bwt


def bwt_decode(bwt) -> np.ndarray:
    """fromBWT on int16 maybe-symbols -> uint8 text."""
    b = np.ascontiguousarray(bwt, dtype=np.int16)
    out = np.empty(max(b.size, 1), dtype=np.uint8)
    n = C.c_uint64(0)
    _check(lib().orc_bwt_decode(_p(b), b.size, _p(out), out.size, C.byref(n)))
    return out[: n.value].copy()


# ---- MTF ------------------------------------------------------------------
def nub_sorted(sym) -> np.ndarray:
    s = np.ascontiguousarray(sym, dtype=np.int16)
    lst = np.empty(257, dtype=np.int16)
    k = lib().orc_nub_sorted(_p(s), s.size, _p(lst))
    return lst[:k].copy()


def mtf_encode(sym):
    """seqToMTF -> (indices int32[N], final list int16[sigma])."""
    s = np.ascontiguousarray(sym, dtype=np.int16)
    idx = np.empty(s.size, dtype=np.int32)
    fin = np.empty(257, dtype=np.int16)
    sg = C.c_uint32(0)
    _check(lib().orc_mtf_encode(_p(s), s.size, _p(idx), _p(fin), C.byref(sg)))
    return idx, fin[: sg.value].copy()


def mtf_decode(idx, final_list) -> np.ndarray:
    i = np.ascontiguousarray(idx, dtype=np.int32)
    f = np.ascontiguousarray(final_list, dtype=np.int16)
    out = np.empty(i.size, dtype=np.int16)
    n = C.c_uint64(0)
    _check(lib().orc_mtf_decode(_p(i), i.size, _p(f), f.size, _p(out), C.byref(n)))
    return out[: n.value].copy()


# ---- RLE ------------------------------------------------------------------
def rle_encode(sym):
    """seqToRLE -> (counts int64[R], symbols int16[R]); flat Seq = [show c, s]..."""
    s = np.ascontiguousarray(sym, dtype=np.int16)
    cap = 2 * s.size + 2
    cnt = np.empty(cap, dtype=np.int64)
    rs = np.empty(cap, dtype=np.int16)
    R = C.c_uint64(0)
    _check(lib().orc_rle_encode(_p(s), s.size, _p(cnt), _p(rs), cap, C.byref(R)))
    return cnt[: R.value].copy(), rs[: R.value].copy()


def rle_decode(count, rsym, has_count=None) -> np.ndarray:
    c = np.ascontiguousarray(count, dtype=np.int64)
    r = np.ascontiguousarray(rsym, dtype=np.int16)
    hc = None if has_count is None else np.ascontiguousarray(has_count, dtype=np.uint8)
    cap = int(np.maximum(c, 1).sum()) + 1
    out = np.empty(cap, dtype=np.int16)
    n = C.c_uint64(0)
    _check(lib().orc_rle_decode(_p(c), _p(r), _p(hc), c.size, _p(out), cap, C.byref(n)))
    return out[: n.value].copy()


# ---- FM-index (dense, as the reference stores it) ---------------------------
class FMIndex:
    def __init__(self, text):
        t = _u8(text)
        self._h = C.c_void_p(None)
        _check(lib().orc_fm_build(_p(t), t.size, C.byref(self._h)))
        self.n = t.size
        if self._h.value:
            L = lib()
            self.N = L.orc_fm_N(self._h)
            self.sigma = L.orc_fm_sigma(self._h)

    def _arr(self, fn, dtype, count):
        ptr = getattr(lib(), fn)(self._h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(count,)).copy()

    @property
    def alphabet(self):
        return self._arr("orc_fm_alpha", np.int16, self.sigma)

    @property
    def Cc(self):
        return self._arr("orc_fm_C", np.int64, self.sigma)

    @property
    def occ(self):
        return self._arr("orc_fm_occ", np.uint32, self.sigma * self.N).reshape(self.sigma, self.N)

    @property
    def bwt(self):
        return self._arr("orc_fm_bwt", np.int16, self.N)

    @property
    def sa(self):
        return self._arr("orc_fm_sa", np.uint32, self.N)

    def count(self, pat) -> int:
        p = _u8(pat)
        return lib().orc_fm_count(self._h, _p(p), p.size)

    def locate(self, pat) -> np.ndarray:
        p = _u8(pat)
        k = lib().orc_fm_locate(self._h, _p(p), p.size, None, 0)
        pos = np.empty(max(k, 1), dtype=np.int64)
        lib().orc_fm_locate(self._h, _p(p), p.size, _p(pos), pos.size)
        return pos[:k].copy()

    def count_batch(self, pats: np.ndarray, off: np.ndarray, nthreads: int = 1) -> np.ndarray:
        pats = _u8(pats)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        q = off.size - 1
        out = np.empty(q, dtype=np.int64)
        lib().orc_fm_count_batch(self._h, _p(pats), _p(off), q, _p(out), nthreads)
        return out

    def __del__(self):
        try:
            if self._h.value:
                lib().orc_fm_free(self._h)
                self._h = C.c_void_p(None)
        except Exception:
            pass


class FMIndexSampled:
    """Checkpointed CPU FM-index for baseline timing at sizes where the dense
    sigma x N table does not fit.  Same search as FMIndex.count."""

    def __init__(self, text):
        t = _u8(text)
        self._h = C.c_void_p(None)
        _check(lib().orc_fms_build(_p(t), t.size, C.byref(self._h)))

    def count(self, pat) -> int:
        p = _u8(pat)
        return lib().orc_fms_count(self._h, _p(p), p.size)

    def count_batch(self, pats, off, nthreads: int = 1) -> np.ndarray:
        pats = _u8(pats)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        q = off.size - 1
        out = np.empty(q, dtype=np.int64)
        lib().orc_fms_count_batch(self._h, _p(pats), _p(off), q, _p(out), nthreads)
        return out

    def __del__(self):
        try:
            if self._h.value:
                lib().orc_fms_free(self._h)
                self._h = C.c_void_p(None)
        except Exception:
            pass


# ---- independent occurrence scan (full-size FM-index checks) -------------
def naive_search(text, pats: np.ndarray, want_pos: bool = False):
    """Brute-force occurrences of q equal-length patterns (rows of `pats`) in `text`:
    counts int64[q] and, with want_pos, (hit_off uint64[q+1], 1-based positions ascending per pattern).
    Not a restatement of the reference: the definition count / locate must agree with (patterns over
    symbols that occur in the text)."""
    t = _u8(text)
    p = np.ascontiguousarray(pats, dtype=np.uint8)
    q, m = p.shape
    cnt = np.empty(q, dtype=np.int64)
    tot = C.c_uint64(0)
    if not want_pos:
        _check(lib().orc_naive_search(_p(t), t.size, _p(p), q, m, _p(cnt), None, None, 0, C.byref(tot)))
        return cnt
    ho = np.empty(q + 1, dtype=np.uint64)
    cap = 1 << 20
    while True:
        pos = np.empty(cap, dtype=np.uint64)
        rc = lib().orc_naive_search(_p(t), t.size, _p(p), q, m, _p(cnt), _p(ho), _p(pos), cap, C.byref(tot))
        if rc == ORC_E_CAP:
            cap = int(tot.value) + 1
            continue
        _check(rc)
        return cnt, ho, pos[: tot.value].copy()


# ---- Seq-level renderings used by the golden tests -----------------------
def seq_from_rle_pairs(cnt, rs, item=lambda b: bytes([b])):
    """Flat `Seq (Maybe b)` of the reference: [Just (show count), sym, ...]."""
    out = []
    for c, s in zip(cnt.tolist(), rs.tolist()):
        out.append(str(c).encode())
        out.append(None if s < 0 else item(s))
    return out


def rle_pairs_from_seq(seq):
    """Inverse of seq_from_rle_pairs for well-formed RLE values."""
    cnt, rs = [], []
    for k in range(0, len(seq) - 1, 2):
        cnt.append(int(seq[k]))
        rs.append(-1 if seq[k + 1] is None else seq[k + 1][0])
    return np.array(cnt, dtype=np.int64), np.array(rs, dtype=np.int16)
