"""ctypes binding of libtc_b200.so (include/tc_b200.h).

There is NO CPU fallback: if the shared library is missing, or no sm_100-class CUDA
device is usable, every compute call raises.  Nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtc_b200.so")

TC_OK, TC_E_CUDA, TC_E_CAP, TC_E_FROMJUST, TC_E_INDEX, TC_E_NOMEM, TC_E_ARG, TC_E_TOOBIG, TC_E_NODEVICE = (
    0, -1, -2, -3, -4, -5, -6, -7, -8)


class TcError(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__(f"libtc_b200: {msg} (rc={rc})")
        self.rc = rc


class FromJustError(TcError):
    """The reference throws `Maybe.fromJust: Nothing` on this input."""


class SeqIndexError(TcError, IndexError):
    """The reference throws `index out of bounds` (Data.Sequence.index) on this input."""


class NoDeviceError(TcError):
    """No usable B200-class CUDA device; this library has no CPU path."""


class BlockInfo(C.Structure):
    _fields_ = [("n", C.c_uint64), ("N", C.c_uint64), ("primary", C.c_uint64), ("sigma", C.c_uint32),
                ("final_list", C.c_int16 * 257), ("R", C.c_uint64)]


class FmInfo(C.Structure):
    _fields_ = [("n", C.c_uint64), ("N", C.c_uint64), ("primary", C.c_uint64), ("sigma", C.c_uint32),
                ("sa_sample_rate", C.c_uint32), ("alphabet", C.c_int16 * 257), ("C", C.c_int64 * 257),
                ("blob_bytes", C.c_uint64), ("n_samples", C.c_uint64)]


_vp, _u64, _u32, _int = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
_pu64, _pu32 = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)

# name -> (restype, argtypes); mirrors include/tc_b200.h one to one
_SIGS = {
    "tc_ctx_create": (_int, [_int, C.POINTER(_vp)]),
    "tc_ctx_create_on_stream": (_int, [_int, _vp, C.POINTER(_vp)]),
    "tc_ctx_destroy": (None, [_vp]),
    "tc_ctx_sync": (_int, [_vp]),
    "tc_strerror": (C.c_char_p, [_int]),
    "tc_last_error": (C.c_char_p, [_vp]),
    "tc_host_alloc": (_vp, [C.c_size_t]),
    "tc_host_free": (None, [_vp]),
    "tc_ctx_launches": (_u64, [_vp]),
    "tc_version": (C.c_char_p, []),
    "tc_ctx_profile": (_int, [_vp, _int]),
    "tc_ctx_profile_report": (_int, [_vp, C.c_char_p, C.c_size_t]),
    "tc_bwt_encode": (_int, [_vp, _vp, _u64, _vp, _pu64, _vp]),
    "tc_bwt_decode": (_int, [_vp, _vp, _u64, _vp, _u64, _pu64]),
    "tc_bwt_decode_u8": (_int, [_vp, _vp, _u64, _u64, _vp, _u64, _pu64]),
    "tc_mtf_encode": (_int, [_vp, _vp, _u64, _vp, _vp, _pu32]),
    "tc_mtf_encode_u8": (_int, [_vp, _vp, _u64, _u64, _vp, _vp, _pu32]),
    "tc_mtf_decode": (_int, [_vp, _vp, _u64, _vp, _u32, _vp]),
    "tc_rle_encode": (_int, [_vp, _vp, _u64, _vp, _vp, _u64, _pu64]),
    "tc_rle_encode_u8": (_int, [_vp, _vp, _u64, _u64, _vp, _vp, _u64, _pu64]),
    "tc_rle_encode_u16": (_int, [_vp, _vp, _u64, _vp, _vp, _u64, _pu64]),
    "tc_rle_decode": (_int, [_vp, _vp, _vp, _u64, _vp, _u64, _pu64]),
    "tc_bwt_rle_encode": (_int, [_vp, _vp, _u64, _vp, _vp, _u64, C.POINTER(BlockInfo)]),
    "tc_bwt_mtf_rle_encode": (_int, [_vp, _vp, _u64, _vp, _vp, _u64, C.POINTER(BlockInfo)]),
    "tc_blocks_encode": (_int, [_vp, _u64, _vp, _vp, _int, _vp, _vp, _vp, C.POINTER(BlockInfo)]),
    "tc_packed_bound": (_u64, [_u64]),
    "tc_blocks_encode_packed": (_int, [_vp, _u64, _vp, _vp, _int, _vp, _vp, _vp, C.POINTER(BlockInfo)]),
    "tc_packed_info": (_int, [_vp, _u64, C.POINTER(BlockInfo), _pu32]),
    "tc_packed_unpack": (_int, [_vp, _u64, _vp, _vp, _u64, C.POINTER(BlockInfo)]),
    "tc_packed_decode": (_int, [_vp, _vp, _u64, _vp, _u64, _pu64]),
    "tc_blocks_decode_packed": (_int, [_vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    "tc_bwt_rle_decode": (_int, [_vp, _vp, _vp, _u64, _vp, _u64, _pu64]),
    "tc_bwt_mtf_rle_decode": (_int, [_vp, _vp, _vp, C.POINTER(BlockInfo), _vp, _u64, _pu64]),
    "tc_bwt_encode_dev": (_int, [_vp, _vp, _u64, _vp, _pu64, _vp]),
    "tc_mtf_encode_u8_dev": (_int, [_vp, _vp, _u64, _u64, _vp, _vp, _pu32]),
    "tc_rle_encode_u8_dev": (_int, [_vp, _vp, _u64, _u64, _vp, _vp, _u64, _pu64]),
    "tc_rle_encode_u16_dev": (_int, [_vp, _vp, _u64, _vp, _vp, _u64, _pu64]),
    "tc_bwt_mtf_rle_encode_dev": (_int, [_vp, _vp, _u64, _vp, _vp, _u64, C.POINTER(BlockInfo)]),
    "tc_blocks_encode_dev": (_int, [_vp, _u64, _vp, _vp, _int, _vp, _vp, _vp, C.POINTER(BlockInfo)]),
    "tc_fm_build": (_int, [_vp, _vp, _u64, _u32, C.POINTER(_vp)]),
    "tc_fm_build_dev": (_int, [_vp, _vp, _u64, _u32, C.POINTER(_vp)]),
    "tc_fm_free": (None, [_vp]),
    "tc_fm_get_info": (_int, [_vp, C.POINTER(FmInfo)]),
    "tc_fm_count": (_int, [_vp, _vp, _vp, _vp, _u64, _vp]),
    "tc_fm_count_dev": (_int, [_vp, _vp, _vp, _vp, _u64, _vp]),
    "tc_fm_locate": (_int, [_vp, _vp, _vp, _vp, _u64, _vp, _vp, _u64, _pu64]),
    "tc_fm_locate_dev": (_int, [_vp, _vp, _vp, _vp, _u64, _vp, _vp, _u64, _pu64]),
    "tc_fm_export": (_int, [_vp, _vp, _vp, _vp]),
    "tc_fm_blob": (_vp, [_vp]),
    "tc_fm_from_blob_dev": (_int, [_vp, _vp, _u64, _int, C.POINTER(_vp)]),
    "tc_device_count": (_int, []),
    "tc_ctx_pool_acquire": (_int, [_int, C.POINTER(_vp)]),
    "tc_ctx_pool_release": (None, [_vp]),
    "tc_mgpu_blocks_encode_packed": (_int, [_int, _vp, _u64, _vp, _vp, _int, _vp, _vp, _vp, C.POINTER(BlockInfo)]),
    "tc_fm_replicate": (_int, [_vp, _int, _vp, _vp]),
    "tc_mgpu_fm_count": (_int, [_int, _vp, _vp, _vp, _vp, _u64, _vp]),
    "tc_mgpu_fm_locate": (_int, [_int, _vp, _vp, _vp, _vp, _u64, _vp, _vp, _u64, _pu64]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None
_lib_lock = threading.Lock()


def load():
    """Load the shared library (no device needed for this step)."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a). text_compression_b200 has no CPU fallback.")
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGS.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def _raise(ctxh, rc):
    L = load()
    msg = L.tc_strerror(rc).decode()
    if rc == TC_E_CUDA and ctxh:
        msg += ": " + L.tc_last_error(ctxh).decode()
    if rc == TC_E_FROMJUST:
        raise FromJustError(rc, msg)
    if rc == TC_E_INDEX:
        raise SeqIndexError(rc, msg)
    if rc == TC_E_NODEVICE:
        raise NoDeviceError(rc, msg)
    raise TcError(rc, msg)


def ptr(a):
    """void* of a numpy array (host) / int device address / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


class Context:
    """One tc_ctx (device + stream + scratch arena).  Not thread-safe: one per thread."""

    def __init__(self, device: int = 0, stream: int | None = None):
        L = load()
        h = C.c_void_p(None)
        rc = (L.tc_ctx_create(device, C.byref(h)) if stream is None
              else L.tc_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h)))
        if rc != TC_OK:
            _raise(None, rc)
        self.h = h
        self.device = device
        self.L = L

    def call(self, name, *args, allow=()):
        rc = getattr(self.L, name)(self.h, *args)
        if rc != TC_OK and rc not in allow:
            _raise(self.h, rc)
        return rc

    def sync(self):
        self.call("tc_ctx_sync")

    def profile(self, on: bool):
        self.call("tc_ctx_profile", 1 if on else 0)

    def profile_report(self) -> dict:
        """{kernel name: (launches, total_ms, algorithmic_bytes)} since profile(True)."""
        buf = C.create_string_buffer(1 << 16)
        self.call("tc_ctx_profile_report", buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms, nbytes = line.split("\t")
            out[name.strip("()")] = (int(n), float(ms), int(nbytes))
        return out

    @property
    def launches(self) -> int:
        return int(self.L.tc_ctx_launches(self.h))

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.tc_ctx_destroy(self.h)
            self.h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_tls = threading.local()


def default_context() -> Context:
    """Thread-local default context on the current device (LOCAL_RANK under torchrun)."""
    ctx = getattr(_tls, "ctx", None)
    if ctx is None:
        dev = int(os.environ.get("TC_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        ctx = Context(dev)
        _tls.ctx = ctx
    return ctx


def set_default_context(ctx: Context | None):
    _tls.ctx = ctx


def pinned_empty(count: int, dtype) -> np.ndarray:
    """numpy array backed by CUDA-pinned memory (tc_host_alloc); freed with the array."""
    L = load()
    dt = np.dtype(dtype)
    nbytes = max(1, count * dt.itemsize)
    p = L.tc_host_alloc(nbytes)
    if not p:
        raise MemoryError("tc_host_alloc failed")

    class _Buf(C.c_char * nbytes):  # ctypes arrays accept attributes, so the owner rides along
        def __del__(self):
            try:
                L.tc_host_free(C.c_void_p(C.addressof(self)))
            except Exception:
                pass

    buf = _Buf.from_address(p)
    return np.frombuffer(buf, dtype=dt, count=count)
