"""text_compression_b200 -- the B200 (sm_100a) implementation of the text-compression hot path
(Data.BWT, Data.MTF, Data.RLE, Data.FMIndex) behind the reference's own API names.

The compute lives in libtc_b200.so (hand-written CUDA, C ABI in include/tc_b200.h); this
package is the host-side mirror of the reference interface.  There is no CPU fallback.
"""
from . import _lib
from ._lib import Context, FromJustError, NoDeviceError, SeqIndexError, TcError, default_context
from .seq import BWT, MTF, RLE, MaybeSeq, TextBWT

__version__ = "0.1.0"
__all__ = ["Context", "default_context", "TcError", "FromJustError", "SeqIndexError", "NoDeviceError",
           "BWT", "MTF", "RLE", "MaybeSeq", "TextBWT", "bwt", "mtf", "rle", "fmindex", "block", "stream", "generic", "multi"]


def __getattr__(name):
    if name in ("bwt", "mtf", "rle", "fmindex", "block", "multi", "stream", "generic"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
