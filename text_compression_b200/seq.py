"""Host-side stand-ins for the reference's value types.

The reference works on `Data.Sequence.Seq (Maybe b)` with `Nothing` as the `$` sentinel
(src/Data/BWT/Internal.hs:83).  Here a `Seq (Maybe b)` over byte-derived symbols is a
`MaybeSeq`: a numpy int16 array (-1 == Nothing) plus the element kind:
  'W'  Word8       (BWT Word8)
  'B'  ByteString  (one-byte strings, `BS.singleton`)
  'T'  Text        (one-character texts, `decodeUtf8 . BS.singleton`: ASCII only, Q6)
`to_list()` materialises the Haskell value shape (None / int / bytes / str).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def _check_text_kind(codes: np.ndarray):
    # decodeUtf8 . BS.singleton throws on any byte >= 0x80 (src/Data/MTF.hs:131,142)
    if codes.size and int(codes.max(initial=-1)) >= 0x80:
        bad = int(codes[codes >= 0x80][0])
        raise UnicodeDecodeError("utf-8", bytes([bad]), 0, 1, "invalid start byte (per-byte decodeUtf8 in the reference)")


class MaybeSeq:
    __slots__ = ("codes", "kind")

    def __init__(self, codes, kind: str = "B"):
        self.codes = np.ascontiguousarray(codes, dtype=np.int16)
        assert kind in ("W", "B", "T")
        if kind == "T":
            _check_text_kind(self.codes)
        self.kind = kind

    @classmethod
    def from_list(cls, items, kind: str = "B") -> "MaybeSeq":
        out = np.empty(len(items), dtype=np.int16)
        for i, x in enumerate(items):
            if x is None:
                out[i] = -1
            elif isinstance(x, (int, np.integer)):
                out[i] = int(x)
            elif isinstance(x, (bytes, bytearray)):
                if len(x) != 1:
                    raise NotImplementedError("multi-byte Pack items are out of scope for the GPU path")
                out[i] = x[0]
            else:
                b = str(x).encode("utf-8")
                if len(b) != 1:
                    raise NotImplementedError("multi-byte Pack items are out of scope for the GPU path")
                out[i] = b[0]
        return cls(out, kind)

    def as_kind(self, kind: str) -> "MaybeSeq":
        return MaybeSeq(self.codes, kind)

    def to_list(self):
        if self.kind == "W":
            return [None if c < 0 else int(c) for c in self.codes.tolist()]
        if self.kind == "B":
            return [None if c < 0 else bytes([c]) for c in self.codes.tolist()]
        return [None if c < 0 else chr(c) for c in self.codes.tolist()]

    def __len__(self):
        return int(self.codes.size)

    def __eq__(self, other):
        return isinstance(other, MaybeSeq) and self.kind == other.kind and np.array_equal(self.codes, other.codes)

    def __repr__(self):
        return f"MaybeSeq<{self.kind}>({self.to_list()!r})"


@dataclass(eq=False)
class BWT:
    """newtype BWT a = BWT (Seq (Maybe a))  (src/Data/BWT/Internal.hs:83)"""
    seq: MaybeSeq

    def __eq__(self, o):
        return isinstance(o, BWT) and self.seq == o.seq

    def __len__(self):
        return len(self.seq)


@dataclass(eq=False)
class TextBWT:
    """newtype TextBWT = TextBWT (BWT Word8)  (src/Data/BWT.hs:74)"""
    bwt: BWT

    def __eq__(self, o):
        return isinstance(o, TextBWT) and self.bwt == o.bwt


@dataclass(eq=False)
class MTF:
    """newtype MTF b = MTF (Seq Int, Seq (Maybe b))  (src/Data/MTF/Internal.hs:67)"""
    indices: np.ndarray      # int64, like Haskell Int
    final_list: MaybeSeq     # seqToMTF returns the FINAL list

    def __eq__(self, o):
        return (isinstance(o, MTF) and np.array_equal(self.indices, o.indices) and self.final_list == o.final_list)

    def to_tuple(self):
        return (self.indices.tolist(), self.final_list.to_list())


def _show(kind: str, n: int):
    s = str(n)
    return s.encode() if kind == "B" else s


@dataclass(eq=False)
class RLE:
    """newtype RLE b = RLE (Seq (Maybe b))  (src/Data/RLE/Internal.hs:95): the flat sequence
    [Just (show count), symbol, ...].  Stored as parallel arrays; `to_list()` renders it."""
    counts: np.ndarray       # uint32 run lengths
    syms: np.ndarray         # int16, -1 == Nothing
    kind: str = "B"

    def __post_init__(self):
        self.counts = np.ascontiguousarray(self.counts, dtype=np.uint32)
        self.syms = np.ascontiguousarray(self.syms, dtype=np.int16)
        if self.kind == "T":
            _check_text_kind(self.syms)

    def to_list(self):
        out = []
        one = (lambda c: bytes([c])) if self.kind == "B" else chr
        for c, s in zip(self.counts.tolist(), self.syms.tolist()):
            out.append(_show(self.kind, c))
            out.append(None if s < 0 else one(s))
        return out

    @classmethod
    def from_list(cls, items, kind: str = "B") -> "RLE":
        """Parse a flat reference-style RLE value.  An odd trailing element is ignored
        (src/Data/RLE/Internal.hs:187-189); `Nothing` in a count slot is where the reference
        throws fromJust, reported here eagerly."""
        from ._lib import FromJustError, TC_E_FROMJUST
        cnt, sym = [], []
        for k in range(0, len(items) - 1, 2):
            y1, y2 = items[k], items[k + 1]
            if y1 is None:
                raise FromJustError(TC_E_FROMJUST, "Nothing in a count slot of an RLE value")
            c = int(y1.decode() if isinstance(y1, (bytes, bytearray)) else y1)
            cnt.append(max(c, 0))  # replicateM_ with n <= 0 pushes nothing
            if y2 is None:
                sym.append(-1)
            else:
                b = y2 if isinstance(y2, (bytes, bytearray)) else str(y2).encode("utf-8")
                if len(b) != 1:
                    raise NotImplementedError("multi-byte Pack items are out of scope for the GPU path")
                sym.append(b[0])
        return cls(np.array(cnt, dtype=np.uint32), np.array(sym, dtype=np.int16), kind)

    def __len__(self):
        return 2 * int(self.counts.size)

    def __eq__(self, o):
        return (isinstance(o, RLE) and self.kind == o.kind and np.array_equal(self.counts, o.counts)
                and np.array_equal(self.syms, o.syms))


def to_bytes(x) -> bytes:
    if isinstance(x, str):
        return x.encode("utf-8")            # DTE.encodeUtf8
    return bytes(x)
