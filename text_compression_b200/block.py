"""Composed helpers with device-resident chaining.

  compress_bwt_rle      = bytestringToBWTToRLEB            (src/Data/RLE.hs:83-85)
  compress_bwt_mtf_rle  = bytestringToBWTToMTFB, then seqToRLE over the index stream
                          (the "BWT+MTF+RLE" composite of BASELINE.json)
BWT -> MTF -> RLE never leaves HBM; only the text goes in and the runs come out.

Packed block container (layout in include/tc_b200.h): compress_blocks_packed
returns each block as one byte string (header + runs at 2 bytes + 1 bit each); unpack_block turns
it back into the run records on the host, decompress_packed into the text on the device.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import TC_E_CAP, TC_OK, BlockInfo, _raise, default_context, load, ptr
from .seq import to_bytes


@dataclass
class CompressedBlock:
    n: int
    N: int
    primary: int
    sigma: int
    final_list: np.ndarray   # int16[sigma]; empty when MTF was not applied
    counts: np.ndarray       # uint32[R]
    syms: np.ndarray         # int16[R]: BWT symbols (-1 == Nothing) or MTF indices
    with_mtf: bool

    @property
    def R(self) -> int:
        return int(self.counts.size)


def _compress(fn, text, ctx, with_mtf) -> CompressedBlock:
    ctx = ctx or default_context()
    t = text if isinstance(text, np.ndarray) else np.frombuffer(to_bytes(text), dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    n = t.size
    cap = n + 3
    cnt = np.empty(cap, dtype=np.uint32)
    sym = np.empty(cap, dtype=np.int16)
    info = BlockInfo()
    ctx.call(fn, ptr(t), n, ptr(cnt), ptr(sym), cap, C.byref(info))
    R = int(info.R)
    fin = np.array(info.final_list[: info.sigma], dtype=np.int16)
    return CompressedBlock(n, int(info.N), int(info.primary), int(info.sigma), fin, cnt[:R].copy(), sym[:R].copy(),
                           with_mtf)


def compress_bwt_rle(text, ctx=None) -> CompressedBlock:
    return _compress("tc_bwt_rle_encode", text, ctx, False)


def compress_bwt_mtf_rle(text, ctx=None) -> CompressedBlock:
    return _compress("tc_bwt_mtf_rle_encode", text, ctx, True)


def compress_blocks(texts, with_mtf: bool = True, ctx=None, pinned: bool = True) -> list:
    """Multi-block compression (BASELINE.json config 5) through tc_blocks_encode: the copy of
    block b+1 to the device and of block b-1's runs back overlap the compression of block b.
    Returns one CompressedBlock per input, in input order."""
    from ._lib import pinned_empty
    ctx = ctx or default_context()
    ts = [np.ascontiguousarray(t if isinstance(t, np.ndarray) else np.frombuffer(to_bytes(t), dtype=np.uint8),
                               dtype=np.uint8) for t in texts]
    nb = len(ts)
    if nb == 0:
        return []
    alloc = pinned_empty if pinned else (lambda k, dt: np.empty(k, dtype=dt))
    ns = (C.c_uint64 * nb)(*[t.size for t in ts])
    caps = (C.c_uint64 * nb)(*[t.size + 3 for t in ts])
    cnts = [alloc(t.size + 3, np.uint32) for t in ts]
    syms = [alloc(t.size + 3, np.int16) for t in ts]
    tp = (C.c_void_p * nb)(*[t.ctypes.data for t in ts])
    cp = (C.c_void_p * nb)(*[c.ctypes.data for c in cnts])
    sp = (C.c_void_p * nb)(*[s.ctypes.data for s in syms])
    infos = (BlockInfo * nb)()
    ctx.call("tc_blocks_encode", nb, tp, ns, 1 if with_mtf else 0, cp, sp, caps, infos)
    out = []
    for b in range(nb):
        i = infos[b]
        R = int(i.R)
        fin = np.array(i.final_list[: i.sigma], dtype=np.int16)
        out.append(CompressedBlock(ts[b].size, int(i.N), int(i.primary), int(i.sigma), fin, cnts[b][:R].copy(),
                                   syms[b][:R].copy(), with_mtf))
    return out


def decompress(blk: CompressedBlock, ctx=None) -> bytes:
    ctx = ctx or default_context()
    if blk.R == 0:
        return b""
    cap = max(blk.n, 1) + 2
    out = np.empty(cap, dtype=np.uint8)
    n_out = C.c_uint64(0)
    cnt = np.ascontiguousarray(blk.counts, dtype=np.uint32)
    sym = np.ascontiguousarray(blk.syms, dtype=np.int16)
    if blk.with_mtf:
        info = BlockInfo()
        info.n, info.N, info.primary, info.sigma, info.R = blk.n, blk.N, blk.primary, blk.sigma, blk.R
        for j, v in enumerate(blk.final_list.tolist()):
            info.final_list[j] = v
        ctx.call("tc_bwt_mtf_rle_decode", ptr(cnt), ptr(sym), C.byref(info), ptr(out), cap, C.byref(n_out))
    else:
        ctx.call("tc_bwt_rle_decode", ptr(cnt), ptr(sym), blk.R, ptr(out), cap, C.byref(n_out))
    return out[: n_out.value].tobytes()


def compress_blocks_packed(texts, with_mtf: bool = True, ctx=None, pinned: bool = True) -> list:
    """tc_blocks_encode_packed: one container (uint8 array) per input block, in input order."""
    from ._lib import pinned_empty
    ctx = ctx or default_context()
    ts = [np.ascontiguousarray(t if isinstance(t, np.ndarray) else np.frombuffer(to_bytes(t), dtype=np.uint8),
                               dtype=np.uint8) for t in texts]
    nb = len(ts)
    if nb == 0:
        return []
    alloc = pinned_empty if pinned else (lambda k, dt: np.empty(k, dtype=dt))
    bound = [int(ctx.L.tc_packed_bound(t.size)) for t in ts]
    outs = [alloc(b, np.uint8) for b in bound]
    ns = (C.c_uint64 * nb)(*[t.size for t in ts])
    caps = (C.c_uint64 * nb)(*bound)
    nbytes = (C.c_uint64 * nb)()
    tp = (C.c_void_p * nb)(*[t.ctypes.data for t in ts])
    op = (C.c_void_p * nb)(*[o.ctypes.data for o in outs])
    infos = (BlockInfo * nb)()
    ctx.call("tc_blocks_encode_packed", nb, tp, ns, 1 if with_mtf else 0, op, caps, nbytes, infos)
    return [outs[b][: int(nbytes[b])].copy() for b in range(nb)]


def unpack_block(blob) -> CompressedBlock:
    """tc_packed_unpack (host only, no device): container -> CompressedBlock with the run records."""
    L = load()
    a = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8) if not isinstance(blob, np.ndarray) else blob)
    info = BlockInfo()
    flags = C.c_uint32(0)
    rc = L.tc_packed_info(ptr(a), a.size, C.byref(info), C.byref(flags))
    if rc != TC_OK:
        _raise(None, rc)
    R = int(info.R)
    cnt = np.empty(R, dtype=np.uint32)
    sym = np.empty(R, dtype=np.int16)
    rc = L.tc_packed_unpack(ptr(a), a.size, ptr(cnt), ptr(sym), R, C.byref(info))
    if rc != TC_OK:
        _raise(None, rc)
    fin = np.array(info.final_list[: info.sigma], dtype=np.int16)
    return CompressedBlock(int(info.n), int(info.N), int(info.primary), int(info.sigma), fin, cnt, sym,
                           bool(flags.value & 1))


def decompress_packed(blob, ctx=None) -> bytes:
    """tc_packed_decode: container -> text (unpacked and decoded on the device)."""
    ctx = ctx or default_context()
    a = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8) if not isinstance(blob, np.ndarray) else blob)
    info = BlockInfo()
    rc = ctx.L.tc_packed_info(ptr(a), a.size, C.byref(info), None)
    if rc != TC_OK:
        _raise(None, rc)
    cap = int(info.n) + 2
    out = np.empty(cap, dtype=np.uint8)
    n_out = C.c_uint64(0)
    ctx.call("tc_packed_decode", ptr(a), a.size, ptr(out), cap, C.byref(n_out))
    return out[: n_out.value].tobytes()


def decompress_blocks_packed(blobs, ctx=None) -> list:
    """tc_blocks_decode_packed: a list of containers -> their texts, several containers in flight on the device
    (the inverse of compress_blocks_packed).  Reference quirk kept: in the BWT -> RLE chain (with_mtf False) seqToRLE
    writes a trailing Nothing twice, so a block whose BWT ends with its Nothing raises FromJustError on the way back,
    as fromBWT does in the reference."""
    ctx = ctx or default_context()
    arrs = [np.ascontiguousarray(np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b) for b in blobs]
    nb = len(arrs)
    if nb == 0:
        return []
    caps = []
    for a in arrs:
        info = BlockInfo()
        rc = ctx.L.tc_packed_info(ptr(a), a.size, C.byref(info), None)
        if rc != TC_OK:
            _raise(None, rc)
        caps.append(int(info.n) + 2)
    outs = [np.empty(c, dtype=np.uint8) for c in caps]
    bp = (C.c_void_p * nb)(*[a.ctypes.data for a in arrs])
    by = (C.c_uint64 * nb)(*[a.size for a in arrs])
    op = (C.c_void_p * nb)(*[o.ctypes.data for o in outs])
    cp = (C.c_uint64 * nb)(*caps)
    n_out = (C.c_uint64 * nb)()
    ctx.call("tc_blocks_decode_packed", nb, bp, by, op, cp, n_out)
    return [outs[b][: int(n_out[b])].tobytes() for b in range(nb)]
