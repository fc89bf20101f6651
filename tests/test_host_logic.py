"""CPU-only tests: the C-ABI library loads and exports every symbol include/tc_b200.h declares,
host-side value types behave like the reference's, and the product fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "tc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from text_compression_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    # and the Python binding declares a signature for each of them
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (no oracle, no CPU path)."""
    import torch
    from text_compression_b200 import _lib, NoDeviceError
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(NoDeviceError):
        _lib.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "text_compression_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+\S*oracle|liborc|tc_oracle|#include\s+\"[^\"]*oracle", txt, re.M), f


def test_value_types():
    from text_compression_b200.seq import MTF, RLE, MaybeSeq
    s = MaybeSeq.from_list([b"a", None, b"b"], "B")
    assert s.codes.tolist() == [97, -1, 98] and s.to_list() == [b"a", None, b"b"]
    assert s.as_kind("T").to_list() == ["a", None, "b"]
    assert s == MaybeSeq(np.array([97, -1, 98]), "B") and s != s.as_kind("T")
    # Q6: the T variants decodeUtf8 every byte on its own -> bytes >= 0x80 throw
    with pytest.raises(UnicodeDecodeError):
        MaybeSeq(np.array([0xC3, 0xA9]), "T")
    r = RLE.from_list([b"4", b"a", b"1", None, b"12", b"c", b"7"], "B")       # odd trailing element ignored
    assert r.counts.tolist() == [4, 1, 12] and r.syms.tolist() == [97, -1, 99]
    assert r.to_list() == [b"4", b"a", b"1", None, b"12", b"c"]
    assert RLE.from_list(["3", "x"], "T").to_list() == ["3", "x"]
    from text_compression_b200 import FromJustError
    with pytest.raises(FromJustError):
        RLE.from_list([None, b"a"], "B")
    m = MTF(np.array([1, 0]), MaybeSeq.from_list([b"b", None], "B"))
    assert m.to_tuple() == ([1, 0], [b"b", None])


def test_pack_patterns():
    from text_compression_b200.fmindex import pack_patterns
    flat, off = pack_patterns([b"abc", "", "de"])
    assert flat.tobytes() == b"abcde" and off.tolist() == [0, 3, 3, 5]
    flat, off = pack_patterns([])
    assert flat.size == 0 and off.tolist() == [0]


def test_sharding_helpers():
    from text_compression_b200 import multi
    assert multi.blocks_of_rank(10, 4, 1) == [1, 5, 9]
    cover = sorted(b for r in range(8) for b in multi.blocks_of_rank(512, 8, r))
    assert cover == list(range(512))
    for q, ws in ((10, 3), (7, 8), (0, 2), (10_000_000, 8)):
        sl = [multi.query_slice(q, ws, r) for r in range(ws)]
        assert sl[0][0] == 0 and sl[-1][1] == q
        assert all(sl[i][1] == sl[i + 1][0] for i in range(ws - 1))
        assert max(b - a for a, b in sl) - min(b - a for a, b in sl) <= 1


def test_generators_are_deterministic():
    from tests.util import gen_acgt, gen_acgtn, gen_bytes, gen_reads
    assert gen_bytes(0xC2, 8).tolist() == gen_bytes(0xC2, 8).tolist()
    t = gen_acgtn(0xC3, 100000)
    assert set(np.unique(t).tolist()) <= set(b"ACGTN") and 500 < int((t == ord("N")).sum()) < 1500
    r = gen_reads(1, gen_acgt(2, 5000), 100, 20)
    assert r.shape == (100, 20)


def test_packed_container_host_side():
    """The packed block container (include/tc_b200.h): tc_packed_unpack / tc_packed_info are host-only
    and must read what an independent numpy writer (tests/util.pack_container) lays out from the
    ORACLE's runs -- long runs (count >= 255 exceptions), Nothing symbols, MTF index 256."""
    import ctypes as C
    from oracle import oracle
    from tests.util import gen_acgtn, gen_bytes, pack_container
    from text_compression_b200 import _lib, block
    L = _lib.load()
    texts = [gen_acgtn(7, 5000), np.frombuffer(b"a" * 3000 + b"b" * 255 + b"c" * 254 + b"ab" * 40, np.uint8),
             np.tile(np.arange(256, dtype=np.uint8), 5), gen_bytes(9, 3000), np.frombuffer(b"x", np.uint8)]
    for t in texts:
        bwt = oracle.bwt_encode(t)
        primary = int(np.nonzero(bwt < 0)[0][0])
        for with_mtf in (False, True):
            if with_mtf:
                idx, fin = oracle.mtf_encode(bwt)
                cnt, sym = oracle.rle_encode(idx.astype(np.int16))
            else:
                fin = np.empty(0, np.int16)
                cnt, sym = oracle.rle_encode(bwt)
            blob = pack_container(t.size, t.size + 1, primary, len(fin), fin, cnt, sym, with_mtf)
            assert blob.size <= L.tc_packed_bound(t.size)
            got = block.unpack_block(blob)
            assert (got.n, got.N, got.primary, got.sigma, got.with_mtf) == (t.size, t.size + 1, primary, len(fin), with_mtf)
            assert got.final_list.tolist() == fin.tolist()
            assert np.array_equal(got.counts, cnt) and np.array_equal(got.syms, sym)
    assert (cnt >= 0).all() and any((oracle.rle_encode(oracle.bwt_encode(texts[1]))[0] >= 255).tolist())
    assert 256 in oracle.mtf_encode(oracle.bwt_encode(texts[2]))[0].tolist()
    # malformed containers are refused, never read out of bounds
    info = _lib.BlockInfo()
    bad = blob.copy(); bad[0] ^= 1
    assert L.tc_packed_info(_lib.ptr(bad), bad.size, C.byref(info), None) == _lib.TC_E_ARG
    assert L.tc_packed_info(_lib.ptr(blob), blob.size - 1, C.byref(info), None) == _lib.TC_E_ARG
    assert L.tc_packed_info(_lib.ptr(blob), 100, C.byref(info), None) == _lib.TC_E_ARG
    bad = blob.copy(); bad[40:48] = np.frombuffer(np.uint64(1 << 40).tobytes(), np.uint8)   # R
    assert L.tc_packed_info(_lib.ptr(bad), bad.size, C.byref(info), None) == _lib.TC_E_ARG
    cnt1 = np.empty(1, np.uint32); sym1 = np.empty(1, np.int16)
    blob2 = pack_container(3, 4, 0, 0, [], [1, 2, 1], [5, 6, -1], False)
    assert L.tc_packed_unpack(_lib.ptr(blob2), blob2.size, _lib.ptr(cnt1), _lib.ptr(sym1), 1, C.byref(info)) == _lib.TC_E_CAP
    assert info.R == 3


def test_header_is_plain_c():
    """include/tc_b200.h is the drop-in boundary: it must compile as C99 (no C++-isms), and
    tc_packed_header must have the 640-byte layout the container format states."""
    import subprocess
    import tempfile
    src = ('#include "tc_b200.h"\n#include <stddef.h>\n'
           'typedef char a1[sizeof(tc_packed_header) == 640 ? 1 : -1];\n'
           'typedef char a2[offsetof(tc_packed_header, R) == 40 ? 1 : -1];\n'
           'typedef char a3[offsetof(tc_packed_header, final_list) == 112 ? 1 : -1];\n'
           'typedef char a4[offsetof(tc_block_info, R) == 544 && sizeof(tc_block_info) == 552 ? 1 : -1];\n'
           'int main(void) { return tc_packed_bound(0) > 0 ? 0 : 1; }\n')
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.c")
        open(p, "w").write(src)
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I",
                            os.path.join(ROOT, "include"), p], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_packed_container_fuzz():
    """tc_packed_unpack against the numpy writer on random run sequences (random counts incl. >= 255,
    all 258 symbol codes), and on corrupted containers: every outcome is TC_OK or TC_E_ARG/TC_E_CAP,
    never a crash, and an accepted container unpacks to R runs."""
    import ctypes as C
    from tests.util import pack_container
    from text_compression_b200 import _lib, block
    L = _lib.load()
    rng = np.random.default_rng(20261018)
    for case in range(150):
        R = int(rng.integers(0, 700))
        cnt = rng.integers(1, 6, size=R).astype(np.uint32)
        big = rng.random(R) < 0.03
        cnt[big] = rng.integers(255, 1 << 31, size=int(big.sum())).astype(np.uint32)
        sym = rng.integers(-1, 257, size=R).astype(np.int16)
        with_mtf = bool(case & 1)
        fin = rng.permutation(np.arange(-1, 256, dtype=np.int16))[: int(rng.integers(0, 258))]
        blob = pack_container(12345, 12346, 7, len(fin), fin, cnt, sym, with_mtf)
        got = block.unpack_block(blob)
        assert np.array_equal(got.counts, cnt) and np.array_equal(got.syms, sym)
        assert got.final_list.tolist() == fin.tolist() and got.with_mtf == with_mtf and got.primary == 7
        # corruption: a few random byte flips anywhere (header included)
        bad = blob.copy()
        for pos in rng.integers(0, bad.size, size=3):
            bad[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))
        info = _lib.BlockInfo()
        c = np.empty(R + 8, np.uint32)
        s = np.empty(R + 8, np.int16)
        rc = L.tc_packed_unpack(_lib.ptr(bad), bad.size, _lib.ptr(c), _lib.ptr(s), R + 8, C.byref(info))
        assert rc in (_lib.TC_OK, _lib.TC_E_ARG, _lib.TC_E_CAP)
        if rc == _lib.TC_OK:
            assert int(info.R) <= R + 8


def test_stream_framing_host_side():
    """TCZ1 framing (no device): split_stream reads what compress_stream's layout states and
    refuses truncated or foreign input."""
    import struct
    from tests.util import pack_container
    from text_compression_b200 import stream
    c1 = pack_container(3, 4, 0, 0, [], [1, 2, 1], [5, 6, -1], False).tobytes()
    c2 = pack_container(0, 0, 0, 0, [], [], [], True).tobytes()
    z = struct.pack("<4sIQQ", b"TCZ1", 3, 3, 2) + struct.pack("<Q", len(c1)) + c1 + struct.pack("<Q", len(c2)) + c2
    bs, total, parts = stream.split_stream(z)
    assert (bs, total, parts) == (3, 3, [c1, c2])
    for bad in (z[:10], z[:40], b"XXXX" + z[4:], z[:-1]):
        with pytest.raises(ValueError):
            stream.split_stream(bad)


# ---- the Haskell boundary, as far as it can be checked without GHC --------------------------------------------
def _hs_exports(path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("extract_exports", os.path.join(ROOT, "tests", "golden", "extract_exports.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.exports(path)


def test_haskell_modules_keep_the_reference_export_lists():
    """haskell/src/Data/{BWT,MTF,RLE,FMIndex}[/Internal].hs export exactly the names the reference modules export
    (tests/golden/reference_exports.json, extracted from the reference by tests/golden/extract_exports.py);
    Data.RLE.Internal adds one helper for its sibling modules."""
    import json
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_exports.json")))
    assert len(want) == 8
    for mod, names in want.items():
        path = os.path.join(ROOT, "haskell", "src", *mod.split(".")) + ".hs"
        got_mod, got = _hs_exports(path)
        assert got_mod == mod
        extra = sorted(set(got) - set(names))
        assert not (set(names) - set(got)), (mod, "missing", sorted(set(names) - set(got)))
        assert extra == (["symbolBytes"] if mod == "Data.RLE.Internal" else []), (mod, extra)
    # every exported function has a top-level type signature in the module
    for mod in want:
        src = open(os.path.join(ROOT, "haskell", "src", *mod.split(".")) + ".hs").read()
        for name in want[mod]:
            if name[0].islower() and "(" not in name:
                assert re.search(r"^%s\s*::" % re.escape(name), src, re.M), (mod, name)


def test_abi_layout_c99():
    """tests/c/abi_layout.c: the struct offsets the Haskell shim reads (544 / 552 / 16) and every foreign import's
    argument list, checked by gcc against include/tc_b200.h."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-c",
                               os.path.join(ROOT, "tests", "c", "abi_layout.c"), "-o", os.path.join(d, "abi.o")])


def test_mtf_decode_step_on_cpu():
    """tests/c/mtfd_select_test.cpp: the select-based MTF decode step the kernels run
    (text_compression_b200/csrc/mtfd_select.cuh, one source for host and device) on plain arrays against a list
    shifted by hand: 4000 chunks, alphabets of 9..257 entries, uniform / small / zero / always-the-back indices."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "mtfd_select_test")
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-Wall", "-I", os.path.join(ROOT, "text_compression_b200", "csrc"),
                               "-x", "c++", os.path.join(ROOT, "tests", "c", "mtfd_select_test.cpp"), "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


_HS2C = {"CInt": "i32", "Word32": "i32", "Int32": "i32", "Word64": "i64", "Int64": "i64", "CSize": "i64",
         "CString": "ptr"}


def _hs_kind(t):
    t = t.strip()
    if t == "()":
        return "void"
    if t.startswith("(") and t.endswith(")"):
        t = t[1:-1].strip()
    if t.startswith(("Ptr", "FunPtr")):
        return "ptr"
    return _HS2C[t]


def _c_kind(t):
    t = t.strip()
    if "*" in t:
        return "ptr"
    t = re.sub(r"\b(const|unsigned)\b", "", t).split()
    base = t[0] if t else "void"
    return {"void": "void", "int": "i32", "uint32_t": "i32", "int32_t": "i32", "uint64_t": "i64", "int64_t": "i64",
            "size_t": "i64"}[base]


def test_foreign_imports_match_the_header():
    """Every `foreign import ccall` in B200.hs against the prototype of the same name in include/tc_b200.h:
    same arity, and per argument the same kind (pointer / 32-bit / 64-bit) in the same order."""
    hs = open(os.path.join(ROOT, "haskell", "src", "Data", "TextCompression", "B200.hs")).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "tc_b200.h")).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"([\w \*]+?)\b(tc_\w+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        args = [a for a in (x.strip() for x in args.split(",")) if a and a != "void"]
        protos[name] = (_c_kind(ret), [_c_kind(re.sub(r"\b\w+$", "", a) if not a.endswith("*") else a) for a in args])
    imports = re.findall(r'foreign import ccall \w+ "(&?)(tc_\w+)"\s*\n?\s*\w+\s*::\s*(.*?)(?=\nforeign import|\n\n|\n--)', hs, flags=re.S)
    assert len(imports) >= 25
    for amp, name, sig in imports:
        assert name in protos, name
        if amp:      # address import (finalisers): only the symbol must exist
            continue
        parts = [p.strip() for p in re.sub(r"\s+", " ", sig).split("->")]
        res = re.sub(r"^IO\s*", "", parts[-1]).strip()
        got = (_hs_kind(res), [_hs_kind(p) for p in parts[:-1]])
        assert got == protos[name], (name, got, protos[name])


def test_struct_offsets_used_by_the_shim():
    """The byte offsets B200.hs hard-codes, against ctypes' layout of the same structs (which follows the C ABI)."""
    from text_compression_b200._lib import BlockInfo, FmInfo
    import ctypes as C
    hs = open(os.path.join(ROOT, "haskell", "src", "Data", "TextCompression", "B200.hs")).read()
    assert C.sizeof(BlockInfo) == 552 and BlockInfo.R.offset == 544
    assert set(re.findall(r"peekByteOff pinfo (\d+)", hs)) == {"544"}
    assert set(re.findall(r"pinned \((\d+) \* nb\)", hs)) == {"552"}
    assert C.sizeof(FmInfo) == 2624 and FmInfo.C.offset == 552
