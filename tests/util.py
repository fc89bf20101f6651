import numpy as np


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n uint64 values of splitmix64 (SURVEY.md 8d generator), vectorised."""
    idx = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) + np.uint64(seed)
    z = idx
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def gen_acgt(seed, n):
    return np.frombuffer(b"ACGT", dtype=np.uint8)[(splitmix64(seed, n) >> np.uint64(62)).astype(np.int64)]


def gen_acgtn(seed, n, p_n=0.01):
    r = splitmix64(seed, n)
    u = (r >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    base = np.frombuffer(b"ACGT", dtype=np.uint8)[(r & np.uint64(3)).astype(np.int64)]
    return np.where(u < p_n, np.uint8(ord("N")), base).astype(np.uint8)


def gen_bytes(seed, n):
    return (splitmix64(seed, n) >> np.uint64(56)).astype(np.uint8)


def gen_ascii(seed, n):
    return (0x20 + (splitmix64(seed, n) >> np.uint64(33)) % np.uint64(95)).astype(np.uint8)


def gen_words(seed, n, vocab=600, dup_every=1 << 18, dup_len=6000):
    """Correlated, repetitive text: Zipf-like draws from a fixed vocabulary of short lower-case words, indented
    lines, and a block copied verbatim from earlier in the text every `dup_every` bytes (long common prefixes,
    like duplicated files in a source tree).  Deterministic."""
    r = splitmix64(seed, 4 * vocab + n // 2 + 16)
    words = []
    for w in range(vocab):
        ln = 2 + int(r[4 * w] % np.uint64(9))
        letters = (r[4 * w + 1] >> (np.arange(ln, dtype=np.uint64) * np.uint64(5))) % np.uint64(26)
        words.append(bytes((0x61 + letters).astype(np.uint8)))
    out = bytearray()
    k = 4 * vocab
    next_dup = dup_every
    while len(out) < n:
        x = int(r[k]); k += 1
        u = (x >> 11) / float(1 << 53)
        out += words[int(vocab * u * u * u)]          # cubic skew: a few words dominate
        sep = (x >> 3) & 31
        out += b"\n        " if sep == 0 else b"\n    " if sep == 1 else b"." if sep == 2 else b" "
        if len(out) >= next_dup:
            src = (x >> 17) % (len(out) - dup_len)
            out += out[src: src + dup_len]
            next_dup += dup_every
    return np.frombuffer(bytes(out[:n]), dtype=np.uint8).copy()


def python_corpus(nbytes):
    """Real, correlated, repetitive text: Python sources of this interpreter's site-packages, concatenated in the
    order of a sorted top-down walk that STOPS as soon as nbytes have been read (enumerating the whole tree first costs
    a minute or more on a freshly started box whose image is paged in on demand).  The same image runs here and on
    the GPU box, so the corpus is the same on both.  Returns fewer bytes if the tree holds less."""
    import os
    import sysconfig
    buf = bytearray()
    for d, dirs, files in os.walk(sysconfig.get_paths()["purelib"]):
        dirs.sort()
        for f in sorted(files):
            if f.endswith(".py"):
                try:
                    with open(os.path.join(d, f), "rb") as fh:
                        buf += fh.read()
                except OSError:
                    pass
                if len(buf) >= nbytes:
                    return np.frombuffer(bytes(buf[:nbytes]), dtype=np.uint8)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def gen_reads(seed, text, q, m, mut_frac=0.10):
    """q reads of m symbols at uniform offsets; mut_frac of them get one random substitution."""
    r = splitmix64(seed, 3 * q)
    offs = (r[:q] % np.uint64(text.size - m + 1)).astype(np.int64)
    reads = text[offs[:, None] + np.arange(m)[None, :]].copy()
    mut = (r[q:2 * q] % np.uint64(1000)) < np.uint64(int(mut_frac * 1000))
    pos = (r[2 * q:] % np.uint64(m)).astype(np.int64)
    sub = np.frombuffer(b"ACGT", dtype=np.uint8)[((r[2 * q:] >> np.uint64(40)) & np.uint64(3)).astype(np.int64)]
    rows = np.nonzero(mut)[0]
    reads[rows, pos[rows]] = sub[rows]
    return reads


def pack_container(n, N, primary, sigma, final_list, counts, syms, with_mtf) -> np.ndarray:
    """Independent numpy writer of the packed block container (include/tc_b200.h, tc_packed_header):
    what tc_blocks_encode_packed must produce byte for byte from the oracle's runs."""
    import struct
    counts = np.asarray(counts, dtype=np.uint32)
    syms = np.asarray(syms, dtype=np.int16)
    R = int(counts.size)
    big = np.nonzero(counts >= 16)[0].astype(np.uint64)
    al = lambda x: (x + 15) & ~15
    off_cnt4 = 640
    off_sym8 = al(off_cnt4 + (R + 1) // 2)
    off_hi = al(off_sym8 + R)
    off_big_idx = al(off_hi + (R + 31) // 32 * 4)
    off_big_cnt = al(off_big_idx + 8 * big.size)
    total = al(off_big_cnt + 4 * big.size)
    out = np.zeros(total, dtype=np.uint8)
    fl = np.zeros(257 + 7, dtype=np.int16)
    fl[: len(final_list)] = np.asarray(final_list, dtype=np.int16)
    hdr = struct.pack("<QII6Q5QII", 0x314B4C4242434254, 2, 1 if with_mtf else 0, n, N, primary, R, int(big.size), total,
                      off_cnt4, off_sym8, off_hi, off_big_idx, off_big_cnt, sigma, 0) + fl.tobytes()
    assert len(hdr) == 640
    out[:640] = np.frombuffer(hdr, dtype=np.uint8)
    code = syms.astype(np.int32) & 0x1FF
    nib = np.zeros((R + 1) // 2 * 2, dtype=np.uint8)
    nib[:R] = np.minimum(counts.astype(np.int64) - 1, 15).astype(np.uint8)
    out[off_cnt4:off_cnt4 + (R + 1) // 2] = nib[0::2] | (nib[1::2] << 4)
    out[off_sym8:off_sym8 + R] = (code & 0xFF).astype(np.uint8)
    bits = np.zeros((R + 31) // 32 * 32, dtype=np.uint8)
    bits[:R] = code >> 8
    out[off_hi:off_hi + bits.size // 8] = np.packbits(bits, bitorder="little")
    out[off_big_idx:off_big_idx + 8 * big.size] = big.view(np.uint8)
    out[off_big_cnt:off_big_cnt + 4 * big.size] = counts[big.astype(np.int64)].view(np.uint8)
    return out
