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
