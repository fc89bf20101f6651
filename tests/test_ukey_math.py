"""CPU-only check of the arithmetic behind the suffix sort's uniform keys (csrc/sufsort.cu:
uk_lut_kernel / ukey_of): a numpy restatement of the same integer recurrences shows that the
32-bit key is weakly order-preserving in the packed suffix key for any alphabet and skew, and
close to uniform for memoryless text -- the two properties the MSD path relies on.  (The CUDA
implementation itself is checked on the GPU against the oracle's suffix arrays.)"""
import numpy as np
import pytest


def build_lut(probs, sigma, b, g):
    """cum16 << 16 and freq16 << 16 per g-gram, like uk_lut_kernel (float32 products, floor)."""
    E = 1 << (g * b)
    nvalid = float(sigma + 1) ** g
    S = np.float32(65536.0 - nvalid - 64.0)
    e = np.arange(E)
    pr = np.ones(E, dtype=np.float32)
    for j in range(g):
        c = (e >> (b * (g - 1 - j))) & ((1 << b) - 1)
        pr = pr * np.where(c <= sigma, probs[np.minimum(c, sigma)], 0).astype(np.float32)
    f = np.where(pr > 0, np.maximum(1, np.floor(pr * S)), 0).astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(f)[:-1]])
    over = cum > 65535
    cum = np.where(over, 65535, cum)
    f = np.where(over, 0, f)
    f = np.minimum(f, 65536 - cum)
    f = np.minimum(f, 65535)
    return (cum << 16).astype(np.uint64), (f << 16).astype(np.uint64)


def ukey(keys, kb, gb, G, cumS, freqS):
    x = np.zeros(keys.shape, dtype=np.uint64)
    r = np.full(keys.shape, 0xFFFFFFFF, dtype=np.uint64)
    sh = kb - gb
    m = np.uint64((1 << gb) - 1)
    for _ in range(G):
        s = (keys >> np.uint64(sh)) & m
        x = x + ((r * cumS[s]) >> np.uint64(32))       # __umulhi(r, cum << 16)
        r = (r * freqS[s]) >> np.uint64(32)
        sh -= gb
    assert int(x.max()) < (1 << 32)
    return x


@pytest.mark.parametrize("sigma,skew", [(4, False), (5, True), (2, False), (20, True), (95, False), (256, False), (256, True)])
def test_ukey_is_order_preserving_and_uniform(sigma, skew):
    rng = np.random.default_rng(sigma * 7 + skew)
    n = 200_000
    p = rng.random(sigma) ** (4 if skew else 0) + 1e-9
    p /= p.sum()
    text = rng.choice(sigma, size=n, p=p) + 1                      # codes 1..sigma, 0 = past the end
    b = max(1, int(sigma).bit_length())
    k = 64 // b
    kb = k * b
    gs = max(1, min(9 // b, k))
    gb = gs * b
    hist = np.bincount(text, minlength=sigma + 1).astype(np.float64)
    probs = hist / (n + 1)
    probs[0] = 1.0 / (n + 1)
    H = -(hist[1:][hist[1:] > 0] / n * np.log2(hist[1:][hist[1:] > 0] / n)).sum()
    G = max(1, min(k // gs, int(np.ceil(34.0 / (gs * max(H, 0.02))))))
    cumS, freqS = build_lut(probs, sigma, b, gs)
    padded = np.concatenate([text, np.zeros(k, dtype=np.int64)])
    keys = np.zeros(n, dtype=np.uint64)
    for j in range(k):                                             # first k symbols, MSB first
        keys = (keys << np.uint64(b)) | padded[j:j + n].astype(np.uint64)
    u = ukey(keys, kb, gb, G, cumS, freqS)
    order = np.argsort(keys, kind="stable")
    assert np.all(np.diff(u[order].astype(np.int64)) >= 0), "ukey must never invert the key order"
    # balance of the two 8-bit partition digits (the reason the key exists): over DISTINCT keys
    # (equal keys necessarily share a bucket) the fullest of the 65,536 top-16-bit buckets stays
    # far below the 512 records one warp sorts
    uu = u[np.unique(keys, return_index=True)[1]]
    buckets = np.bincount((uu >> np.uint64(16)).astype(np.int64), minlength=1 << 16)
    assert buckets.max() <= max(64, 8 * uu.size // 65536), (sigma, skew, int(buckets.max()), uu.size)
