"""Dropout oracle (test infrastructure): the keep mask of csrc/dropout.cu restated in numpy.

The reference uses torch.nn.Dropout (pkg/models/pet_models/pet_cnn.py:26-27,38-40), whose CUDA RNG stream is an
implementation detail of ATen that no other implementation can reproduce; what is pinned here is (a) the generator
itself - Philox4x32-10 against the Random123 known-answer vectors - and (b) the product's documented mapping
(seed, offset, element index) -> keep bit, so that the CUDA mask can be checked bit for bit.  The nn.Dropout contract
(scale 1/(1-p), same mask in backward, identity in eval) is checked by the tests on top of this.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint64 arrays holding 32-bit values; returns four uint32-valued uint64 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) for v in (c0, c1, c2, c3))
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def keep_mask(n, p, seed, offset):
    """Boolean keep mask of `n` elements: element i belongs to vector v = i // 8; the vector draws Philox counters
    (2v, 2v+1) with (c2, c3) = offset and key = seed; lane j of the 8 words is kept iff (word >> 8) >= round(p * 2^24)."""
    nvec = (n + 7) // 8
    v = np.arange(nvec, dtype=np.uint64)
    thresh = np.uint64(int(p * 16777216.0 + 0.5))
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    c2 = np.full(nvec, offset & 0xFFFFFFFF, dtype=np.uint64)
    c3 = np.full(nvec, (offset >> 32) & 0xFFFFFFFF, dtype=np.uint64)
    words = []
    for half in (0, 1):
        ctr = v * np.uint64(2) + np.uint64(half)
        words += list(philox4x32_10(ctr & MASK32, ctr >> np.uint64(32), c2, c3, k0, k1))
    w = np.stack(words, axis=1)                       # (nvec, 8)
    return ((w >> np.uint64(8)) >= thresh).reshape(-1)[:n]
