"""TEST INFRASTRUCTURE (oracle): numpy restatement of the in-kernel random draws of csrc/sampling.cuh.

Philox4x32-10 is the published counter-based generator of Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as
1, 2, 3" (SC'11) -- the algorithm behind torch.rand on CUDA (the reference's helper.py:126,227 draw with torch.rand); it is
restated here from the paper (multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments 0x9E3779B9 / 0xBB67AE85, 10 rounds) and
pinned against the known-answer vectors of the authors' Random123 distribution (tests/test_oracle.py::test_philox_known_answers).
The library's use of it -- key = seed, counter = (column / 4, row, offset lo, offset hi | stream << 30), word column % 4,
(x >> 8) * 2^-24 -- is this module's `uniform`.  Only tests/ may import this file.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: 4 uint32 arrays (broadcastable), key: 2 ints -> 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & MASK for x in np.broadcast_arrays(*ctr)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def uniform(seed: int, offset: int, stream: int, rows: int, cols: int) -> np.ndarray:
    """float32 [rows, cols] in [0, 1): the draws aon_rng_uniform writes (stream 0 = stratified jitter, 1 = inverse cdf)."""
    seed &= 0xFFFFFFFFFFFFFFFF
    r, c = np.meshgrid(np.arange(rows, dtype=np.uint64), np.arange(cols, dtype=np.uint64), indexing="ij")
    c3 = ((offset >> 32) & 0x3FFFFFFF) | (stream << 30)
    x = philox4x32_10((c >> np.uint64(2), r, np.uint64(offset & 0xFFFFFFFF), np.uint64(c3)), (seed & 0xFFFFFFFF, seed >> 32))
    w = np.choose((c & np.uint64(3)).astype(np.int64), x)
    return ((w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
