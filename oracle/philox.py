"""NumPy restatement of the engine's in-kernel noise generator -- TEST INFRASTRUCTURE.

The reference has no counterpart: its noise is the global legacy NumPy stream
(``autompc/control/mppi.py:23``), which a GPU cannot reproduce cheaply.  The
engine's performance mode uses Philox4x32-10 (Salmon et al., SC'11) keyed by
(seed, solve counter, global sample, step) + Box-Muller; this file restates that
generator so tests can check the device stream bit-for-bit at the integer level
and to ~1e-6 after the float transforms.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32).copy() for c in (c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def mppi_noise(seed, counter, H, K, nu, sigma, k_offset=0):
    """(H, K, nu) float64: what ``ampc_mppi_get_noise`` returns (up to float32 rounding)."""
    nblk = (nu + 3) // 4
    h, k, blk = np.meshgrid(np.arange(H, dtype=np.uint32), np.arange(K, dtype=np.uint32),
                            np.arange(nblk, dtype=np.uint32), indexing="ij")
    r = philox4x32_10(k + np.uint32(k_offset), h | (blk << np.uint32(16)),
                      np.full(h.shape, counter & 0xFFFFFFFF, dtype=np.uint32),
                      np.full(h.shape, (counter >> 32) & 0xFFFFFFFF, dtype=np.uint32),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.zeros((H, K, nblk * 4))
    for q in range(2):
        u1 = ((r[2 * q] >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24
        u2 = (r[2 * q + 1] >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
        rad = np.sqrt(-2.0 * np.log(u1))
        n0, n1 = rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)
        for b in range(nblk):
            out[:, :, 4 * b + 2 * q] = n0[:, :, b]
            out[:, :, 4 * b + 2 * q + 1] = n1[:, :, b]
    return out[:, :, :nu] * np.sqrt(sigma)
