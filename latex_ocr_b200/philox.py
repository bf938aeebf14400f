"""Host mirror (numpy) of the device Philox4x32-10 generator in csrc/lo_common.cuh — used by tests to reproduce the in-kernel
dropout masks and beam-search diversity draws, and documenting the stream layout:

  dropout multiplier of (row b, step t, unit j), call counter n, 64-bit seed s:
      words = philox4x32_10(counter = (j >> 2, t, b, n), key = (s & 0xffffffff, s >> 32));  u = (words[j & 3] >> 8) * 2**-24
      multiplier = 1/(1-p) if u >= p else 0                                   (nn.Dropout semantics, seq2seq_torch.py:316)
"""
import numpy as np

_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (broadcast)."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_M0) * c0
        p1 = np.uint64(_M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(_W0)) & mask
        k1 = (k1 + np.uint64(_W1)) & mask
    return c0, c1, c2, c3


def u01(words):
    return ((np.asarray(words, dtype=np.uint64) >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def dropout_multipliers(seed, call, B, T, D, p):
    """[B, T, D] float32 multipliers exactly as lstm_pw_fwd_kernel draws them."""
    b = np.arange(B, dtype=np.uint64)[:, None, None]
    t = np.arange(T, dtype=np.uint64)[None, :, None]
    j4 = np.arange(D // 4 + (1 if D % 4 else 0), dtype=np.uint64)[None, None, :]
    w = philox4x32_10(j4 + 0 * b + 0 * t, t + 0 * b + 0 * j4, b + 0 * t + 0 * j4, np.uint64(call), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = np.stack([u01(x) for x in w], axis=-1).reshape(B, T, -1)[:, :, :D]
    return np.where(u >= np.float32(p), np.float32(1.0 / (1.0 - p)), np.float32(0.0)).astype(np.float32)
