"""Deterministic stand-in forwards shared by the GPU parity tests (numpy on both sides: the oracle pipeline calls them
directly, the CUDA session through CallableWorker)."""
import zlib

import numpy as np


class NpWorker:
    """deterministic stand-in for RettoInnerWorker used on BOTH sides (oracle gets numpy, the CUDA session
    gets the same functions through CallableWorker)."""

    def __init__(self, probmap, n_classes=6625):
        self.probmap = probmap
        self.C = n_classes

    def det(self, x):
        assert x.shape[0] == 1 and x.shape[1] == 3
        assert x.shape[2:] == self.probmap.shape, (x.shape, self.probmap.shape)
        return self.probmap[None, None]

    def cls(self, x):
        n = x.shape[0]
        out = np.zeros((n, 2), np.float32)
        for i in range(n):
            left, right = float(x[i, :, :, :96].sum()), float(x[i, :, :, 96:].sum())
            s = np.float32(0.95 if (int(abs(left) * 7) % 3 == 0) else 0.6)
            out[i] = (1 - s, s) if left > right else (s, 1 - s)
        return out

    def rec(self, x):
        n, _, _, W = x.shape
        T = W // 8
        out = np.zeros((n, T, self.C), np.float32)
        for i in range(n):
            rng = np.random.default_rng(zlib.crc32(np.ascontiguousarray(x[i]).tobytes()))
            out[i] = rng.random((T, self.C), dtype=np.float32) * np.float32(1e-3)
            cls = rng.integers(0, self.C, T)
            cls[rng.random(T) < 0.4] = 0
            out[i, np.arange(T), cls] = 0.5 + 0.5 * rng.random(T, dtype=np.float32)
        return out
