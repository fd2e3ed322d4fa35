"""Replay forwards shared by the reference-side dump (oracle/ref_dump: `replay_cls` / `replay_rec` in the Rust patch) and the
oracle pipeline: every output is an INTEGER-hash function of the bytes of the input tensor, so both sides produce bit-identical
logits without exchanging them.  det returns the probability map that was generated with the page."""
import numpy as np

MASK = (1 << 64) - 1


def fnv1a(a: np.ndarray) -> int:
    h = 0xCBF29CE484222325
    for b in np.ascontiguousarray(a, dtype="<f4").tobytes():
        h ^= b
        h = (h * 0x100000001B3) & MASK
    return h


def splitmix(x: int) -> int:
    z = (x + 0x9E3779B97F4A7C15) & MASK
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
    return z ^ (z >> 31)


class ReplayWorker:
    def __init__(self, prob, n_classes=6625):
        self.prob, self.C = prob, n_classes

    def det(self, x):
        assert x.shape[2:] == self.prob.shape
        return self.prob[None, None]

    def cls(self, x):
        n = x.shape[0]
        out = np.zeros((n, 2), np.float32)
        for i in range(n):
            h = splitmix(fnv1a(x[i]))
            is180 = h % 10 < 3
            s = np.float32(0.95) if (h >> 8) % 10 < 3 else np.float32(0.55)
            out[i] = (np.float32(1.0) - s, s) if is180 else (s, np.float32(1.0) - s)
        return out

    def rec(self, x):
        n, T = x.shape[0], x.shape[3] // 8
        out = np.zeros((n, T, self.C), np.float32)
        for i in range(n):
            seed, prev = fnv1a(x[i]), 0
            for k in range(T):
                h = splitmix((seed + k) & MASK)
                u = h % 100
                c = 0 if u < 45 else (prev if u < 65 else 1 + ((h >> 16) % (self.C - 1)))
                prev = c
                out[i, k, c] = np.float32(0.5) + np.float32((h >> 40) % 1000) / np.float32(2000)
        return out
