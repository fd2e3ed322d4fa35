"""Stand-in forward passes for demos and tests of the entry points (retto_b200.cli --worker tools.demo_worker:make_worker):
every output is a deterministic function of the input tensor only, so any batching of the pages gives the same rows.
NOT a model: det marks dark pixels as text, cls/rec hash the batch tensor."""
import zlib

import numpy as np


class StatelessWorker:
    def det(self, x):
        g = (x[0].mean(axis=0) + 1.0) / 2.0
        return np.clip(1.0 - g, 0.0, 1.0).astype(np.float32)[None, None]

    def cls(self, x):
        s = x.reshape(x.shape[0], -1).mean(axis=1)
        return np.stack([np.where(s > -0.6, 0.95, 0.05), np.where(s > -0.6, 0.05, 0.95)], 1).astype(np.float32)

    def rec(self, x, n_classes=6625):
        n, T = x.shape[0], x.shape[3] // 8
        out = np.zeros((n, T, n_classes), np.float32)
        for i in range(n):
            rng = np.random.default_rng(zlib.crc32(np.ascontiguousarray(x[i]).tobytes()))
            out[i, np.arange(T), rng.integers(0, n_classes, T)] = 0.5 + 0.5 * rng.random(T, dtype=np.float32)
        return out


def make_worker(device_id, args=None):
    from retto_b200.session import CallableWorker
    w = StatelessWorker()
    return CallableWorker(w.det, w.cls, w.rec)
