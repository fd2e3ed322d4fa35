"""pure copy-engine H2D bandwidth from pinned memory: one 1.26 GB copy, then the same bytes as 256 page-sized copies
(B200 pod, this round: 55.6 GB/s and 53.0 GB/s) — the ceiling the e2e number of bench.py is measured against"""
import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from retto_b200.api import Context
ctx = Context(0); L, H = ctx._L, ctx._h
n = 256 * 1280 * 1280 * 3
hp = C.c_void_p(); ctx._check(L.retto_b200_host_alloc(H, n, C.byref(hp)))
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for chunk in (n, 1280 * 1280 * 3):
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        for off in range(0, n, chunk):
            L.retto_b200_h2d(H, d.data_ptr() + off, hp.value + off, chunk)
        ctx.sync(); dt = time.perf_counter() - t
    print("chunk", chunk, "GB/s", n / dt / 1e9, "ms", dt * 1e3)
