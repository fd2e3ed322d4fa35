"""Side benchmark of the device JPEG decode (csrc/jpeg_decode.cu): 256 rendered 1280x1280 pages (32 unique), quality 90, 4:2:0,
with one restart interval per MCU row and without restart markers; per-kernel CUDA-event times.  python tools/bench_jpeg.py [n] [size]"""
import io
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import ctypes as C
    import torch
    from PIL import Image
    from retto_b200 import _lib
    from retto_b200.api import Context, image_info
    from tools.synth import gen_page
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
    uniq = [gen_page(4 + i, S, S)[0] for i in range(32)]
    ctx = Context(0)
    out = {}
    for tag, kw in (("dri_mcu_row", dict(restart_marker_rows=1)), ("no_dri", dict()), ("dri_mcu_row_444", dict(restart_marker_rows=1, subsampling=0))):
        files = []
        for u in uniq:
            b = io.BytesIO()
            k = dict(quality=90, subsampling=2)
            k.update(kw)
            Image.fromarray(u).save(b, "JPEG", **k)
            files.append(b.getvalue())
        files = [files[i % 32] for i in range(n)]
        bufs = [np.frombuffer(f, np.uint8) for f in files]
        # pinned copies of the files
        pinned = []
        for b in bufs:
            t = torch.empty(len(b), dtype=torch.uint8).pin_memory()
            t.numpy()[:] = b
            pinned.append(t)
        enc = (_lib.Encoded * n)()
        ptrs = (C.c_void_p * n)()
        outs = []
        for i in range(n):
            enc[i] = _lib.Encoded(pinned[i].data_ptr(), len(bufs[i]))
            t = torch.empty((S, S, 3), dtype=torch.uint8, device="cuda")
            outs.append(t)
            ptrs[i] = t.data_ptr()
        status = (C.c_int32 * n)()
        L = _lib.lib()
        for _ in range(3):
            st = L.retto_b200_decode_images(ctx.handle, enc, n, ptrs, status)
            assert st == 0, st
        ref = np.asarray(Image.open(io.BytesIO(files[0])).convert("RGB"))
        assert np.array_equal(outs[0].cpu().numpy(), ref) and np.array_equal(outs[n - 1].cpu().numpy(), np.asarray(Image.open(io.BytesIO(files[n - 1])).convert("RGB")))
        ctx.enable_kernel_timing(True)
        ctx.reset_kernel_times()
        K = 5
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            L.retto_b200_decode_images(ctx.handle, enc, n, ptrs, status)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
        kt = {k: round(v[1] / K, 4) for k, v in ctx.kernel_times().items() if k.startswith("jpeg")}
        ctx.enable_kernel_timing(False)
        out[tag] = {"pages": n, "size": S, "mean_file_bytes": float(np.mean([len(f) for f in files])), "wall_ms_incl_h2d_and_parse": round(dt * 1e3, 3),
                    "pages_per_s": round(n / dt, 1), "kernel_ms": kt, "kernel_sum_ms": round(sum(kt.values()), 4)}
        print(tag, json.dumps(out[tag]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/r02_bench_jpeg.json", "w"), indent=1)


if __name__ == "__main__":
    main()
