"""Side benchmarks for BASELINE.json configs[1] (DB postprocess only, 1024 maps 960x960) and configs[2] (CTC decode only,
16384 lines 40 x 6625).  bench.py measures the headline config; these numbers go to DESIGN.md / profiles/."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from retto_b200.api import Context
from tools.synth import gen_probmap, synth_dict_text

PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(ctx, fn, steps=5, warmup=3, kernel_times=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s = ctx.torch_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if kernel_times:
        ctx.enable_kernel_timing(True); ctx.reset_kernel_times()
    e0.record(s)
    for _ in range(steps):
        fn()
    e1.record(s)
    torch.cuda.synchronize()
    kt = ctx.kernel_times() if kernel_times else {}
    ctx.enable_kernel_timing(False)
    return e0.elapsed_time(e1) / steps, {k: v[1] / steps for k, v in kt.items() if v[0]}


def main():
    ctx = Context(0)
    # config 2
    uniq = [torch.from_numpy(gen_probmap(2000 + i, 960, 960)).cuda() for i in range(32)]
    maps = [uniq[i % 32].clone() for i in range(1024)]
    torch.cuda.synchronize()
    # the C-ABI call itself: descriptor table and result buffers built once, as a C / Rust caller holds them (the Python wrapper's
    # per-call ctypes table building is reported beside it)
    import ctypes as C
    from retto_b200._lib import DetPostDesc, Box
    n, cap = 1024, 1024 * 80
    descs = (DetPostDesc * n)(*[DetPostDesc(p.data_ptr(), 960, 960, 960, 960) for p in maps])
    status, offs, boxes = (C.c_int32 * n)(), (C.c_int32 * (n + 1))(), (Box * cap)()

    def abi():
        st = ctx._L.retto_b200_det_postprocess(ctx._h, descs, n, status, offs, boxes, cap)
        assert st == 0, st
    ms, _ = timed(ctx, abi, kernel_times=False)
    _, kt = timed(ctx, abi)
    ms_py, _ = timed(ctx, lambda: ctx.det_postprocess(maps, [(960, 960)] * 1024, max_boxes_total=cap), kernel_times=False)
    ab5, ab9 = 5.0 * 960 * 960 * 1024, 9.0 * 960 * 960 * 1024
    print(json.dumps({"config": "DB postprocess only: 1024 synthetic 960x960 prob maps", "ms": ms, "maps_per_s": 1024 / ms * 1e3,
                      "boxes": int(offs[n]), "ms_through_python_wrapper": ms_py, "kernel_sum_ms": sum(kt.values()),
                      "moved_bytes_5HW": ab5, "gbs_5HW": ab5 / ms / 1e6, "frac_of_hbm_peak_5HW": ab5 / ms / 1e6 / PEAK,
                      "survey_bytes_9HW": ab9, "gbs_9HW": ab9 / ms / 1e6, "frac_of_hbm_peak_9HW": ab9 / ms / 1e6 / PEAK, "kernels_ms": kt}))
    del maps, uniq
    torch.cuda.empty_cache()
    # config 3
    ctx.dict_load(synth_dict_text())
    N, T, C = 16384, 40, 6625
    x = torch.rand((N, T, C), device="cuda") * 1e-3
    win = torch.randint(1, C, (N, T), device="cuda")
    win[torch.rand((N, T), device="cuda") < 0.45] = 0
    x.scatter_(2, win.unsqueeze(-1), (0.5 + 0.5 * torch.rand((N, T), device="cuda")).unsqueeze(-1))
    torch.cuda.synchronize()
    ms, kt = timed(ctx, lambda: ctx.ctc_decode([x]))
    ab = 4.0 * N * T * C
    print(json.dumps({"config": "CTC decode only: 16k rec logits 48x320 lines, 6625 classes", "ms": ms, "lines_per_s": N / ms * 1e3,
                      "algorithmic_bytes": ab, "gbs": ab / ms / 1e6, "frac_of_hbm_peak": ab / ms / 1e6 / PEAK, "kernels_ms": kt,
                      "note": "ms includes D2H of strings/scores and host packing of 16384 strings in Python ctypes wrapper"}))


if __name__ == "__main__":
    main()
