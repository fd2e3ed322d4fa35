"""CPU tests: the C-ABI library loads and exports every symbol include/retto_b200.h declares; host-side
planning entry points (no GPU needed) agree with the oracle's restatement of the reference rules."""
import ctypes as C
import os
import re

import numpy as np

from oracle import oracle as O
from oracle.pipeline import stable_order_desc_ratio
from retto_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "retto_b200.h")).read()
    declared = set(re.findall(r"\b(retto_b200_[a-z0-9_]+)\s*\(", hdr)) - {"retto_b200_forward_fn", "retto_b200_status"}
    L = _lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert L.retto_b200_abi_version() == 2


def test_rust_sys_crate_covers_the_whole_header():
    """rust/retto-b200-sys/src/lib.rs is generated from include/retto_b200.h (tools/gen_rust_sys.py): it is up to date, and every
    function, struct and status code the header declares appears in it (the extern block is complete, not a sample)"""
    import subprocess
    import sys
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"]).returncode == 0
    rs = open(os.path.join(ROOT, "rust", "retto-b200-sys", "src", "lib.rs")).read()
    hdr = open(os.path.join(ROOT, "include", "retto_b200.h")).read()
    for sym in _lib.EXPORTS:
        assert re.search(r"pub fn %s\(" % sym, rs), sym
    for st in re.findall(r"\} (retto_b200_\w+);", hdr):
        assert ("pub struct %s" % st) in rs or ("pub type %s" % st) in rs, st
    for code in re.findall(r"(RETTO_B200_ERR_\w+) = \d+", hdr):
        assert ("pub const %s:" % code) in rs, code
    assert 'pub type retto_b200_forward_fn = Option<unsafe extern "C" fn(' in rs and "pub type retto_b200_stage_fn" in rs


def test_config_defaults_match_reference():
    c = _lib.Config()
    _lib.lib().retto_b200_config_default(C.byref(c))
    assert (c.max_side_len, c.min_side_len) == (2000, 30)                                   # session.rs:33-34
    assert (c.det_limit_side_len, c.det_limit_type, c.det_min_mini_box_size) == (736, 0, 3)  # det_processor.rs:75-93
    assert np.float32(c.det_scale) == np.float32(1) / np.float32(255)
    assert (np.float32(c.det_thresh), np.float32(c.det_box_thresh), np.float32(c.det_unclip_ratio)) == (np.float32(.3), np.float32(.5), np.float32(1.6))
    assert list(c.cls_image_shape) == [3, 48, 192] and c.cls_batch_num == 6 and np.float32(c.cls_thresh) == np.float32(.9) and list(c.cls_label) == [0, 180]
    assert list(c.rec_image_shape) == [3, 48, 320] and c.rec_batch_num == 6
    assert c.det_dilation_2x2 == 1


def test_no_cuda_device_fails_loudly():
    """the product has no CPU fallback: creating a context without a GPU is an error, not a silent CPU path"""
    import torch
    if torch.cuda.is_available():
        return
    h = C.c_void_p()
    assert _lib.lib().retto_b200_create(0, None, C.byref(h)) == _lib.ERR_CUDA


def test_resize_plans_match_oracle():
    from retto_b200.api import resize_both_plan, resize_either_plan
    rng = np.random.default_rng(0)
    for _ in range(300):
        h, w = int(rng.integers(1, 5000)), int(rng.integers(1, 5000))
        assert resize_both_plan(h, w) == O.resize_both_plan(h, w)
        assert resize_either_plan(h, w) == O.resize_either_plan(h, w)
        assert resize_either_plan(h, w, 1, 960) == O.resize_either_plan(h, w, 1, 960)


def _plan(kind, dims):
    L = _lib.lib()
    cfg = _lib.Config()
    L.retto_b200_config_default(C.byref(cfg))
    n = len(dims)
    infos = (_lib.CropInfo * max(n, 1))(*[_lib.CropInfo(w, h, 0, 0, 0) for (h, w) in dims])
    lines = (_lib.LineJob * max(n, 1))()
    batches = (_lib.Batch * (n + 1))()
    nb, tot = C.c_int32(), C.c_uint64()
    assert L.retto_b200_plan_batches(C.byref(cfg), kind, infos, n, lines, batches, C.byref(nb), C.byref(tot)) == 0
    return [lines[i] for i in range(n)], [batches[i] for i in range(nb.value)], tot.value


def test_plan_batches_follows_cls_and_rec_rules():
    rng = np.random.default_rng(1)
    for trial in range(50):
        n = int(rng.integers(0, 40))
        dims = [(int(rng.integers(8, 60)), int(rng.integers(8, 900))) for _ in range(n)]
        if n > 4:
            dims[3] = dims[1]                       # equal ratios: stable order keeps detection order
        order = stable_order_desc_ratio(dims)
        for kind in (0, 1):
            lines, batches, tot = _plan(kind, dims)
            assert [l.crop for l in lines] == order
            assert len(batches) == (n + 5) // 6
            mx = np.float32(320) / np.float32(48)   # carried across batches, never reset (rec_processor.rs:227)
            off = 0
            for b in batches:
                idxs = order[b.first_line:b.first_line + b.n]
                if kind == 1:
                    for i in idxs:
                        mx = max(mx, np.float32(dims[i][1]) / np.float32(dims[i][0]))
                    assert np.float32(b.max_wh_ratio) == mx
                for k, i in enumerate(idxs):
                    iw, rw = O.resize_norm_plan(dims[i][0], dims[i][1], 48, 192 if kind == 0 else 320, float(mx) if kind == 1 else None)
                    l = lines[b.first_line + k]
                    assert (l.img_w, l.resized_w, b.img_w) == (iw, rw, iw)
                    assert l.dst_offset == off + k * 3 * 48 * iw
                assert b.offset == off
                off += b.n * 3 * 48 * b.img_w
            assert tot == off


def test_shard_indices_lpt():
    from retto_b200.shard import shard_indices
    sizes = [100, 1, 50, 50, 7, 7, 30, 2]
    parts = shard_indices(sizes, 3)
    assert sorted(sum(parts, [])) == list(range(8))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) == 100 and all(p == sorted(p) for p in parts)
    assert shard_indices([5, 5, 5, 5], 2) == [[0, 2], [1, 3]]


def test_cli_surface(tmp_path):
    """retto-cli flags (main.rs:18-39) are accepted under the same names, files are found like WalkDir does"""
    from retto_b200 import cli
    a = cli.build_parser().parse_args(["-i", str(tmp_path), "--det-model-path", "d.onnx", "--cls-model-path", "c.onnx", "--rec-model-path", "r.onnx",
                                       "--rec-keys-path", "k.txt", "--device", "b200", "--device-id", "1", "--gpus", "2"])
    assert (a.images, a.det_model_path, a.device, a.device_id, a.gpus) == (str(tmp_path), "d.onnx", "b200", 1, 2)
    assert cli.build_parser().prog == "ratio-cli"
    (tmp_path / "sub").mkdir()
    for n in ("b.png", "a.png", "sub/c.png"):
        (tmp_path / n).write_bytes(b"x")
    assert [os.path.relpath(f, tmp_path) for f in cli.find_files(str(tmp_path))] == ["a.png", "b.png", "sub/c.png"]
    assert cli.find_files(str(tmp_path / "a.png")) == [str(tmp_path / "a.png")]
    # the log lines use the reference's Debug shapes: custom PointBox impl (points.rs:70-82), OrderedFloat coordinates, f32 `{:?}`
    # digits, str escape_debug
    from retto_b200.session import ClsPostProcessLabel, DetProcessorInnerResult, RecProcessorSingleResult, RettoWorkerResult
    import numpy as np
    r = RettoWorkerResult([DetProcessorInnerResult(np.array([[1, 2], [30, 2], [30, 9], [1, 9]], np.float32), 0.8235294)],
                          [ClsPostProcessLabel(180, 0.95)], [RecProcessorSingleResult('a"b\\c\n中', float("nan"))])
    d, c, t = cli.fmt_debug(r)
    assert d == ("Det result: DetProcessorResult([DetProcessorInnerResult { boxes: PointBox { tl: Point { x: OrderedFloat(1.0), y: OrderedFloat(2.0) }, "
                 "tr: Point { x: OrderedFloat(30.0), y: OrderedFloat(2.0) }, br: Point { x: OrderedFloat(30.0), y: OrderedFloat(9.0) }, "
                 "bl: Point { x: OrderedFloat(1.0), y: OrderedFloat(9.0) } }, score: 0.8235294 }])")
    assert c == "Cls result: ClsProcessorResult([ClsProcessorSingleResult { label: ClsPostProcessLabel { label: 180, score: 0.95 } }])"
    assert t == 'Rec result: RecProcessorResult([RecProcessorSingleResult { text: "a\\"b\\\\c\\n中", score: NaN }])'
    assert cli._f32(1e-7) == "1e-7" and cli._f32(100000.0) == "100000.0"


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU oracle on the host cores) prints ONE JSON line with the driver's keys; it needs no GPU"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "4",
                          "--size", "640"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "impl", "cpu_baseline", "e2e"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "pages/s" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bind_host_to_gpu_is_harmless_without_nvml():
    from retto_b200.shard import bind_host_to_gpu
    before = os.sched_getaffinity(0)
    n = bind_host_to_gpu(0)
    assert isinstance(n, int) and n >= 0
    if n == 0:
        assert os.sched_getaffinity(0) == before


def _build_c_demo(tmp_path):
    import subprocess
    exe = str(tmp_path / "c_api_demo")
    lib_dir = os.path.join(ROOT, "retto_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_api_demo.c"),
           "-L" + lib_dir, "-lretto_b200", "-L/usr/local/cuda/lib64", "-Wl,-rpath-link,/usr/local/cuda/lib64", "-Wl,-rpath," + lib_dir, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_c99_and_the_c_demo_links(tmp_path):
    """the boundary is a C ABI: the header must compile as plain C99 (pedantic, warnings as errors), and a C program that uses nothing
    but the header and the .so (examples/c_api_demo.c: no CUDA headers, no torch) must link; without a CUDA device it fails loudly"""
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "retto_b200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exe = _build_c_demo(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode in (0, 2), (r.returncode, r.stdout, r.stderr)         # 2 = retto_b200_create found no CUDA device
    if r.returncode == 2:
        assert "retto_b200_create" in r.stderr


def test_design_md_is_the_generators_output(tmp_path):
    """DESIGN.md is generated (tools/fill_design.py: docs_src/DESIGN.md.in + the evidence under profiles/): the committed file must be what
    the generator writes from the committed evidence, so the numbers in the document cannot drift from the JSON they cite"""
    import shutil
    import subprocess
    import sys
    committed = open(os.path.join(ROOT, "DESIGN.md")).read()
    backup = tmp_path / "DESIGN.md"
    shutil.copy(os.path.join(ROOT, "DESIGN.md"), backup)
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fill_design.py")], capture_output=True, text=True, cwd=ROOT)
        assert r.returncode == 0, r.stderr
        assert open(os.path.join(ROOT, "DESIGN.md")).read() == committed
    finally:
        shutil.copy(backup, os.path.join(ROOT, "DESIGN.md"))
