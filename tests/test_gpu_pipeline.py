"""GPU parity: rotate-crop (K7), cls/rec batch build (K8), cls postprocess + flip (K9) and the whole
RettoSession::process_pipeline against the CPU oracle pipeline — boxes, labels, strings bit-exact."""
import os
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


from _workers import NpWorker  # noqa: E402


@pytest.mark.parametrize("lazy", [False, True])
def test_crop_parity(ctx, lazy, monkeypatch):
    """lazy = the session's mode: direct crops (whole-pixel translations inside the page) are left to the batch build and only
    materialised when retto_b200_crop_fetch asks for them"""
    import torch
    from oracle import oracle as O
    if lazy:
        monkeypatch.setenv("RETTO_B200_CROP_LAZY", "1")
    else:
        monkeypatch.delenv("RETTO_B200_CROP_LAZY", raising=False)
    rng = np.random.default_rng(0)
    page = rng.integers(0, 256, (700, 900, 3), dtype=np.uint8)
    boxes = []
    # axis-aligned (translation class), rotated rectangles (projection class), tall (rotate270), near the border (white fill)
    boxes.append([[50, 60], [350, 60], [350, 100], [50, 100]])
    boxes.append([[100, 200], [500, 230], [497, 270], [97, 240]])
    boxes.append([[600, 100], [640, 100], [640, 400], [600, 400]])
    boxes.append([[0, 0], [200, 0], [200, 30], [0, 30]])
    boxes.append([[700, 650], [899, 655], [898, 699], [699, 694]])
    boxes.append([[300, 400], [420, 520], [390, 550], [270, 430]])
    for _ in range(10):
        cx, cy = rng.uniform(150, 750), rng.uniform(100, 600)
        w, h, a = rng.uniform(30, 280), rng.uniform(10, 60), rng.uniform(-0.5, 0.5)
        c, s = np.cos(a), np.sin(a)
        pts = [(-w / 2, -h / 2), (w / 2, -h / 2), (w / 2, h / 2), (-w / 2, h / 2)]
        boxes.append([[round(cx + x * c - y * s), round(cy + x * s + y * c)] for x, y in pts])
    boxes = np.array(boxes, np.float32)
    g = _t(page)
    torch.cuda.synchronize()
    infos = ctx.crop_boxes([g], [0] * len(boxes), boxes)
    for i, b in enumerate(boxes):
        ref = O.get_crop_img(page, b)
        cw, ch, rot = O.crop_dims(b)
        assert (infos[i].w, infos[i].h, infos[i].rotated270) == (cw, ch, rot)
        got = ctx.crop_fetch(i, infos[i])
        assert got.shape == ref.shape
        assert np.array_equal(got, ref), f"crop {i}: {np.abs(got.astype(int) - ref.astype(int)).max()} max diff"


def test_batches_and_cls_flip_parity(ctx):
    import torch
    from oracle import oracle as O
    from oracle.pipeline import stable_order_desc_ratio
    from retto_b200._lib import CropInfo, LineJob
    rng = np.random.default_rng(1)
    page = rng.integers(0, 256, (600, 1200, 3), dtype=np.uint8)
    boxes = []
    for k in range(15):
        x0, y0 = int(rng.integers(5, 300)), 10 + 38 * k
        w, h = int(rng.integers(40, 850)), int(rng.integers(12, 34))
        boxes.append([[x0, y0], [x0 + w, y0 + 1], [x0 + w, y0 + h], [x0, y0 + h - 1]])
    boxes.append([[1000, 100], [1030, 100], [1030, 300], [1000, 300]])  # tall -> rotate270
    boxes = np.array(boxes, np.float32)
    g = _t(page)
    torch.cuda.synchronize()
    infos = ctx.crop_boxes([g], [0] * len(boxes), boxes)
    crops = [O.get_crop_img(page, b) for b in boxes]
    dims = [c.shape[:2] for c in crops]
    order = stable_order_desc_ratio(dims)
    # cls batches
    lines, batches, total = ctx.plan_batches(0, infos)
    assert [l.crop for l in lines] == order
    base = ctx.build_batches(0, lines, total)
    host = np.zeros(total, np.float32)
    ctx._check(ctx._L.retto_b200_d2h(ctx._h, host.ctypes.data, base, total * 4))
    ctx.sync()
    for l in lines:
        ref = O.resize_norm_image(crops[l.crop], (3, 48, 192), None)
        got = host[l.dst_offset:l.dst_offset + 3 * 48 * 192].reshape(3, 48, 192)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"cls line crop {l.crop}"
    # cls postprocess: flip crops 0, 3, 7 (label 180 & score >= .9), not 5 (score < .9), not 2 (label 0)
    n = len(lines)
    logits = np.tile(np.array([[0.8, 0.2]], np.float32), (n, 1))
    pos = {l.crop: k for k, l in enumerate(lines)}
    for c in (0, 3, 7):
        logits[pos[c]] = (0.05, 0.95)
    logits[pos[5]] = (0.2, 0.8)
    logits[pos[2]] = (0.97, 0.03)
    logits[pos[9]] = (0.5, 0.5)   # tie -> first max -> label 0
    gl = _t(logits)
    torch.cuda.synchronize()
    res = ctx.cls_postprocess(gl, [l.crop for l in lines])
    st, am, sc = O.cls_postprocess(logits)
    for k in range(n):
        assert res[k][0] == (0, 180)[am[k]] and np.float32(res[k][1]) == sc[k]
    flipped = {c: (c in (0, 3, 7)) for c in range(len(crops))}
    for c in (0, 2, 3):
        got = ctx.crop_fetch(c, infos[c])
        ref = crops[c][::-1, ::-1] if flipped[c] else crops[c]
        assert np.array_equal(got, ref)
    # rec batches (running max_wh_ratio, flips applied as an index transform)
    lines, batches, total = ctx.plan_batches(1, infos)
    assert [l.crop for l in lines] == order
    base = ctx.build_batches(1, lines, total)
    host = np.zeros(total, np.float32)
    ctx._check(ctx._L.retto_b200_d2h(ctx._h, host.ctypes.data, base, total * 4))
    ctx.sync()
    mx = np.float32(320) / np.float32(48)
    for b in batches:
        for k in range(b.n):
            i = lines[b.first_line + k].crop
            wh = np.float32(dims[i][1]) / np.float32(dims[i][0])
            mx = max(mx, wh)
        assert np.float32(b.max_wh_ratio) == mx
        for k in range(b.n):
            l = lines[b.first_line + k]
            ref = O.resize_norm_image(crops[l.crop], (3, 48, 320), float(mx), flip180=flipped[l.crop])
            assert ref.shape[2] == b.img_w == l.img_w
            got = host[l.dst_offset:l.dst_offset + 3 * 48 * b.img_w].reshape(3, 48, b.img_w)
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"rec line crop {l.crop}"
    assert len({b.img_w for b in batches}) > 1


@pytest.mark.parametrize("generic,lazy,page_w", [(False, False, 2000), (True, False, 2000), (False, True, 2000), (False, True, 2001), (True, True, 2003)])
def test_batches_dims_matrix(ctx, generic, lazy, page_w, monkeypatch):
    """K8 over a matrix of crop sizes that reaches every class of build_batches_kernel (rec_batch.cu): up-scaling on
    both axes (FF), block means (BB: lines taller than 48 px), wide lines squeezed into 192 columns (BF), mixed /
    very large windows (generic) — each bit-exact against the oracle's resize_norm_image, with and without the
    180-degree flip, and once more with every line forced through the generic thumbnail_pixel path.
    lazy = the session's fused mode: these axis-aligned boxes are "direct" crops, so the batch build reads the 3-byte page pixels
    itself (every byte phase: page widths 2000 / 2001 / 2003) instead of a materialised RGBX copy."""
    import torch
    if lazy:
        monkeypatch.setenv("RETTO_B200_CROP_LAZY", "1")
    else:
        monkeypatch.delenv("RETTO_B200_CROP_LAZY", raising=False)
    from oracle import oracle as O
    from oracle.pipeline import stable_order_desc_ratio
    if generic:
        monkeypatch.setenv("RETTO_B200_BB_GENERIC", "1")
    else:
        monkeypatch.delenv("RETTO_B200_BB_GENERIC", raising=False)
    rng = np.random.default_rng(11)
    page = rng.integers(0, 256, (1500, page_w, 3), dtype=np.uint8)
    boxes = []
    for h in (6, 20, 31, 47, 48, 49, 60, 76, 97, 130, 200, 420):
        for w in (10, 40, 100, 191, 192, 193, 400, 777, 1100, 1900):
            x0, y0 = int(rng.integers(3, 1996 - w)), int(rng.integers(3, 1496 - h))
            boxes.append([[x0, y0], [x0 + w, y0], [x0 + w, y0 + h], [x0, y0 + h]])
    boxes = np.array(boxes, np.float32)
    g = _t(page)
    torch.cuda.synchronize()
    infos = ctx.crop_boxes([g], [0] * len(boxes), boxes)
    crops = [O.get_crop_img(page, b) for b in boxes]
    dims = [c.shape[:2] for c in crops]
    order = stable_order_desc_ratio(dims)
    lines, batches, total = ctx.plan_batches(0, infos)
    assert [l.crop for l in lines] == order
    base = ctx.build_batches(0, lines, total)
    host = np.zeros(total, np.float32)
    ctx._check(ctx._L.retto_b200_d2h(ctx._h, host.ctypes.data, base, total * 4))
    ctx.sync()
    for l in lines:
        ref = O.resize_norm_image(crops[l.crop], (3, 48, 192), None)
        got = host[l.dst_offset:l.dst_offset + 3 * 48 * 192].reshape(3, 48, 192)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"cls line crop {l.crop} dims {dims[l.crop]}"
    # flip every third crop through the cls postprocess, then build the rec batches
    n = len(lines)
    logits = np.tile(np.array([[0.8, 0.2]], np.float32), (n, 1))
    flipped = {c: (c % 3 == 0) for c in range(n)}
    for k, l in enumerate(lines):
        if flipped[l.crop]:
            logits[k] = (0.05, 0.95)
    gl = _t(logits)
    torch.cuda.synchronize()
    ctx.cls_postprocess(gl, [l.crop for l in lines])
    lines, batches, total = ctx.plan_batches(1, infos)
    base = ctx.build_batches(1, lines, total)
    host = np.zeros(total, np.float32)
    ctx._check(ctx._L.retto_b200_d2h(ctx._h, host.ctypes.data, base, total * 4))
    ctx.sync()
    for b in batches:
        for k in range(b.n):
            l = lines[b.first_line + k]
            ref = O.resize_norm_image(crops[l.crop], (3, 48, 320), float(np.float32(b.max_wh_ratio)), flip180=flipped[l.crop])
            assert ref.shape[2] == b.img_w == l.img_w
            got = host[l.dst_offset:l.dst_offset + 3 * 48 * b.img_w].reshape(3, 48, b.img_w)
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"rec line crop {l.crop} dims {dims[l.crop]} flip {flipped[l.crop]}"


def _session(ctx, worker, synth_dict):
    from retto_b200.session import CallableWorker, RettoSession
    ctx.dict_load(synth_dict)
    return RettoSession(worker=CallableWorker(worker.det, worker.cls, worker.rec), ctx=ctx)


@pytest.mark.parametrize("seed,hw", [(4, (1280, 1280)), (5, (960, 960)), (6, (736, 1000))])
def test_session_matches_oracle_pipeline(ctx, synth_dict, seed, hw):
    from oracle import oracle as O
    from oracle.pipeline import run_page
    from tools.synth import gen_page, probmap_from_rects
    h, w = hw
    img, rects = gen_page(seed, h, w)
    dh, dw = O.resize_either_plan(h, w)
    sx, sy = dw / w, dh / h
    prob = probmap_from_rects(seed, [(r[0] * sx, r[1] * sy, r[2] * sx, r[3] * sy, r[4]) for r in rects], dh, dw)
    wk = NpWorker(prob)
    taps = {}
    ref = run_page(img, wk, synth_dict, taps=taps)
    assert not taps["det"].comparator_inconsistent
    sess = _session(ctx, wk, synth_dict)
    got = sess.run(img)
    assert got.status == 0
    assert len(got.det_result) == len(ref["boxes"]) > 5
    for i in range(len(ref["boxes"])):
        assert np.array_equal(got.det_result[i].boxes, ref["boxes"][i])
        assert np.float32(got.det_result[i].score) == ref["scores"][i]
        assert (got.cls_result[i].label, np.float32(got.cls_result[i].score)) == (ref["cls"][i][0], np.float32(ref["cls"][i][1]))
        assert got.rec_result[i].text == ref["rec"][i][0]
        a, b = np.float32(got.rec_result[i].score), np.float32(ref["rec"][i][1])
        assert a == b or (np.isnan(a) and np.isnan(b))
    # the worker saw bit-identical tensors on both sides
    seen = sess.worker.seen
    assert np.array_equal(seen[0][0].view(np.uint32), taps["det_in"].view(np.uint32))
    assert len(seen[1]) == len(taps["cls_batches"]) and len(seen[2]) == len(taps["rec_batches"])
    for a, b in zip(seen[1] + seen[2], taps["cls_batches"] + taps["rec_batches"]):
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert any(taps["flipped"])


def test_session_batch_of_pages_with_resizes(ctx, synth_dict):
    """several pages in one call, including resize_both (>2000 px) and resize_either up-scaling (<736 px)"""
    from oracle import oracle as O
    from oracle.pipeline import run_page
    from tools.synth import gen_page, probmap_from_rects
    specs = [(11, 900, 1400), (12, 2300, 1700), (13, 500, 640), (14, 1280, 1280)]
    imgs, refs, workers = [], [], []
    for seed, h, w in specs:
        img, rects = gen_page(seed, h, w, n_lines=(6, 14))
        page = O.resize_both(img)
        ah, aw = page.shape[:2]
        dh, dw = O.resize_either_plan(ah, aw)
        sx, sy = dw / w, dh / h
        prob = probmap_from_rects(seed, [(r[0] * sx, r[1] * sy, r[2] * sx, r[3] * sy, r[4]) for r in rects], dh, dw)
        wk = NpWorker(prob)
        imgs.append(img)
        workers.append(wk)
        refs.append(run_page(img, wk, synth_dict))

    class Multi:
        def __init__(self):
            self.k = 0

        def det(self, x):
            w = workers[self.k]
            self.k += 1
            return w.det(x)

        def cls(self, x):
            return workers[0].cls(x)

        def rec(self, x):
            return workers[0].rec(x)

    sess = _session(ctx, Multi(), synth_dict)
    got = sess.run_pages(imgs)
    for g, r in zip(got, refs):
        assert g.status == 0 and len(g.det_result) == len(r["boxes"]) > 0
        for i in range(len(r["boxes"])):
            assert np.array_equal(g.det_result[i].boxes, r["boxes"][i])
            assert g.cls_result[i].label == r["cls"][i][0]
            assert g.rec_result[i].text == r["rec"][i][0]


def test_session_empty_page(ctx, synth_dict):
    """no detections -> cls/rec run zero batches and return empty vectors (chunks of an empty slice)"""
    wk = NpWorker(np.zeros((736, 736), np.float32))
    sess = _session(ctx, wk, synth_dict)
    r = sess.run(np.full((736, 736, 3), 255, np.uint8))
    assert r.status == 0 and r.det_result == [] and r.cls_result == [] and r.rec_result == []


from tools.demo_worker import StatelessWorker  # noqa: E402  (forwards that depend on the input tensor only)


def _small_pages(n, seed):
    rng = np.random.default_rng(seed)
    imgs = []
    for i in range(n):
        im = np.full((160 + 8 * (i % 5), 240 + 16 * (i % 3), 3), 255, np.uint8)
        for k in range(int(rng.integers(1, 4))):
            y0, x0 = 20 + 40 * k, int(rng.integers(10, 60))
            im[y0:y0 + int(rng.integers(10, 22)), x0:x0 + int(rng.integers(60, 150))] = int(rng.integers(0, 60))
        imgs.append(im)
    return imgs


def _same_results(a, b):
    assert len(a.det_result) == len(b.det_result)
    for x, y in zip(a.det_result, b.det_result):
        assert np.array_equal(x.boxes, y.boxes) and x.score == y.score
    assert [(c.label, c.score) for c in a.cls_result] == [(c.label, c.score) for c in b.cls_result]
    assert [r.text for r in a.rec_result] == [r.text for r in b.rec_result]
    assert [r.score for r in a.rec_result] == [r.score for r in b.rec_result] or all(
        (x.score == y.score) or (x.score != x.score and y.score != y.score) for x, y in zip(a.rec_result, b.rec_result))


def test_session_lanes_equal_serial(ctx, synth_dict):
    """device-resident pages: units of 16 pages software-pipelined on two lanes (csrc/session.cu) give exactly the
    results of the same batch run as one unit on one stream"""
    import torch
    imgs = _small_pages(75, 9)
    dev = [torch.from_numpy(im).cuda() for im in imgs]
    torch.cuda.synchronize()
    sess = _session(ctx, StatelessWorker(), synth_dict)
    try:
        ctx.set_pipeline(1, 1 << 20)
        serial = sess.run_pages(dev, on_device=True)
        ctx.set_pipeline(2, 16)
        l0 = ctx.launch_count
        piped = sess.run_pages(dev, on_device=True)
        assert ctx.launch_count - l0 > 5 * 20   # five units, each with its own kernel chain
    finally:
        ctx.set_pipeline(0, 0)
    assert len(serial) == len(piped) == 75 and sum(len(r.det_result) for r in serial) > 75
    for a, b in zip(serial, piped):
        _same_results(a, b)


def test_session_device_crop_table_paths(synth_dict, monkeypatch):
    """csrc/crop.cu builds the crop descriptor table on the device and enqueues the crop kernels before the host has the
    boxes; a batch that does not fit the capacities of the moment (table entries, pixel bytes, row grid) must fall back to the
    host-built table.  A fresh context sees: a small batch (fallback: nothing sized yet), a much larger one (fallback again:
    more rows / bytes than the hints), the same large one (device path), the small one (device path, oversized grid) — all
    four must equal the host-table results."""
    import torch
    from retto_b200.api import Context
    small, large = _small_pages(6, 21), _small_pages(90, 22)
    dev_s = [torch.from_numpy(im).cuda() for im in small]
    dev_l = [torch.from_numpy(im).cuda() for im in large]
    torch.cuda.synchronize()
    monkeypatch.setenv("RETTO_B200_HOST_CROP_TABLE", "1")
    c0 = Context(0)
    try:
        s0 = _session(c0, StatelessWorker(), synth_dict)
        ref_s, ref_l = s0.run_pages(dev_s, on_device=True), s0.run_pages(dev_l, on_device=True)
    finally:
        c0.close()
    assert sum(len(r.det_result) for r in ref_l) > 90
    monkeypatch.delenv("RETTO_B200_HOST_CROP_TABLE")
    c1 = Context(0)
    try:
        s1 = _session(c1, StatelessWorker(), synth_dict)
        for dev, ref in [(dev_s, ref_s), (dev_l, ref_l), (dev_l, ref_l), (dev_s, ref_s), (dev_l, ref_l)]:
            got = s1.run_pages(dev, on_device=True)
            assert len(got) == len(ref)
            for a, b in zip(got, ref):
                _same_results(a, b)
    finally:
        c1.close()


@pytest.mark.parametrize("pinned", [False, True])
def test_session_chunked_pipeline_equals_unchunked(ctx, synth_dict, pinned):
    """more host pages than one chunk take the chunked H2D/compute pipeline (csrc/session.cu): results must equal
    page-by-page runs.  pinned pages are pulled by pull_pages_kernel, pageable ones by cudaMemcpyAsync."""
    rng = np.random.default_rng(8)
    imgs = []
    for i in range(70):
        im = np.full((160 + 8 * (i % 5), 240 + 16 * (i % 3), 3), 255, np.uint8)
        for k in range(int(rng.integers(1, 4))):
            y0, x0 = 20 + 40 * k, int(rng.integers(10, 60))
            im[y0:y0 + int(rng.integers(10, 22)), x0:x0 + int(rng.integers(60, 150))] = int(rng.integers(0, 60))
        if pinned:
            import torch
            im = torch.from_numpy(im).pin_memory().numpy()
        imgs.append(im)

    sess = _session(ctx, StatelessWorker(), synth_dict)
    batch = sess.run_pages(imgs)
    assert len(batch) == 70 and sum(len(r.det_result) for r in batch) > 70
    for i in (0, 1, 63, 64, 65, 69):
        one = sess.run(imgs[i])
        assert len(one.det_result) == len(batch[i].det_result)
        for a, b in zip(one.det_result, batch[i].det_result):
            assert np.array_equal(a.boxes, b.boxes) and a.score == b.score
        assert [c.label for c in one.cls_result] == [c.label for c in batch[i].cls_result]
        assert [r.text for r in one.rec_result] == [r.text for r in batch[i].rec_result]


def test_cli_entry_point(ctx, synth_dict, tmp_path, capsys):
    """retto-cli surface (main.rs:18-95): files in a directory -> the three result lines per image + the summary line;
    the JSON objects equal a direct RettoSession run on the decoded pages"""
    import json
    from PIL import Image
    from retto_b200 import cli
    imgs = _small_pages(5, 21)
    (tmp_path / "imgs" / "sub").mkdir(parents=True)
    names = ["imgs/a.png", "imgs/b.png", "imgs/sub/c.png", "imgs/sub/d.png", "imgs/z.png"]
    for n, im in zip(names, imgs):
        Image.fromarray(im).save(tmp_path / n)
    (tmp_path / "keys.txt").write_text(synth_dict, encoding="utf-8")
    rc = cli.main(["-i", str(tmp_path / "imgs"), "--device", "b200", "--rec-keys-path", str(tmp_path / "keys.txt"),
                   "--worker", "tools.demo_worker:make_worker", "--batch-pages", "2", "--json"])
    assert rc == 0
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "Found 5 files, processing..." and out[-1].startswith("Successfully processed 5 images, avg time: ")
    assert sum(l.startswith("Det result: DetProcessorResult([") for l in out) == 5
    assert sum(l.startswith("Cls result: ClsProcessorResult([") for l in out) == 5
    assert sum(l.startswith("Rec result: RecProcessorResult([") for l in out) == 5
    js = [json.loads(l) for l in out if l.startswith("{")]
    files = sorted(str(tmp_path / n) for n in names)
    assert [j["file"] for j in js] == files
    sess = _session(ctx, StatelessWorker(), synth_dict)
    for j in js:
        im = imgs[names.index(os.path.relpath(j["file"], tmp_path))]
        want = sess.run(im).to_json()
        assert len(want["det_result"]) > 0
        assert json.dumps(want, sort_keys=True) == json.dumps({k: j[k] for k in ("det_result", "cls_result", "rec_result")}, sort_keys=True)


def test_cli_jpeg_files_and_debug_shapes(ctx, synth_dict, tmp_path, capsys):
    """JPEG files reach the device as file bytes (decoded by csrc/jpeg_decode.cu), PNG files are decoded on the host: both give
    the results of a session run on the libjpeg-turbo / Pillow pixels; the log lines have the reference's Debug shapes
    (PointBox {tl, tr, br, bl}, OrderedFloat(..), points.rs:70-82); `--worker standin` runs torch networks through the seam"""
    import json
    from PIL import Image
    from retto_b200 import cli
    imgs = _small_pages(4, 31)
    (tmp_path / "imgs").mkdir()
    for k, im in enumerate(imgs):
        if k % 2 == 0:
            Image.fromarray(im).save(tmp_path / "imgs" / f"p{k}.jpg", quality=92, restart_marker_rows=1)
        else:
            Image.fromarray(im).save(tmp_path / "imgs" / f"p{k}.png")
    (tmp_path / "keys.txt").write_text(synth_dict, encoding="utf-8")
    rc = cli.main(["-i", str(tmp_path / "imgs"), "--device", "b200", "--rec-keys-path", str(tmp_path / "keys.txt"),
                   "--worker", "tools.demo_worker:make_worker", "--json"])
    cap = capsys.readouterr()
    assert rc == 0 and "2 JPEG files decoded on the device, 2 files decoded on the host" in cap.err
    out = cap.out.splitlines()
    det = [l for l in out if l.startswith("Det result: ")]
    assert len(det) == 4 and "PointBox { tl: Point { x: OrderedFloat(" in det[0] and "inner" not in det[0]
    js = {os.path.basename(j["file"]): j for j in (json.loads(l) for l in out if l.startswith("{"))}
    sess = _session(ctx, StatelessWorker(), synth_dict)
    for k in range(4):
        name = f"p{k}.jpg" if k % 2 == 0 else f"p{k}.png"
        px = np.asarray(Image.open(tmp_path / "imgs" / name).convert("RGB"))
        want = sess.run(px).to_json()
        assert len(want["det_result"]) > 0
        assert json.dumps(want, sort_keys=True) == json.dumps({q: js[name][q] for q in ("det_result", "cls_result", "rec_result")}, sort_keys=True)
    rc = cli.main(["-i", str(tmp_path / "imgs" / "p0.jpg"), "--device", "b200", "--rec-keys-path", str(tmp_path / "keys.txt"), "--worker", "standin"])
    assert rc == 0 and "Successfully processed 1 images" in capsys.readouterr().out


def test_stage_api_orders_with_torch_current_stream(ctx):
    """The stage wrappers read tensors produced on torch's current stream and hand back tensors torch consumes next, while the
    context enqueues on its own non-blocking stream: no host synchronisation is needed either side (Context._ordered).  The
    producer here is kept busy by a long matmul chain first, so an unordered call would read the page before it is written."""
    import torch
    from oracle import oracle as O
    rng = np.random.default_rng(77)
    page = rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    ref = O.det_preprocess(page, 0, 736)
    pinned = torch.from_numpy(page).pin_memory()
    a = torch.randn(4096, 4096, device="cuda")
    side = torch.cuda.Stream()
    for s in (torch.cuda.current_stream(), side):
        with torch.cuda.stream(s):
            g = torch.zeros((720, 1280, 3), dtype=torch.uint8, device="cuda")
            b = a
            for _ in range(12):
                b = (b @ a) * 1e-3                     # ~ms of queued work ahead of the page upload
            g.copy_(pinned, non_blocking=True)
            out = ctx.det_preprocess([g])[0]
            got = (out + b[0, 0] * 0).cpu().numpy()    # consumed by torch on the same stream, no sync in between
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_c_api_demo_runs(tmp_path):
    """examples/c_api_demo.c — plain C99 over include/retto_b200.h + libretto_b200.so only — runs one page through
    retto_b200_run_pages with a stand-in forward callback and finds its three lines"""
    import subprocess
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_cpu_abi import _build_c_demo
    exe = _build_c_demo(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "c_api_demo ok" in r.stdout, (r.returncode, r.stdout[-1500:], r.stderr[-1500:])
    assert "lines 3" in r.stdout and 'text "abcdef"' in r.stdout
