"""GPU parity: DetProcessor::postprocess (K2..K6) vs the CPU oracle — bitmaps, labels, boxes bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["run-table", "pixel-planes", "run-table+concurrent-geometry"])
def ccl_path(request, monkeypatch):
    """every test of this module runs on both CCL formulations of csrc/db_post.cu — the run-table path (default) and the
    pixel-plane passes (its fallback) — and on the opt-in mode that runs unclip beside the score kernel on a second stream"""
    monkeypatch.delenv("RETTO_B200_PIXEL_CCL", raising=False)
    monkeypatch.delenv("RETTO_B200_GEOM_CONCURRENT", raising=False)
    if request.param == "pixel-planes":
        monkeypatch.setenv("RETTO_B200_PIXEL_CCL", "1")
    elif request.param == "run-table+concurrent-geometry":
        monkeypatch.setenv("RETTO_B200_GEOM_CONCURRENT", "1")
    return request.param


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _labels_ref(bitmap):
    """8-connected components, label = min linear index (cv2 as an independent labeller)."""
    import cv2
    n, lab = cv2.connectedComponents((bitmap > 0).astype(np.uint8), connectivity=8)
    h, w = bitmap.shape
    idx = np.arange(h * w, dtype=np.int64).reshape(h, w)
    out = np.full((h, w), -1, np.int32)
    mins = np.full(n, h * w, np.int64)
    np.minimum.at(mins, lab.ravel(), idx.ravel())
    fg = lab > 0
    out[fg] = mins[lab[fg]].astype(np.int32)
    return out


def _consistent_maps(seed0, n, h, w, **kw):
    """planted-rect maps whose sorted_boxes comparator is a consistent order (SURVEY hard part #6)"""
    from oracle import oracle as O
    from tools.synth import gen_probmap
    out, seed = [], seed0
    while len(out) < n:
        p = gen_probmap(seed, h, w, **kw)
        seed += 1000
        if not O.det_postprocess(p, h, w).comparator_inconsistent:
            out.append(p)
    return out


def _check_pages(ctx, probs, ori_hw=None, allow_inconsistent=False, diag=None):
    """asserted mode: the oracle on glibc trig (tests/conftest.py); `diag` = the libm_diag fixture for a second,
    counted-only run with the CUDA path's trig hooked into the oracle"""
    import torch
    from oracle import oracle as O
    assert O.libm_mode() == 0
    gs = [_t(p) for p in probs]
    torch.cuda.synchronize()
    ori = ori_hw or [p.shape for p in probs]
    out = ctx.det_postprocess(gs, ori)
    for i, p in enumerate(probs):
        ref = O.det_postprocess(p, ori[i][0], ori[i][1], want_bitmap=True)
        bm = ctx.fetch_bitmap(i, *p.shape)
        assert np.array_equal(bm, ref.bitmap), f"bitmap mismatch page {i}"
        lab = ctx.fetch_labels(i, *p.shape)
        assert np.array_equal(lab, _labels_ref(ref.bitmap)), f"label mismatch page {i}"
        boxes, scores = out.page(i)
        # (an inconsistent sorted_boxes comparator makes the REFERENCE's order depend on Rust's sort internals; the
        #  oracle and the CUDA path both use the n <= 20 insertion sort, so they still agree with each other)
        assert allow_inconsistent or not ref.comparator_inconsistent
        assert ref.status >= 0 and out.page_status[i] == 0
        assert boxes.shape == ref.boxes.shape, f"page {i}: {len(boxes)} vs {len(ref.boxes)} boxes"
        assert np.array_equal(boxes, ref.boxes), f"box mismatch page {i}"
        assert np.array_equal(scores.view(np.uint32), ref.scores.view(np.uint32)), f"score mismatch page {i}"
    if diag is not None:
        diag("det_post_tests", lambda: [O.det_postprocess(p, ori[i][0], ori[i][1]).boxes for i, p in enumerate(probs)],
             [out.page(i)[0] for i in range(len(probs))])
    return out


def test_single_rect(ctx):
    p = np.full((64, 96), 0.05, np.float32)
    p[20:40, 10:80] = 0.9
    out = _check_pages(ctx, [p])
    assert len(out.boxes) == 1


def test_empty_and_full(ctx):
    a = np.zeros((64, 64), np.float32)
    b = np.ones((96, 160), np.float32)
    out = _check_pages(ctx, [a, b])
    assert out.offsets[1] == 0


@pytest.mark.parametrize("seed", range(6))
def test_planted_rects_256(ctx, seed, libm_diag):
    from tools.synth import gen_probmap
    probs = sum([_consistent_maps(100 + seed * 4 + i, 1, 256, 256, k_range=(3, 10), wide_angle=(i % 2 == 1), border_touch_p=0.3) for i in range(4)], [])
    _check_pages(ctx, probs, diag=libm_diag)


def test_planted_rects_960_batch(ctx, libm_diag):
    from tools.synth import gen_probmap
    probs = sum([_consistent_maps(7 + i, 1, 960, 960) for i in range(6)], [])
    out = _check_pages(ctx, probs, diag=libm_diag)
    assert out.offsets[-1] > 60


def test_mixed_sizes_and_scaling(ctx):
    from tools.synth import gen_probmap
    shapes = [(736, 992), (1280, 1280), (96, 1504), (320, 132)]
    probs = [gen_probmap(50 + i, h, w, k_range=(2, 12)) for i, (h, w) in enumerate(shapes)]
    ori = [(480, 640), (1280, 1280), (64, 1000), (640, 260)]
    _check_pages(ctx, probs, ori)


def test_wide_angle_rects_glibc_and_diag(ctx, libm_diag):
    """wide-angle rectangles (the f64 trig of min_area_rect / Clipper matters most here): asserted against the glibc
    oracle like every other test; the diagnostic second run counts boxes that change with the CUDA path's trig"""
    from tools.synth import gen_probmap
    probs = sum([_consistent_maps(900 + i, 1, 512, 512, k_range=(5, 15), wide_angle=True) for i in range(8)], [])
    _check_pages(ctx, probs, diag=libm_diag)


def test_noise_speckles(ctx):
    rng = np.random.default_rng(3)
    p = (rng.random((200, 300)) < 0.02).astype(np.float32) * 0.9
    p[50:90, 40:200] = 0.8
    _check_pages(ctx, [p])


def test_hole_borders(ctx):
    """find_contours also returns hole borders (no border_type filter in retto): rings, tiny holes whose border
    box survives, nested shapes, holes next to the frame"""
    import cv2
    maps = []
    a = np.full((128, 192), 0.05, np.float32)
    a[10:60, 10:180] = 0.9
    a[25:45, 30:160] = 0.05            # big hole: hole contour fails the score filter
    a[80:110, 20:80] = 0.9
    a[93:96, 40:43] = 0.1              # 3x3 raw hole -> 2x2 after dilation: its border box survives
    a[93:96, 60:64] = 0.1
    maps.append(a)
    b = np.full((96, 96), 0.02, np.float32)
    b[2:94, 2:94] = 0.95               # frame-filling blob with several holes, one containing an island
    b[10:40, 10:40] = 0.0
    b[20:30, 20:30] = 0.9              # island inside the hole (its own outer border)
    b[50:54, 50:90] = 0.0
    b[70:73, 10:13] = 0.0
    maps.append(b)
    # a page-filling foreground whose first row starts at x = 0, with letters as holes (an inverted page): imageproc never starts its
    # outer border — every run start / end at x > 0 faces one of its own holes, whose border was traced first — so only the hole
    # borders yield contours
    c = np.full((120, 200), 0.9, np.float32)
    c[20:40, 30:60] = 0.05
    c[20:40, 80:130] = 0.05
    c[70:100, 50:150] = 0.05
    c[80:90, 90:110] = 0.9             # an island inside the third hole
    maps.append(c)
    # the same, but the blob leaves the frame further down (an L-shaped outside notch): the outer border IS discovered there, late
    d = c.copy()
    d[60:120, 170:200] = 0.05          # outside background reaching the right / bottom frame
    d[50:56, 0:20] = 0.05              # and a notch on the left frame
    maps.append(d)
    rng = np.random.default_rng(11)
    for _ in range(4):                 # random blobs: many irregular holes
        m = (rng.random((80, 120)) < 0.62).astype(np.uint8) * 255
        m = cv2.morphologyEx(m, cv2.MORPH_OPEN, np.ones((3, 3), np.uint8))
        maps.append(np.where(m > 0, 0.9, 0.05).astype(np.float32))
    ctx.enable_trace(True)
    out = _check_pages(ctx, maps, allow_inconsistent=True)
    from oracle import oracle as O
    for i, p in enumerate(maps):
        # true holes (textbook Suzuki-Abe typing; with the frame quirk some OUTER borders are merely typed "hole")
        n_holes_ref = sum(h for _, h in O.find_contours(O.threshold_dilate(p), quirk_x0=0))
        assert ctx.fetch_trace(i)["n_holes"] == n_holes_ref
    ctx.enable_trace(False)
    assert ctx.fetch_trace(0)["n_holes"] == 3 and len(out.page(0)[0]) == 3   # 2 outer boxes + 1 surviving hole-border box


def test_run_table_overflow_falls_back(ctx, ccl_path):
    """a page with more horizontal runs than ccl_runs_kernel keeps on chip (8192) sends the batch to the pixel-plane
    passes: same boxes, labels and bitmaps, also for the well-behaved page next to it"""
    from tools.synth import gen_probmap
    rng = np.random.default_rng(5)
    noisy = np.full((1024, 1024), 0.05, np.float32)
    ys, xs = rng.integers(2, 1020, 7000), rng.integers(2, 1020, 7000)
    noisy[ys, xs] = 0.9                     # isolated pixels -> 2x2 blobs after dilation: two runs each
    noisy[400:440, 100:700] = 0.85
    probs = [_consistent_maps(4242, 1, 512, 512, k_range=(5, 12))[0], noisy]
    out = _check_pages(ctx, probs, allow_inconsistent=True)
    assert len(out.page(1)[0]) >= 1


def test_no_spurious_hole_path(ctx, ccl_path):
    """pages without holes must not take the hole-border path: the Euler number (#components - #holes) has to come out
    right also for runs that cross the 128-px strips of the bitmap kernel (wide rectangles on a 1280-px page)"""
    import torch
    from oracle import oracle as O
    p = np.full((640, 1280), 0.05, np.float32)
    p[40:80, 30:1250] = 0.9          # crosses nine strips
    p[120:170, 100:700] = 0.85
    p[200:230, 127:129] = 0.9        # a thin bar on a strip boundary
    for k in range(12):              # a staircase: diagonal contacts between consecutive rows
        p[300 + 4 * k:304 + 4 * k, 200 + 37 * k:260 + 37 * k] = 0.9
    assert sum(h for _, h in O.find_contours(O.threshold_dilate(p), quirk_x0=0)) == 0
    g = _t(p)
    torch.cuda.synchronize()
    l0 = ctx.launch_count
    out = ctx.det_postprocess([g], [p.shape])
    used = ctx.launch_count - l0
    assert used <= {"run-table": 9, "run-table+concurrent-geometry": 10}.get(ccl_path, 12), used    # no bg_* / hole_* kernels
    ref = O.det_postprocess(p, *p.shape)
    assert np.array_equal(out.page(0)[0], ref.boxes)
