"""GPU parity: DetProcessor::postprocess (K2..K6) vs the CPU oracle — bitmaps, labels, boxes bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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


def _check_pages(ctx, probs, ori_hw=None, libm=1):
    import torch
    from oracle import oracle as O
    O.set_libm(libm)
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
        assert not ref.comparator_inconsistent
        assert ref.status >= 0 and out.page_status[i] == 0
        assert boxes.shape == ref.boxes.shape, f"page {i}: {len(boxes)} vs {len(ref.boxes)} boxes"
        assert np.array_equal(boxes, ref.boxes), f"box mismatch page {i}"
        assert np.array_equal(scores.view(np.uint32), ref.scores.view(np.uint32)), f"score mismatch page {i}"
    O.set_libm(0)
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
def test_planted_rects_256(ctx, seed):
    from tools.synth import gen_probmap
    probs = sum([_consistent_maps(100 + seed * 4 + i, 1, 256, 256, k_range=(3, 10), wide_angle=(i % 2 == 1), border_touch_p=0.3) for i in range(4)], [])
    _check_pages(ctx, probs)


def test_planted_rects_960_batch(ctx):
    from tools.synth import gen_probmap
    probs = sum([_consistent_maps(7 + i, 1, 960, 960) for i in range(6)], [])
    out = _check_pages(ctx, probs)
    assert out.offsets[-1] > 60


def test_mixed_sizes_and_scaling(ctx):
    from tools.synth import gen_probmap
    shapes = [(736, 992), (1280, 1280), (96, 1504), (320, 132)]
    probs = [gen_probmap(50 + i, h, w, k_range=(2, 12)) for i, (h, w) in enumerate(shapes)]
    ori = [(480, 640), (1280, 1280), (64, 1000), (640, 260)]
    _check_pages(ctx, probs, ori)


def test_glibc_vs_rtmath_same_boxes(ctx):
    """the oracle in reference-faithful libm mode (glibc) gives the same boxes as the CUDA path here"""
    from tools.synth import gen_probmap
    probs = sum([_consistent_maps(900 + i, 1, 512, 512, k_range=(5, 15), wide_angle=True) for i in range(4)], [])
    _check_pages(ctx, probs, libm=0)


def test_noise_speckles(ctx):
    rng = np.random.default_rng(3)
    p = (rng.random((200, 300)) < 0.02).astype(np.float32) * 0.9
    p[50:90, 40:200] = 0.8
    _check_pages(ctx, [p])
