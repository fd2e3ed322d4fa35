"""CPU tests: the oracle against its golden vectors and against independent implementations on this box
(cv2 / numpy), as SURVEY.md §8(c) lists.  No GPU needed."""
import os
import zlib

import numpy as np
import pytest

from oracle import oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def test_normalise_formulas_bits():
    # det: (x as f32 * (1f32/255f32) - .5) / .5 ; cls/rec: (x as f32 / 255f32 - .5) / .5  
    x = np.arange(256, dtype=np.float32)
    det = ((x * (np.float32(1) / np.float32(255))) - np.float32(0.5)) / np.float32(0.5)
    rec = ((x / np.float32(255)) - np.float32(0.5)) / np.float32(0.5)
    assert np.array_equal(G["norm_det"].view(np.uint32), det.view(np.uint32))
    assert np.array_equal(G["norm_rec"].view(np.uint32), rec.view(np.uint32))
    assert 50 < int((det.view(np.uint32) != rec.view(np.uint32)).sum()) < 200   # the two really differ


def test_det_preprocess_layout_bgr():
    img = np.zeros((32, 32, 3), np.uint8)
    img[..., 0], img[..., 1], img[..., 2] = 10, 20, 30   # R,G,B
    t = O.det_preprocess(img, limit_len=32)
    assert t.shape == (1, 3, 32, 32)
    f = lambda v: (np.float32(v) * (np.float32(1) / np.float32(255)) - np.float32(.5)) / np.float32(.5)  # noqa
    assert t[0, 0, 0, 0] == f(30) and t[0, 1, 0, 0] == f(20) and t[0, 2, 0, 0] == f(10)   # planes are B,G,R


def test_thumbnail_golden_and_identity():
    for k in range(int(G["thumb_n"])):
        src, ref = G[f"thumb_in_{k}"], G[f"thumb_out_{k}"]
        assert np.array_equal(O.thumbnail(src, ref.shape[0], ref.shape[1]), ref)
    img = np.random.default_rng(0).integers(0, 256, (17, 23, 3), dtype=np.uint8)
    assert np.array_equal(O.thumbnail(img, 17, 23), img)
    # exact 2:1 box filter == rounded mean of 2x2 blocks
    img = np.random.default_rng(1).integers(0, 256, (8, 12, 3), dtype=np.uint8)
    ref = ((img.reshape(4, 2, 6, 2, 3).astype(np.uint32).sum((1, 3)) + 2) // 4).astype(np.uint8)
    assert np.array_equal(O.thumbnail(img, 4, 6), ref)


def test_resize_plans():
    assert O.resize_both_plan(960, 960) == []
    assert O.resize_both_plan(2896, 4096) == [(1408, 1984)]
    assert O.resize_both_plan(2000, 3000) == [(1312, 1984)]
    assert O.resize_both_plan(20, 200) == [(32, 288)]           # < min_side_len: scale 1.5 -> floor(30/32).round()*32
    assert O.resize_either_plan(480, 640) == (736, 992)
    assert O.resize_either_plan(1280, 1280) == (1280, 1280)
    assert O.resize_either_plan(2000, 1500) == (2016, 1504)            # 62.5 rounds half away from zero


def test_dilate_equals_cv2():
    import cv2
    rng = np.random.default_rng(0)
    for _ in range(10):
        m = (rng.random((40, 55)) > 0.8).astype(np.uint8) * 255
        pred = np.where(m > 0, 0.9, 0.1).astype(np.float32)
        assert np.array_equal(O.threshold_dilate(pred, 0.3, True), cv2.dilate(m, np.ones((2, 2), np.uint8)))
    p = np.array([[0.3, np.nan, 0.30000004]], np.float32)      # strict >, NaN -> 0
    assert O.threshold_dilate(p, 0.3, False).tolist() == [[0, 0, 255]]


def _sets(cs):
    return sorted(tuple(sorted(set(map(tuple, np.asarray(c).reshape(-1, 2).tolist())))) for c in cs)


def test_find_contours_point_sets_equal_cv2():
    import cv2
    rng = np.random.default_rng(2)
    for it in range(150):
        h, w = int(rng.integers(6, 50)), int(rng.integers(6, 50))
        m = (rng.random((h, w)) < rng.choice([0.05, 0.2, 0.4])).astype(np.uint8) * 255
        if it % 2:
            m = cv2.dilate(m, np.ones((2, 2), np.uint8))
        ref, _ = cv2.findContours(m, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
        assert _sets([c for c, _ in O.find_contours(m, quirk_x0=0)]) == _sets(ref)


def test_find_contours_order_holes_and_frame_quirk():
    m = np.zeros((12, 16), np.uint8)
    m[2:10, 3:13] = 255
    m[4:8, 6:10] = 0                       # ring: outer + hole border
    cs = O.find_contours(m)
    assert [h for _, h in cs] == [0, 1]
    assert cs[0][0][0].tolist() == [3, 2]  # discovered at the component's first pixel in raster order
    assert cs[1][0][0].tolist() == [5, 4]  # hole border starts at the pixel left of the hole's first pixel
    # imageproc's scan never starts a border at the frame itself (RECALLED quirk): a blob touching column 0 is
    # discovered at the right end of its first run, with the same point set
    m = np.zeros((8, 10), np.uint8)
    m[2:5, 0:4] = 255
    a, b = O.find_contours(m, quirk_x0=1), O.find_contours(m, quirk_x0=0)
    assert len(a) == len(b) == 1 and _sets([a[0][0]]) == _sets([b[0][0]])
    assert a[0][0][0].tolist() == [3, 2] and b[0][0][0].tolist() == [0, 2]
    full = np.full((6, 6), 255, np.uint8)
    assert len(O.find_contours(full, quirk_x0=1)) == 0 and len(O.find_contours(full, quirk_x0=0)) == 1


def test_convex_hull_equals_cv2_and_order():
    import cv2
    rng = np.random.default_rng(3)
    for _ in range(100):
        pts = rng.integers(0, 25, (int(rng.integers(1, 50)), 2)).astype(np.int32)
        mine = O.convex_hull(pts)
        ref = cv2.convexHull(pts.reshape(-1, 1, 2)).reshape(-1, 2)
        assert set(map(tuple, mine.tolist())) == set(map(tuple, ref.tolist()))
        start = min(map(tuple, pts.tolist()), key=lambda p: (p[1], p[0]))
        assert tuple(mine[0]) == start
    sq = O.convex_hull(np.array([[0, 0], [4, 0], [4, 3], [0, 3], [2, 1], [2, 0]], np.int32))
    assert sq.tolist() == [[0, 0], [4, 0], [4, 3], [0, 3]]     # top-left, then along the top, collinear dropped


def test_min_area_rect_axis_aligned_and_degenerate():
    r = O.min_area_rect(np.array([[10, 10], [50, 10], [50, 20], [10, 20], [30, 15]], np.int32))
    assert r.tolist() == [[10, 10], [50, 10], [50, 20], [10, 20]]
    assert O.min_area_rect(np.array([[3, 4]], np.int32)).tolist() == [[3, 4]] * 4
    assert O.min_area_rect(np.array([[1, 1], [5, 5], [3, 3]], np.int32)).tolist() == [[1, 1], [5, 5], [5, 5], [1, 1]]


def _row_cover_py(px, py, bw, bh, y):
    """Python port of the CUDA path's closed-form row coverage (db_geom.cuh polygon_row_cover)"""
    cover = set()
    inter = []
    for e in range(4):
        x0, y0, x1, y1 = px[e], py[e], px[(e + 1) % 4], py[(e + 1) % 4]
        if (y0 <= y <= y1) or (y1 <= y <= y0):
            if y0 == y1:
                inter += [x0, x1]
            elif y0 == y or y1 == y:
                if y1 > y:
                    inter.append(x0)
                if y0 > y:
                    inter.append(x1)
            else:
                fr = np.float32(y - y0) / np.float32(y1 - y0)
                v = np.float32(x0) + fr * np.float32(x1 - x0)
                inter.append(int(np.sign(v) * np.floor(np.abs(v) + np.float32(0.5))))
    inter.sort()
    for k in range(0, len(inter) - 1, 2):
        f, t = min(inter[k], bw), min(inter[k + 1], bw - 1)
        if f < bw and t >= 0:
            cover.update(range(max(0, f), max(0, t) + 1))
    for e in range(4):
        x0, y0, x1, y1 = px[e], py[e], px[(e + 1) % 4], py[(e + 1) % 4]
        steep = abs(y1 - y0) > abs(x1 - x0)
        if steep:
            x0, y0, x1, y1 = y0, x0, y1, x1
        if x0 > x1:
            x0, x1, y0, y1 = x1, x0, y1, y0
        dx, dy = x1 - x0, abs(y1 - y0)
        ys = 1 if y0 < y1 else -1
        if steep:
            if y < x0 or y > x1:
                continue
            k = y - x0
            A = 2 * k * dy - dx
            nk = 0 if (A <= 0 or dx == 0) else (A + 2 * dx - 1) // (2 * dx)
            xx = y0 + ys * nk
            if 0 <= xx < bw:
                cover.add(xx)
        else:
            t = (y - y0) * ys
            if t < 0 or t > dy:
                continue
            if dy == 0:
                klo, khi = 0, dx
            else:
                klo = 0 if t == 0 else (dx * (2 * t - 1)) // (2 * dy) + 1
                khi = min(dx, (dx * (2 * t + 1)) // (2 * dy))
            for xx in range(x0 + klo, x0 + khi + 1):
                if 0 <= xx < bw:
                    cover.add(xx)
    return cover


def test_polygon_mask_closed_form_matches_scan_fill_plus_bresenham():
    """the per-row interval formulation used on the GPU reproduces draw_polygon_mut's mask exactly"""
    rng = np.random.default_rng(5)
    for it in range(300):
        bw, bh = int(rng.integers(4, 60)), int(rng.integers(4, 40))
        if it % 3 == 0:   # rotated-rectangle-like quads, partly outside the canvas
            q = rng.integers(-6, [bw + 6, bh + 6], (4, 2))
        else:
            cx, cy, w, h, a = rng.uniform(0, bw), rng.uniform(0, bh), rng.uniform(2, bw), rng.uniform(1, bh), rng.uniform(-1.6, 1.6)
            c, s = np.cos(a), np.sin(a)
            q = np.array([[cx + x * c - y * s, cy + x * s + y * c] for x, y in [(-w / 2, -h / 2), (w / 2, -h / 2), (w / 2, h / 2), (-w / 2, h / 2)]])
            q = np.round(q).astype(int)
        if (q[0] == q[3]).all():
            continue
        st, mask = O.polygon_mask(bw, bh, q)
        assert st == 0
        px, py = [int(v) for v in q[:, 0]], [int(v) for v in q[:, 1]]
        for y in range(bh):
            assert _row_cover_py(px, py, bw, bh, y) == set(np.nonzero(mask[y])[0].tolist()), (it, y, q.tolist())
    assert O.polygon_mask(8, 8, np.array([[1, 1], [5, 1], [5, 1], [1, 1]]))[0] == -1   # reference panics: first == last


def test_box_score_is_masked_mean():
    pred = np.random.default_rng(7).random((30, 40)).astype(np.float32)
    q = np.array([[5, 4], [30, 6], [29, 20], [4, 18]], np.int32)
    st, sc = O.box_score_fast(pred, q)
    _, mask = O.polygon_mask(30 - 4 + 1, 20 - 4 + 1, q - np.array([4, 4]))
    ref = np.float32(0)
    for v, m in zip(pred[4:21, 4:31].ravel(), mask.ravel()):
        ref = np.float32(ref + v * np.float32(m))
    assert st == 0 and np.float32(sc) == np.float32(ref / np.float32(mask.sum()))


def test_unclip_offset_properties():
    q = np.array([[10, 10], [110, 10], [110, 40], [10, 40]], np.int32)
    pts, d = O.unclip(q, 1.6)
    assert abs(d - 100 * 30 * 1.6 / 260) < 1e-4
    assert pts[:, 0].min() == round(10 - d) and pts[:, 0].max() == round(110 + d)
    assert pts[:, 1].min() == round(10 - d) and pts[:, 1].max() == round(40 + d)
    r = O.min_area_rect(pts)
    assert r.tolist() == [[round(10 - d), round(10 - d)], [round(110 + d), round(10 - d)], [round(110 + d), round(40 + d)], [round(10 - d), round(40 + d)]]


@pytest.mark.parametrize("libm", [0, 1])
def test_det_postprocess_golden(libm):
    """boxes, scores, per-contour trace; identical in both libm modes (glibc vs the CUDA path's rt_fmath)"""
    O.set_libm(libm)
    try:
        for k in range(int(G["det_n"])):
            p = G[f"det_prob_{k}"]
            r = O.det_postprocess(p, 96, 128, want_bitmap=True)
            assert zlib.crc32(r.bitmap.tobytes()) == int(G[f"det_bitmap_crc_{k}"])
            assert np.array_equal(r.boxes, G[f"det_boxes_{k}"])
            assert np.array_equal(r.scores.view(np.uint32), G[f"det_scores_{k}"].view(np.uint32))
            rect1, ss, sc, st = O.det_trace(p, 96, 128)
            assert np.array_equal(rect1, G[f"det_rect1_{k}"]) and np.array_equal(st, G[f"det_status_{k}"])
    finally:
        O.set_libm(0)


def test_det_ring_yields_hole_box_and_diagonal_merge():
    k_ring, k_diag = int(G["det_n"]) - 2, int(G["det_n"]) - 1
    # App. A #4: the hole border is a contour of its own (no border_type filter); here its mean score over the
    # mostly-background hole fails box_thresh (status 2) while the outer border is kept (status 0)
    assert int(G[f"det_ncontours_{k_ring}"]) == 2 and G[f"det_status_{k_ring}"].tolist() == [0, 2]
    assert int(G[f"det_ncontours_{k_diag}"]) == 1                                            # 8-connectivity


def test_scale_and_clip_rounding():
    b = np.array([[0.5, 1.5], [2.5, -3], [99.6, 49.5], [1000, 1000]], np.float32)
    out = O.scale_and_clip(b, 100, 50, 100, 50)
    assert out.tolist() == [[1, 2], [3, 0], [99, 49], [99, 49]]       # half away from zero, clamp to [0, ori-1]
    assert O.scale_and_clip(np.array([[10, 10]] * 4, np.float32), 100, 100, 250, 50)[0].tolist() == [25, 5]


def test_crop_golden_and_properties():
    page, boxes = G["crop_page"], G["crop_boxes"]
    for k, b in enumerate(boxes):
        assert np.array_equal(O.get_crop_img(page, b), G[f"crop_out_{k}"])
    # axis-aligned box == translation class: interior is an exact copy, the 1-px bicubic border falls back to white
    r, t, cls = O.projection(boxes[0])
    assert r == 0 and cls == 0 and t[2] == 10 and t[5] == 10
    c = O.get_crop_img(page, boxes[0])
    assert c.shape == (30, 140, 3) and np.array_equal(c[1:-3, 1:-3], page[11:37, 11:147])
    assert O.crop_dims(boxes[2]) == (95, 25, 1)                      # h/w >= 1.5 -> rotate270 (dims swap)
    assert (O.get_crop_img(page, boxes[3])[0] == 255).all()          # taps outside the page -> white


def test_resize_norm_image_pad_and_ratio():
    crop = np.random.default_rng(9).integers(0, 256, (30, 100, 3), dtype=np.uint8)
    a = O.resize_norm_image(crop, (3, 48, 192), None)
    assert a.shape == (3, 48, 192) and (a[:, :, 160:] == 0).all() and (a[:, :, :160] != 0).any()
    b = O.resize_norm_image(crop, (3, 48, 320), 7.5)
    assert b.shape == (3, 48, 360)
    f = O.resize_norm_image(crop, (3, 48, 192), None, flip180=True)
    assert np.array_equal(f, O.resize_norm_image(np.ascontiguousarray(crop[::-1, ::-1]), (3, 48, 192), None))


def test_cls_and_ctc_rules():
    st, idx, sc = O.cls_postprocess(np.array([[0.5, 0.5], [0.1, 0.9], [0.7, 0.3]], np.float32))
    assert st == 0 and idx.tolist() == [0, 1, 0]                      # first maximum wins
    assert O.cls_postprocess(np.array([[np.nan, 1.0]], np.float32))[0] == -1
    x = np.zeros((2, 6, 5), np.float32)
    for t, c in enumerate([1, 1, 0, 1, 2, 2]):
        x[0, t, c] = 0.5 + 0.1 * t
    x[1, :, 0] = 1.0                                                  # all blank
    st, idx, prob, tok, cnt, sc = O.ctc_decode(x)
    assert tok[0, :3].tolist() == [1, 1, 2] and cnt.tolist() == [3, 0]
    assert np.float32(sc[0]) == np.float32((np.float32(0.5) + np.float32(0.8) + np.float32(0.9)) / np.float32(3)) and np.isnan(sc[1])


def test_ctc_golden():
    from tools.synth import gen_ctc_logits, synth_dict_text
    st, idx, prob, tok, cnt, sc = O.ctc_decode(G["ctc_small_logits"])
    assert np.array_equal(tok, G["ctc_small_tokens"]) and np.array_equal(cnt, G["ctc_small_counts"])
    assert np.array_equal(sc, G["ctc_small_scores"], equal_nan=True)
    big = gen_ctc_logits(int(G["ctc_big_seed"]), 16, 40, 6625)
    assert zlib.crc32(big.tobytes()) == int(G["ctc_big_crc"])
    st, idx, prob, tok, cnt, sc = O.ctc_decode(big)
    assert np.array_equal(tok, G["ctc_big_tokens"]) and np.array_equal(sc, G["ctc_big_scores"], equal_nan=True)
    chars = O.rec_character(synth_dict_text())
    assert len(chars) == 6625 and chars[0] == "blank" and chars[-1] == " "
    assert O.tokens_to_text(tok[0], cnt[0], chars).encode() == G["ctc_big_text0"].tobytes()
    assert O.rec_character(" a \n　\nb\r\n") == ["blank", "a", "", "b", " "]   # trim, whitespace-only key -> ""
