"""Generates tests/golden/golden_v1.npz — small known-answer fixtures for every stage of the path.

Provenance: the reference (Rust) cannot be built or run in this image and its own tests hold no vectors
(SURVEY.md §8c), so these vectors come from the CPU oracle (oracle/retto_oracle.cpp, libm mode 0 = glibc,
i.e. the reference-faithful mode).  They pin the ORACLE against regressions and give the CUDA path a
fixture that does not depend on the oracle being importable; they are not reference outputs.
Run:  python tools/gen_golden.py
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tools.synth import gen_ctc_logits, gen_probmap, synth_dict_text  # noqa: E402


def main():
    O.set_libm(0)
    rng = np.random.default_rng(20261017)
    g = {}
    # (1) normalise tables for the two formulas (det: x*(1/255) -.5 /.5 ; cls/rec: x/255 -.5 /.5)
    ramp = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    g["norm_det"] = O.det_preprocess(np.ascontiguousarray(np.tile(ramp, (2, 2, 1))), limit_len=32)[0, 0, :16, :16].reshape(256)
    g["norm_rec"] = O.resize_norm_image(np.ascontiguousarray(ramp[:, :, :].reshape(1, 256, 3).repeat(48, 0)), (3, 48, 256), None)[0, 0, :256]
    # (2) thumbnail: identity, 2:1, down to 1984-like ratio, up-scale, crop -> 48 x w
    cases = [((24, 36), (24, 36)), ((40, 64), (20, 32)), ((58, 82), (28, 40)), ((30, 40), (46, 62)), ((21, 90), (48, 206)), ((37, 33), (48, 43)), ((9, 7), (32, 32))]
    for k, ((h, w), (nh, nw)) in enumerate(cases):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        g[f"thumb_in_{k}"] = img
        g[f"thumb_out_{k}"] = O.thumbnail(img, nh, nw)
    g["thumb_n"] = np.int32(len(cases))
    # (3) planted-rect probability maps with per-stage dumps (incl. border-touching, rotated, a ring with a hole, noise)
    maps = []
    for k in range(5):
        maps.append(gen_probmap(700 + k, 96, 128, k_range=(2, 4), wide_angle=(k % 2 == 1), border_touch_p=0.4))
    ring = np.full((96, 128), 0.05, np.float32)
    ring[20:70, 20:110] = 0.9
    ring[35:55, 40:90] = 0.05          # hole -> find_contours returns an extra (hole) border
    maps.append(ring)
    diag = np.full((96, 128), 0.05, np.float32)
    diag[10:30, 10:60] = 0.8
    diag[31:50, 61:120] = 0.8          # touches the first blob diagonally after the 2x2 dilation (8-connectivity)
    maps.append(diag)
    for k, p in enumerate(maps):
        r = O.det_postprocess(p, 96, 128, want_bitmap=True)
        rect1, ss, sc, st = O.det_trace(p, 96, 128)
        g[f"det_prob_{k}"] = p
        g[f"det_bitmap_crc_{k}"] = np.uint32(zlib.crc32(r.bitmap.tobytes()))
        g[f"det_boxes_{k}"] = r.boxes
        g[f"det_scores_{k}"] = r.scores
        g[f"det_rect1_{k}"] = rect1
        g[f"det_status_{k}"] = st
        g[f"det_ncontours_{k}"] = np.int32(len(st))
    g["det_n"] = np.int32(len(maps))
    # (4) rotate-crops
    page = rng.integers(0, 256, (120, 200, 3), dtype=np.uint8)
    boxes = np.array([[[10, 10], [150, 10], [150, 40], [10, 40]], [[20, 50], [180, 62], [178, 90], [18, 78]],
                      [[160, 5], [185, 5], [185, 100], [160, 100]], [[0, 0], [60, 0], [60, 20], [0, 20]]], np.float32)
    g["crop_page"] = page
    g["crop_boxes"] = boxes
    for k, b in enumerate(boxes):
        g[f"crop_out_{k}"] = O.get_crop_img(page, b)
    # (5) CTC: a small-alphabet case stored in full, and a full-size case keyed by seed (tokens only)
    small = gen_ctc_logits(5, 24, 17, 97, tie_frac=0.1, blank_line_frac=0.1)
    st, idx, prob, tok, cnt, sc = O.ctc_decode(small)
    g["ctc_small_logits"], g["ctc_small_tokens"], g["ctc_small_counts"], g["ctc_small_scores"] = small, tok, cnt, sc
    big = gen_ctc_logits(3, 16, 40, 6625)
    st, idx, prob, tok, cnt, sc = O.ctc_decode(big)
    g["ctc_big_seed"] = np.int32(3)
    g["ctc_big_tokens"], g["ctc_big_counts"], g["ctc_big_scores"] = tok, cnt, sc
    g["ctc_big_crc"] = np.uint32(zlib.crc32(big.tobytes()))
    chars = O.rec_character(synth_dict_text())
    g["ctc_big_text0"] = np.frombuffer(O.tokens_to_text(tok[0], cnt[0], chars).encode("utf-8"), np.uint8)
    out = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
