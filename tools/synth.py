"""Synthetic inputs for tests and bench.  The real `ppocr_keys_v1.txt` is not on
disk, so the dictionary is a synthetic 6623-line file (ASCII 0x21-0x7E, then CJK from U+4E00) which
gives the same 6625 classes after RecCharacter::new adds "blank" and " "."""
from __future__ import annotations

import numpy as np


def synth_dict_text(n_lines: int = 6623) -> str:
    chars = [chr(c) for c in range(0x21, 0x7F)]
    c = 0x4E00
    while len(chars) < n_lines:
        chars.append(chr(c))
        c += 1
    return "\n".join(chars[:n_lines]) + "\n"


def gen_ctc_logits(seed: int, n: int, T: int, C: int, tie_frac=0.01, blank_line_frac=0.005) -> np.ndarray:
    """Config-3 style logits: per step blank p=.45, repeat-previous p=.2, else uniform class; background
    U[0,1e-3), winner U[.5,1); a fraction of steps carries an exact tie between two classes (first-max
    rule), a fraction of lines is all-blank (NaN score)."""
    rng = np.random.default_rng(seed)
    x = (rng.random((n, T, C), dtype=np.float32) * np.float32(1e-3)).astype(np.float32)
    for i in range(n):
        allblank = rng.random() < blank_line_frac
        prev = 0
        for t in range(T):
            u = rng.random()
            if allblank or u < 0.45:
                c = 0
            elif u < 0.65:
                c = prev
            else:
                c = int(rng.integers(1, C))
            prev = c
            v = np.float32(0.5 + 0.5 * rng.random())
            x[i, t, c] = v
            if rng.random() < tie_frac:
                c2 = int(rng.integers(0, C))
                x[i, t, c2] = v
    return x


def gen_probmap(seed: int, h: int = 960, w: int = 960, k_range=(20, 60), wide_angle: bool | None = None,
                border_touch_p: float = 0.05, ring_p: float = 0.0) -> np.ndarray:
    """Config-2 style DB probability map with planted rotated text rectangles (generator only: uses cv2).
    inside ~ clip(N(.85,.05), .55, 1) with a 2-px linear edge ramp, background U[0,.2)."""
    import cv2
    rng = np.random.default_rng(seed)
    if wide_angle is None:
        wide_angle = rng.random() < 0.10
    occ = np.zeros((h, w), np.uint8)
    mask = np.zeros((h, w), np.uint8)
    k = int(rng.integers(k_range[0], k_range[1] + 1))
    placed, tries = 0, 0
    while placed < k and tries < k * 40:
        tries += 1
        rw = float(rng.uniform(40, min(400, w * 0.8)))
        rh = float(rng.uniform(12, 48))
        ang = float(rng.uniform(-90, 90) if wide_angle else rng.uniform(-15, 15))
        if rng.random() < 0.3:
            ang = 0.0
        if rng.random() < border_touch_p:
            cx = float(rng.choice([rw * 0.3, w - rw * 0.3]))
            cy = float(rng.uniform(0, h))
        else:
            cx, cy = float(rng.uniform(0, w)), float(rng.uniform(0, h))
        box = cv2.boxPoints(((cx, cy), (rw, rh), ang)).astype(np.int32)
        one = np.zeros((h, w), np.uint8)
        cv2.fillPoly(one, [box], 255)
        if one.sum() == 0:
            continue
        grown = cv2.dilate(one, np.ones((29, 29), np.uint8))
        if (grown & occ).any():
            continue
        if ring_p > 0 and rng.random() < ring_p and rh > 30:
            inner = cv2.boxPoints(((cx, cy), (rw * 0.6, rh * 0.4), ang)).astype(np.int32)
            cv2.fillPoly(one, [inner], 0)
        occ |= one
        mask |= one
        placed += 1
    dist = cv2.distanceTransform(mask, cv2.DIST_L2, 3)
    ramp = np.clip(dist / 2.0, 0.0, 1.0).astype(np.float32)
    inside = np.clip(rng.normal(0.85, 0.05, (h, w)), 0.55, 1.0).astype(np.float32)
    bg = (rng.random((h, w)) * 0.2).astype(np.float32)
    prob = np.where(mask > 0, bg + (inside - bg) * ramp, bg).astype(np.float32)
    return np.ascontiguousarray(prob)


def gen_page(seed: int, h: int = 1280, w: int = 1280, n_lines=(20, 45), invert_p: float = 0.25):
    """Synthetic rendered-text page (Pillow default font) + the text-line rectangles that were drawn.
    returns (rgb uint8 [h,w,3], rects [(x0,y0,x1,y1,rot180)])"""
    from PIL import Image, ImageDraw, ImageFont
    rng = np.random.default_rng(seed)
    inv = rng.random() < invert_p
    bg, fg = (0, 255) if inv else (255, 0)
    img = Image.new("RGB", (w, h), (bg, bg, bg))
    drw = ImageDraw.Draw(img)
    rects = []
    y = int(rng.integers(10, 40))
    target = int(rng.integers(n_lines[0], n_lines[1] + 1))
    words = ["retto", "B200", "kernel", "ocr", "page", "line", "text", "probability", "contour", "0123456789", "HBM3e", "roofline"]
    while y < h - 60 and len(rects) < target:
        size = int(rng.integers(14, 41))
        try:
            font = ImageFont.load_default(size=size)
        except TypeError:
            font = ImageFont.load_default()
        x = int(rng.integers(10, max(11, w // 4)))
        s = " ".join(rng.choice(words) for _ in range(int(rng.integers(2, 9))))
        box = drw.textbbox((x, y), s, font=font)
        if box[2] >= w - 5:
            s = s[: max(3, int(len(s) * (w - 10 - x) / max(1, box[2] - x)) - 1)]
            box = drw.textbbox((x, y), s, font=font)
        rot = rng.random() < 0.3
        if rot:
            tmp = Image.new("RGB", (box[2] - box[0] + 4, box[3] - box[1] + 4), (bg, bg, bg))
            ImageDraw.Draw(tmp).text((2 - (box[0] - x), 2 - (box[1] - y)), s, font=font, fill=(fg, fg, fg))
            img.paste(tmp.rotate(180), (box[0] - 2, box[1] - 2))
        else:
            drw.text((x, y), s, font=font, fill=(fg, fg, fg))
        rects.append((box[0], box[1], box[2], box[3], bool(rot)))
        y = box[3] + int(rng.integers(14, 40))
    return np.asarray(img, dtype=np.uint8).copy(), rects


def probmap_from_rects(seed: int, rects, h: int, w: int, shrink: float = 0.3) -> np.ndarray:
    """A DBNet-like probability map for a rendered page: shrunk text-line kernels at ~.85, 2-px ramp,
    background U[0,.2) (what the det model would emit for that page, without running a model)."""
    import cv2
    rng = np.random.default_rng(seed)
    mask = np.zeros((h, w), np.uint8)
    for (x0, y0, x1, y1, _r) in rects:
        bh = y1 - y0
        d = int(bh * shrink * 0.5)
        cv2.rectangle(mask, (int(x0) + d, int(y0) + d), (int(x1) - d, int(y1) - d), 255, -1)
    dist = cv2.distanceTransform(mask, cv2.DIST_L2, 3)
    ramp = np.clip(dist / 2.0, 0.0, 1.0).astype(np.float32)
    inside = np.clip(rng.normal(0.85, 0.05, (h, w)), 0.55, 1.0).astype(np.float32)
    bg = (rng.random((h, w)) * 0.2).astype(np.float32)
    return np.ascontiguousarray(np.where(mask > 0, bg + (inside - bg) * ramp, bg).astype(np.float32))
