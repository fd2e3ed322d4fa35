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
