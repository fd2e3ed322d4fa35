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
