"""ctypes wrapper around oracle/libretto_oracle.so — the CPU oracle.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under retto_b200/ imports this module.
Parity status of the oracle itself: see the header of retto_oracle.cpp ("PARITY UNPINNED" for the
third-party crate semantics; cross-checked against cv2/numpy where an independent implementation
exists on this box).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libretto_oracle.so")
_SHIM_SO = os.path.join(_HERE, "diag", "libcrmath_shim.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "retto_oracle.cpp")
    shim_src = os.path.join(_HERE, "diag", "crmath_shim.cpp")
    hdr = os.path.join(_HERE, "..", "retto_b200", "csrc", "rt_fmath.h")

    def _stale(so, deps):
        return (not os.path.exists(so)) or any(os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(so) for p in deps)

    stale = _stale(_SO, (src, os.path.join(_HERE, "jpeg_oracle.cpp"))) or _stale(_SHIM_SO, (shim_src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_det_postprocess.restype = C.c_int
        _lib.orc_find_contours.restype = C.c_int
        _lib.orc_set_trig_hooks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_trig_hooked.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class DetCfg(C.Structure):
    _fields_ = [
        ("thresh", C.c_float),
        ("box_thresh", C.c_float),
        ("unclip_ratio", C.c_float),
        ("min_mini_box_size", C.c_int),
        ("dilate", C.c_int),
        ("quirk_x0", C.c_int),
    ]


def default_det_cfg(**kw) -> DetCfg:
    # det_processor.rs:75-93 defaults
    d = dict(thresh=0.3, box_thresh=0.5, unclip_ratio=1.6, min_mini_box_size=3, dilate=1, quirk_x0=1)
    d.update(kw)
    return DetCfg(**d)


_shim = None


def set_libm(mode: int):
    """0 = host glibc: what the reference's f64 trig calls, the ASSERTED mode of every parity test.
    1 = DIAGNOSTIC: install oracle/diag/libcrmath_shim.so (the CUDA path's correctly-rounded trig) through the
    oracle's hooks, for second runs that count glibc-vs-CUDA differences; never the asserted mode."""
    global _shim
    L = lib()
    if mode == 0:
        L.orc_set_trig_hooks(None, None, None)
        return
    if _shim is None:
        _shim = C.CDLL(_SHIM_SO)
    L.orc_set_trig_hooks(C.cast(_shim.diag_cr_atan2, C.c_void_p), C.cast(_shim.diag_cr_sincos, C.c_void_p),
                         C.cast(_shim.diag_cr_acos, C.c_void_p))


def libm_mode() -> int:
    return int(lib().orc_trig_hooked())


def thumbnail(img: np.ndarray, nh: int, nw: int) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, c = img.shape
    out = np.zeros((nh, nw, c), np.uint8)
    lib().orc_thumbnail(_p(img, C.c_uint8), h, w, c, _p(out, C.c_uint8), nh, nw)
    return out


def resize_both_plan(h: int, w: int, max_len: int = 2000, min_len: int = 30):
    dims = (C.c_int * 4)()
    n = lib().orc_resize_both_plan(h, w, max_len, min_len, dims)
    return [(dims[2 * i], dims[2 * i + 1]) for i in range(n)]


def resize_both(img: np.ndarray, max_len: int = 2000, min_len: int = 30) -> np.ndarray:
    h, w, _ = img.shape
    for nh, nw in resize_both_plan(h, w, max_len, min_len):
        img = thumbnail(img, nh, nw)
    return img


def resize_either_plan(h: int, w: int, limit_type: int = 0, limit_len: int = 736):
    oh, ow = C.c_int(), C.c_int()
    lib().orc_resize_either_plan(h, w, limit_type, limit_len, C.byref(oh), C.byref(ow))
    return oh.value, ow.value


def det_preprocess(img: np.ndarray, limit_type: int = 0, limit_len: int = 736, scale=None, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, _ = img.shape
    oh, ow = resize_either_plan(h, w, limit_type, limit_len)
    if scale is None:
        scale = np.float32(1.0) / np.float32(255.0)
    out = np.zeros((1, 3, oh, ow), np.float32)
    m = np.asarray(mean, np.float32)
    s = np.asarray(std, np.float32)
    lib().orc_det_preprocess(_p(img, C.c_uint8), h, w, C.c_float(float(scale)), _p(m, C.c_float), _p(s, C.c_float), _p(out, C.c_float), oh, ow)
    return out


def threshold_dilate(pred: np.ndarray, thr: float = 0.3, dilate: bool = True) -> np.ndarray:
    pred = np.ascontiguousarray(pred, dtype=np.float32)
    h, w = pred.shape
    out = np.zeros((h, w), np.uint8)
    lib().orc_threshold_dilate(_p(pred, C.c_float), h, w, C.c_float(thr), int(dilate), _p(out, C.c_uint8))
    return out


def find_contours(mask: np.ndarray, quirk_x0: int = 1):
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w = mask.shape
    max_pts = 8 * h * w + 16
    max_c = h * w + 1
    pts = np.zeros((max_pts, 2), np.int32)
    offs = np.zeros(max_c + 1, np.int32)
    hole = np.zeros(max_c, np.int32)
    n = lib().orc_find_contours(_p(mask, C.c_uint8), h, w, quirk_x0, _p(pts, C.c_int), C.c_longlong(max_pts), _p(offs, C.c_int), _p(hole, C.c_int), max_c)
    assert n >= 0
    return [(pts[offs[i]:offs[i + 1]].copy(), int(hole[i])) for i in range(n)]


def convex_hull(pts: np.ndarray) -> np.ndarray:
    pts = np.ascontiguousarray(pts, dtype=np.int32)
    out = np.zeros((len(pts) + 1, 2), np.int32)
    n = lib().orc_convex_hull(_p(pts, C.c_int), len(pts), _p(out, C.c_int))
    return out[:n]


def min_area_rect(pts: np.ndarray) -> np.ndarray:
    pts = np.ascontiguousarray(pts, dtype=np.int32)
    out = np.zeros(8, np.float64)
    lib().orc_min_area_rect(_p(pts, C.c_int), len(pts), _p(out, C.c_double))
    return out.reshape(4, 2)


def box_score_fast(pred: np.ndarray, quad: np.ndarray):
    pred = np.ascontiguousarray(pred, dtype=np.float32)
    q = np.ascontiguousarray(quad, dtype=np.int32).reshape(8)
    sc = C.c_float()
    r = lib().orc_box_score_fast(_p(pred, C.c_float), pred.shape[0], pred.shape[1], _p(q, C.c_int), C.byref(sc))
    return r, sc.value


def polygon_mask(cw: int, ch: int, quad: np.ndarray):
    q = np.ascontiguousarray(quad, dtype=np.int32).reshape(8)
    out = np.zeros((ch, cw), np.uint8)
    r = lib().orc_polygon_mask(_p(out, C.c_uint8), cw, ch, _p(q, C.c_int))
    return r, out


def unclip(quad: np.ndarray, ratio: float = 1.6):
    q = np.ascontiguousarray(quad, dtype=np.int32).reshape(8)
    out = np.zeros((4096, 2), np.int32)
    d = C.c_float()
    n = lib().orc_unclip(_p(q, C.c_int), C.c_float(ratio), _p(out, C.c_int), 4096, C.byref(d))
    return out[:max(n, 0)].copy(), d.value


@dataclass
class DetResult:
    boxes: np.ndarray  # [n,4,2] f32 (tl,tr,br,bl)
    scores: np.ndarray  # [n] f32
    status: int  # n or -1 (reference would panic)
    bitmap: np.ndarray | None
    comparator_inconsistent: bool


def det_postprocess(pred: np.ndarray, ori_h: int, ori_w: int, cfg: DetCfg | None = None, want_bitmap=False, max_boxes=65536) -> DetResult:
    pred = np.ascontiguousarray(pred, dtype=np.float32)
    h, w = pred.shape
    cfg = cfg or default_det_cfg()
    boxes = np.zeros((max_boxes, 8), np.float32)
    scores = np.zeros(max_boxes, np.float32)
    bm = np.zeros((h, w), np.uint8) if want_bitmap else None
    inc = C.c_int(0)
    n = lib().orc_det_postprocess(
        _p(pred, C.c_float), h, w, ori_h, ori_w, C.byref(cfg), _p(boxes, C.c_float), _p(scores, C.c_float), max_boxes,
        _p(bm, C.c_uint8) if want_bitmap else None, C.byref(inc))
    assert n != -2, "oracle box buffer overflow"
    k = max(n, 0)
    return DetResult(boxes[:k].reshape(k, 4, 2).copy(), scores[:k].copy(), n, bm, bool(inc.value))


def det_trace(pred: np.ndarray, ori_h: int, ori_w: int, cfg: DetCfg | None = None, max_contours=1 << 20):
    pred = np.ascontiguousarray(pred, dtype=np.float32)
    h, w = pred.shape
    cfg = cfg or default_det_cfg()
    rect1 = np.zeros((max_contours, 8), np.int32)
    ss = np.zeros(max_contours, np.float32)
    sc = np.zeros(max_contours, np.float32)
    st = np.zeros(max_contours, np.int32)
    n = lib().orc_det_trace(_p(pred, C.c_float), h, w, ori_h, ori_w, C.byref(cfg), _p(rect1, C.c_int), _p(ss, C.c_float), _p(sc, C.c_float), _p(st, C.c_int), max_contours)
    assert n >= 0
    return rect1[:n].copy(), ss[:n].copy(), sc[:n].copy(), st[:n].copy()


def scale_and_clip(box: np.ndarray, bw, bh, ow, oh) -> np.ndarray:
    b = np.ascontiguousarray(box, dtype=np.float32).reshape(8).copy()
    lib().orc_scale_and_clip(_p(b, C.c_float), C.c_double(bw), C.c_double(bh), C.c_double(ow), C.c_double(oh))
    return b.reshape(4, 2)


def crop_dims(box: np.ndarray):
    b = np.ascontiguousarray(box, dtype=np.float32).reshape(8)
    cw, ch, rot = C.c_int(), C.c_int(), C.c_int()
    lib().orc_crop_dims(_p(b, C.c_float), C.byref(cw), C.byref(ch), C.byref(rot))
    return cw.value, ch.value, rot.value


def get_crop_img(img: np.ndarray, box: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    b = np.ascontiguousarray(box, dtype=np.float32).reshape(8)
    cw, ch, _ = crop_dims(b)
    out = np.zeros((ch, cw, 3), np.uint8)
    r = lib().orc_get_crop_img(_p(img, C.c_uint8), img.shape[0], img.shape[1], _p(b, C.c_float), _p(out, C.c_uint8))
    assert r == 0, "reference would panic: degenerate projection"
    return out


def projection(box: np.ndarray):
    b = np.ascontiguousarray(box, dtype=np.float32).reshape(8)
    t = np.zeros(9, np.float32)
    cls = C.c_int()
    r = lib().orc_projection(_p(b, C.c_float), _p(t, C.c_float), C.byref(cls))
    return r, t, cls.value


def resize_norm_plan(ori_h, ori_w, img_h, img_w_cfg, ratio=None):
    iw, rw = C.c_int(), C.c_int()
    lib().orc_resize_norm_plan(ori_h, ori_w, img_h, img_w_cfg, int(ratio is not None), C.c_float(float(ratio or 0.0)), C.byref(iw), C.byref(rw))
    return iw.value, rw.value


def resize_norm_image(crop: np.ndarray, shape=(3, 48, 192), ratio=None, flip180=False, ori_hw=None) -> np.ndarray:
    crop = np.ascontiguousarray(crop, dtype=np.uint8)
    ch, cw, _ = crop.shape
    oh, ow = ori_hw or (ch, cw)
    img_w, resized_w = resize_norm_plan(oh, ow, shape[1], shape[2], ratio)
    out = np.zeros((3, shape[1], img_w), np.float32)
    lib().orc_resize_norm_image(_p(crop, C.c_uint8), ch, cw, int(flip180), shape[1], img_w, resized_w, _p(out, C.c_float))
    return out


def cls_postprocess(logits: np.ndarray):
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    n, k = logits.shape
    idx = np.zeros(n, np.int32)
    sc = np.zeros(n, np.float32)
    r = lib().orc_cls_postprocess(_p(logits, C.c_float), n, k, _p(idx, C.c_int), _p(sc, C.c_float))
    return r, idx, sc


def ctc_decode(logits: np.ndarray):
    """returns (status, idx[n,T], prob[n,T], tokens[n,T] (-1 padded), counts[n], score[n])"""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    n, T, Cc = logits.shape
    idx = np.zeros((n, T), np.int32)
    prob = np.zeros((n, T), np.float32)
    tok = np.zeros((n, T), np.int32)
    cnt = np.zeros(n, np.int32)
    sc = np.zeros(n, np.float32)
    r = lib().orc_ctc_decode(_p(logits, C.c_float), n, T, Cc, _p(idx, C.c_int), _p(prob, C.c_float), _p(tok, C.c_int), _p(cnt, C.c_int), _p(sc, C.c_float))
    return r, idx, prob, tok, cnt, sc


# ---- host-side restatements that need no C (rec_processor.rs:29-46, 48-97) ---------------
def rec_character(dict_text: str):
    """RecCharacter::new: lines trimmed, 'blank' prepended, ' ' appended (rec_processor.rs:29-46)."""
    d = [ln.strip() for ln in dict_text.splitlines()]  # str::lines + str::trim
    return ["blank"] + d + [" "]


def tokens_to_text(tokens_row, count, chars):
    return "".join(chars[t] for t in tokens_row[:count])


def jpeg_decode(data: bytes) -> np.ndarray:
    """libjpeg-turbo's default decode of a baseline JPEG (jpeg_oracle.cpp) -> HWC u8 RGB.  Raises ValueError(status) for files the
    restatement does not cover (1 = not a JPEG / truncated, 2 = unsupported kind: progressive, CMYK, ...)."""
    buf = np.frombuffer(data, np.uint8)
    h, w = C.c_int(), C.c_int()
    L = lib()
    L.orc_jpeg_decode.restype = C.c_int
    L.orc_jpeg_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    st = L.orc_jpeg_decode(buf.ctypes.data, len(buf), None, 0, C.byref(h), C.byref(w))
    if st:
        raise ValueError(st)
    out = np.zeros((h.value, w.value, 3), np.uint8)
    st = L.orc_jpeg_decode(buf.ctypes.data, len(buf), out.ctypes.data, out.size, C.byref(h), C.byref(w))
    if st:
        raise ValueError(st)
    return out
