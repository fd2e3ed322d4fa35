"""CPU oracle of RettoSession::process_pipeline (session.rs:75-106) built from the oracle stage functions.
TEST INFRASTRUCTURE ONLY (see oracle.py).  worker = object with det/cls/rec numpy callables
(det: [1,3,H,W]->[1,1,H,W]; cls: [n,3,48,192]->[n,2]; rec: [n,3,48,W]->[n,T,C])."""
from __future__ import annotations

import numpy as np

from . import oracle as O


def stable_order_desc_ratio(dims):
    """indices sorted by Reverse(OrderedFloat(h/w)) with a stable sort (cls_processor.rs:137-138)"""
    ratios = [float(h) / float(w) for (h, w) in dims]
    return sorted(range(len(dims)), key=lambda i: -ratios[i])  # Python's sort is stable


def run_page(img: np.ndarray, worker, dict_text: str, max_side_len=2000, min_side_len=30, det_cfg=None, cls_shape=(3, 48, 192),
             rec_shape=(3, 48, 320), batch_num=6, cls_thresh=0.9, cls_label=(0, 180), taps=None, limit_type=0, limit_len=736,
             cls_batch_num=None, rec_batch_num=None):
    """limit_type 0 = LimitType::Min, 1 = Max (det_processor.rs:75-93); cls_/rec_batch_num default to batch_num"""
    cls_batch_num = cls_batch_num or batch_num
    rec_batch_num = rec_batch_num or batch_num
    chars = O.rec_character(dict_text)
    ori_h, ori_w = img.shape[:2]
    page = O.resize_both(img, max_side_len, min_side_len)                       # session.rs:81-82
    after_h, after_w = page.shape[:2]
    det_in = O.det_preprocess(page, limit_type, limit_len)                      # det_processor.rs:256-274
    pred = np.ascontiguousarray(worker.det(det_in), dtype=np.float32)           # session.rs:86
    det = O.det_postprocess(pred[0, 0], after_h, after_w, det_cfg)              # det_processor.rs:279-335
    assert det.status >= 0, "reference would panic in det postprocess"
    crops = [O.get_crop_img(page, b) for b in det.boxes]                        # session.rs:88-92
    boxes = [O.scale_and_clip(b, after_w, after_h, ori_w, ori_h) for b in det.boxes]   # session.rs:94-97
    dims = [c.shape[:2] for c in crops]
    flipped = [False] * len(crops)
    # cls (cls_processor.rs:127-172)
    order = stable_order_desc_ratio(dims)
    cls_res = [None] * len(crops)
    cls_batches = []
    for b0 in range(0, len(order), cls_batch_num):
        idxs = order[b0:b0 + cls_batch_num]
        batch = np.stack([O.resize_norm_image(crops[i], cls_shape, None) for i in idxs])
        cls_batches.append(batch)
        logits = np.ascontiguousarray(worker.cls(batch), dtype=np.float32)
        st, am, sc = O.cls_postprocess(logits)
        assert st == 0
        for k, i in enumerate(idxs):
            label = cls_label[am[k]]
            if label == 180 and sc[k] >= np.float32(cls_thresh):
                flipped[i] = not flipped[i]                                     # rotate_180_in_place
            cls_res[i] = (int(label), float(sc[k]))
    # rec (rec_processor.rs:214-270)
    rec_res = [None] * len(crops)
    max_wh_ratio = np.float32(rec_shape[2]) / np.float32(rec_shape[1])
    rec_batches = []
    for b0 in range(0, len(order), rec_batch_num):
        idxs = order[b0:b0 + rec_batch_num]
        for i in idxs:
            wh = np.float32(dims[i][1]) / np.float32(dims[i][0])
            if wh > max_wh_ratio:
                max_wh_ratio = wh
        batch = np.stack([O.resize_norm_image(crops[i], rec_shape, float(max_wh_ratio), flip180=flipped[i]) for i in idxs])
        rec_batches.append(batch)
        logits = np.ascontiguousarray(worker.rec(batch), dtype=np.float32)
        st, idx, prob, tok, cnt, sc = O.ctc_decode(logits)
        assert st == 0
        for k, i in enumerate(idxs):
            rec_res[i] = (O.tokens_to_text(tok[k], cnt[k], chars), float(sc[k]))
    if taps is not None:
        taps.update(page=page, det_in=det_in, crops=crops, flipped=flipped, cls_batches=cls_batches, rec_batches=rec_batches, det=det)
    return dict(boxes=np.array(boxes, np.float32).reshape(-1, 4, 2), scores=det.scores, cls=cls_res, rec=rec_res)
