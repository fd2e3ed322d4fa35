"""GPU vs the committed golden fixtures (tests/golden/golden_v1.npz, made by tools/gen_golden.py from the
oracle in glibc mode) — does not need the oracle at run time."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_thumbnail_golden(ctx):
    import torch
    srcs = [_t(G[f"thumb_in_{k}"]) for k in range(int(G["thumb_n"]))]
    dims = [G[f"thumb_out_{k}"].shape[:2] for k in range(int(G["thumb_n"]))]
    torch.cuda.synchronize()
    outs = ctx.thumbnail(srcs, dims)
    ctx.sync()
    for k, o in enumerate(outs):
        assert np.array_equal(o.cpu().numpy(), G[f"thumb_out_{k}"])


def test_det_postprocess_golden(ctx):
    import torch
    n = int(G["det_n"])
    probs = [_t(G[f"det_prob_{k}"]) for k in range(n)]
    torch.cuda.synchronize()
    ctx.enable_trace(True)
    out = ctx.det_postprocess(probs, [(96, 128)] * n)
    for k in range(n):
        tr = ctx.fetch_trace(k)
        boxes, scores = out.page(k)
        assert np.array_equal(boxes, G[f"det_boxes_{k}"])
        assert np.array_equal(scores.view(np.uint32), G[f"det_scores_{k}"].view(np.uint32))
        keep = tr["status"] != 6
        order = np.argsort(tr["key"][keep], kind="stable")
        assert np.array_equal(tr["rect1"][keep][order], G[f"det_rect1_{k}"])
        assert np.array_equal(tr["status"][keep][order], G[f"det_status_{k}"])
    ctx.enable_trace(False)
    ring = n - 2
    assert ctx.fetch_trace(ring)["n_holes"] == 1 and int(G[f"det_ncontours_{ring}"]) == 2


def test_crop_golden(ctx):
    import torch
    page, boxes = G["crop_page"], G["crop_boxes"]
    g = _t(page)
    torch.cuda.synchronize()
    infos = ctx.crop_boxes([g], [0] * len(boxes), boxes)
    for k in range(len(boxes)):
        assert np.array_equal(ctx.crop_fetch(k, infos[k]), G[f"crop_out_{k}"])


def test_ctc_golden(ctx, synth_dict):
    import torch
    from tools.synth import gen_ctc_logits
    ctx.dict_load(synth_dict)
    big = gen_ctc_logits(int(G["ctc_big_seed"]), 16, 40, 6625)
    g = _t(big)
    torch.cuda.synchronize()
    texts, scores, tokens, counts = ctx.ctc_decode([g], want_tokens=True)
    assert np.array_equal(tokens, G["ctc_big_tokens"]) and np.array_equal(counts, G["ctc_big_counts"])
    assert np.array_equal(scores, G["ctc_big_scores"], equal_nan=True)
    assert texts[0].encode() == G["ctc_big_text0"].tobytes()
