"""BASELINE.json full-size configurations, checked through size-independent properties plus oracle spot checks.
  config 2: DB postprocess on 1024 synthetic 960x960 probability maps in ONE batched call
  config 3: CTC decode of 16384 lines of 40 x 6625 logits (17.4 GB) in ONE call
  config 4: 256 pages 1280x1280 end to end from host memory (batch independence + oracle spot checks)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gen960(seed):
    from tools.synth import gen_probmap
    return gen_probmap(seed, 960, 960)


def test_config2_dbpost_1024_maps_960(ctx, libm_diag):
    """asserted against the glibc oracle on NU unique maps (~10 k boxes); the diagnostic run counts how many of them
    change when the oracle uses the CUDA path's correctly-rounded trig instead"""
    import multiprocessing as mp
    import torch
    from oracle import oracle as O
    NU = 256
    with mp.get_context("fork").Pool(min(16, mp.cpu_count())) as pool:
        cand = pool.map(_gen960, range(2000, 2000 + NU + 192))
    uniq = [p for p in cand if not O.det_postprocess(p, 960, 960).comparator_inconsistent][:NU]
    assert len(uniq) == NU
    refs = [O.det_postprocess(p, 960, 960) for p in uniq]
    libm_diag("config2_960", lambda: [O.det_postprocess(p, 960, 960).boxes for p in uniq], [r.boxes for r in refs])
    dev = [torch.from_numpy(p).cuda() for p in uniq]
    maps = [dev[i % NU].clone() for i in range(1024)]          # 1024 distinct buffers (3.8 GB)
    torch.cuda.synchronize()
    out = ctx.det_postprocess(maps, [(960, 960)] * 1024, max_boxes_total=1024 * 80)
    assert (out.page_status == 0).all()
    total = 0
    for i in range(1024):
        boxes, scores = out.page(i)
        r = refs[i % NU]
        assert np.array_equal(boxes, r.boxes), f"map {i}"
        assert np.array_equal(scores.view(np.uint32), r.scores.view(np.uint32)), f"map {i}"
        total += len(boxes)
    assert total == (1024 // NU) * sum(len(r.boxes) for r in refs) and total > 1024 * 15
    # bitmap / labels of a few pages, incl. the last one of the batch
    for i in (0, 517, 1023):
        bm = ctx.fetch_bitmap(i, 960, 960)
        assert np.array_equal(bm, O.threshold_dilate(uniq[i % NU]))
        lab = ctx.fetch_labels(i, 960, 960)
        fg = lab >= 0
        assert np.array_equal(fg, bm > 0)
        ys, xs = np.nonzero(fg)
        assert (lab[fg] <= ys * 960 + xs).all()                # label = min raster index of the component


def test_config3_ctc_16k_lines(ctx, synth_dict):
    import torch
    from oracle import oracle as O
    ctx.dict_load(synth_dict)
    N, T, C = 16384, 40, 6625
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    x = torch.rand((N, T, C), device="cuda", generator=g) * 1e-3           # 17.4 GB
    win = torch.randint(1, C, (N, T), device="cuda", generator=g)
    win[torch.rand((N, T), device="cuda", generator=g) < 0.45] = 0
    rep = torch.rand((N, T), device="cuda", generator=g) < 0.2
    win[:, 1:][rep[:, 1:]] = win[:, :-1][rep[:, 1:]]
    val = 0.5 + 0.5 * torch.rand((N, T), device="cuda", generator=g)
    x.scatter_(2, win.unsqueeze(-1), val.unsqueeze(-1))
    tie = torch.rand((N, T), device="cuda", generator=g) < 0.01            # exact ties: first maximum must win
    other = torch.randint(0, C, (N, T), device="cuda", generator=g)
    x.scatter_(2, other.unsqueeze(-1), torch.where(tie, val, x.gather(2, other.unsqueeze(-1)).squeeze(-1)).unsqueeze(-1))
    x[::200, :, :] = 0
    x[::200, :, 0] = 1.0                                                    # all-blank lines -> NaN score
    torch.cuda.synchronize()
    texts, scores, tokens, counts = ctx.ctc_decode([x], want_tokens=True)
    idx, prob = ctx.ctc_argmax(x)
    torch.cuda.synchronize()
    ctx.sync()
    # properties at full size: prob == row max; idx is the FIRST index attaining it
    mx = x.amax(dim=2)
    assert torch.equal(prob, mx)
    first = (x == mx.unsqueeze(-1)).int().argmax(dim=2).int()
    assert torch.equal(idx, first)
    ih = idx.cpu().numpy()
    keep = (ih != 0) & np.concatenate([np.ones((N, 1), bool), ih[:, 1:] != ih[:, :-1]], 1)
    assert np.array_equal(counts, keep.sum(1))
    assert np.isnan(scores[::200]).all() and (counts[::200] == 0).all()
    assert sum(len(t) for t in texts) == int(counts.sum())                  # one code point per kept class in the synthetic dictionary
    # oracle spot check on 256 lines spread over the tensor
    sel = np.linspace(0, N - 1, 256).astype(int)
    sub = x[torch.from_numpy(sel).cuda()].cpu().numpy()
    st, oi, op, ot, oc, osc = O.ctc_decode(sub)
    chars = O.rec_character(synth_dict)
    assert np.array_equal(tokens[sel][:, :T], ot) and np.array_equal(scores[sel], osc, equal_nan=True)
    assert [texts[i] for i in sel] == [O.tokens_to_text(ot[k], oc[k], chars) for k in range(256)]


def test_config4_256_pages_1280_end_to_end(ctx, synth_dict, libm_diag):
    """config 4: 256 rendered 1280x1280 pages (32 unique x 8) from HOST memory through retto_b200_run_pages — the chunked
    upload pipeline, every stage batched over the unit.  Properties: replicas of a page give identical results wherever
    they sit in the batch (batch independence), every detected line has a box / label / string, and all 32 unique pages
    are checked against the CPU oracle pipeline run with the same stand-in forwards."""
    from oracle import oracle as O
    from oracle.pipeline import run_page
    from retto_b200.session import CallableWorker, RettoSession
    from tools.demo_worker import StatelessWorker
    from tools.synth import gen_page
    uniq = [gen_page(4 + i, 1280, 1280)[0] for i in range(32)]
    pages = [uniq[i % 32] for i in range(256)]
    w = StatelessWorker()
    ctx.dict_load(synth_dict)
    sess = RettoSession(worker=CallableWorker(w.det, w.cls, w.rec), ctx=ctx)
    got = sess.run_pages(pages)
    assert len(got) == 256 and all(r.status == 0 for r in got)
    n_lines = sum(len(r.det_result) for r in got)
    assert n_lines > 256 * 10
    for i in range(32, 256):
        a, b = got[i % 32], got[i]
        assert len(a.det_result) == len(b.det_result) == len(b.cls_result) == len(b.rec_result)
        for x, y in zip(a.det_result, b.det_result):
            assert np.array_equal(x.boxes, y.boxes) and x.score == y.score
        assert [c.label for c in a.cls_result] == [c.label for c in b.cls_result]
        assert [r.text for r in a.rec_result] == [r.text for r in b.rec_result]
    # every unique page against the CPU oracle pipeline (glibc trig = asserted mode) ...
    refs = [run_page(u, w, synth_dict) for u in uniq]
    for i, ref in enumerate(refs):
        assert len(ref["boxes"]) == len(got[i].det_result) > 0
        for k in range(len(ref["boxes"])):
            assert np.array_equal(got[i].det_result[k].boxes, ref["boxes"][k])
            assert got[i].cls_result[k].label == ref["cls"][k][0]
            assert got[i].rec_result[k].text == ref["rec"][k][0]
    # ... and the diagnostic count of boxes that change with the CUDA path's trig hooked into the oracle
    libm_diag("config4_1280_pages", lambda: [np.asarray(run_page(u, w, synth_dict)["boxes"]) for u in uniq],
              [np.asarray(r["boxes"]) for r in refs])
