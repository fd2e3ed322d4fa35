"""GPU parity: CTC decode (K10), det preprocess (K1), thumbnail — CUDA path vs the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_ctc_decode_parity(ctx, synth_dict):
    import torch
    from oracle import oracle as O
    from tools.synth import gen_ctc_logits
    ctx.dict_load(synth_dict)
    assert ctx.dict_size == 6625
    chars = O.rec_character(synth_dict)
    assert len(chars) == 6625
    for (n, T, C) in [(64, 40, 6625), (7, 13, 6625), (3, 97, 6625)]:
        x = gen_ctc_logits(3, n, T, C)
        st, idx, prob, tok, cnt, sc = O.ctc_decode(x)
        assert st == 0
        xg = _t(x)
        torch.cuda.synchronize()
        texts, scores, tokens, counts = ctx.ctc_decode([xg], want_tokens=True)
        assert (counts == cnt).all()
        assert (tokens[:, :T] == tok).all()
        ref_texts = [O.tokens_to_text(tok[i], cnt[i], chars) for i in range(n)]
        assert texts == ref_texts
        # scores bit-exact (sequential f32 mean), NaN == NaN for empty decodes
        assert np.array_equal(scores.view(np.uint32), sc.view(np.uint32)) or np.allclose(scores, sc, equal_nan=True, rtol=0, atol=0)
        gi, gp = ctx.ctc_argmax(xg)
        torch.cuda.synchronize()
        ctx.sync()
        assert (gi.cpu().numpy() == idx).all()
        assert np.array_equal(gp.cpu().numpy(), prob)


def test_ctc_misaligned_and_multi(ctx, synth_dict):
    """several tensors with different T in one call; rows start at every 4-byte phase."""
    import torch
    from oracle import oracle as O
    from tools.synth import gen_ctc_logits
    ctx.dict_load(synth_dict)
    chars = O.rec_character(synth_dict)
    xs = [gen_ctc_logits(10 + k, n, T, 6625) for k, (n, T) in enumerate([(6, 40), (5, 41), (1, 96), (6, 55)])]
    gs = [_t(x) for x in xs]
    torch.cuda.synchronize()
    texts, scores = ctx.ctc_decode(gs)
    ref_t, ref_s = [], []
    for x in xs:
        st, idx, prob, tok, cnt, sc = O.ctc_decode(x)
        ref_t += [O.tokens_to_text(tok[i], cnt[i], chars) for i in range(x.shape[0])]
        ref_s.append(sc)
    assert texts == ref_t
    assert np.array_equal(np.concatenate(ref_s), scores, equal_nan=True)


def test_ctc_nan_is_error(ctx, synth_dict):
    import torch
    from retto_b200 import RettoB200Error
    ctx.dict_load(synth_dict)
    x = np.zeros((2, 5, 6625), np.float32)
    x[1, 3, 77] = np.nan
    g = _t(x)
    torch.cuda.synchronize()
    with pytest.raises(RettoB200Error) as e:
        ctx.ctc_decode([g])
    assert e.value.status == 5  # ERR_NAN_LOGITS (reference: argmax().unwrap() panics, rec_processor.rs:198)


# identity, both axes up, and every mix of an axis rounded up / down to the multiple of 32 (windows of 1-2 pixels): the
# column-per-thread kernel of det_preprocess.cu, and the generic thumbnail_pixel kernel forced through the environment switch
@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("hw", [(960, 960), (736, 1280), (640, 480), (480, 640), (100, 333), (1000, 740), (1010, 745), (745, 1010), (750, 1500),
                                (1111, 1999), (737, 737), (1487, 751), (300, 1400)])
def test_det_preprocess_parity(ctx, hw, generic, monkeypatch):
    import torch
    from oracle import oracle as O
    if generic:
        monkeypatch.setenv("RETTO_B200_DETPRE_GENERIC", "1")
    h, w = hw
    rng = np.random.default_rng(h * 7 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = O.det_preprocess(img)
    g = _t(img)
    torch.cuda.synchronize()
    out = ctx.det_preprocess([g])[0]
    ctx.sync()
    got = out.cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))  # bit-exact f32


def test_det_preprocess_batch_mixed(ctx):
    import torch
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    imgs = [rng.integers(0, 256, s + (3,), dtype=np.uint8) for s in [(768, 800), (500, 900), (736, 736), (64, 2000)]]
    gs = [_t(i) for i in imgs]
    torch.cuda.synchronize()
    outs = ctx.det_preprocess(gs)
    ctx.sync()
    for i, o in zip(imgs, outs):
        assert np.array_equal(o.cpu().numpy().view(np.uint32), O.det_preprocess(i).view(np.uint32))


@pytest.mark.parametrize("generic", [False, True])   # thumbnail_cols_kernel (ratios in [1, 3]) / the generic kernel for every case
@pytest.mark.parametrize("case", [((200, 300), (200, 300)), ((400, 600), (200, 300)), ((2896, 4096), (1408, 1984)),
                                  ((480, 640), (736, 992)), ((37, 211), (48, 274)), ((61, 150), (48, 118)), ((20, 20), (32, 32)),
                                  ((3000, 2000), (1984, 1312)), ((900, 2500), (704, 1984)), ((2001, 2001), (1984, 1984)), ((96, 96), (32, 32)),
                                  ((100, 97), (32, 32))])
def test_thumbnail_parity(ctx, case, generic, monkeypatch):
    import torch
    from oracle import oracle as O
    if generic:
        monkeypatch.setenv("RETTO_B200_DETPRE_GENERIC", "1")
    (h, w), (nh, nw) = case
    rng = np.random.default_rng(h + 3 * w + nh)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = O.thumbnail(img, nh, nw)
    g = _t(img)
    torch.cuda.synchronize()
    out = ctx.thumbnail([g], [(nh, nw)])[0]
    ctx.sync()
    assert np.array_equal(out.cpu().numpy(), ref)
