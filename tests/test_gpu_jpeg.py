"""GPU parity of the device image decode (csrc/jpeg_decode.cu) and of the entry points built on it:
decoded pixels bit-exact vs the oracle AND vs Pillow (libjpeg-turbo); RettoSession.run on file bytes == run on decoded RGB;
per-stage callbacks fire in the reference's order; run_pages_multi == run_pages."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _enc(a, **kw):
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(a).save(b, "JPEG", **kw)
    return b.getvalue()


def _pil(data):
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


def test_decode_matrix(ctx):
    """sub-samplings x qualities x {no DRI, optimised tables, restart per MCU row, restart every 3 MCUs} x odd sizes, one batch"""
    import cv2
    from PIL import Image
    from oracle import oracle as O
    from tools.synth import gen_page
    rng = np.random.default_rng(1)
    imgs = [gen_page(5, 333, 517, n_lines=(4, 8))[0], rng.integers(0, 256, (123, 77, 3), dtype=np.uint8),
            cv2.GaussianBlur(rng.integers(0, 256, (200, 310, 3), dtype=np.uint8), (0, 0), 3), rng.integers(0, 256, (9, 5, 3), dtype=np.uint8),
            gen_page(6, 736, 992, n_lines=(6, 12))[0], rng.integers(0, 256, (20, 4, 3), dtype=np.uint8), rng.integers(0, 256, (3, 2, 3), dtype=np.uint8)]
    files = []
    for img in imgs:
        for sub in (0, 1, 2):
            for q in (35, 90, 100):
                for kw in (dict(), dict(optimize=True), dict(restart_marker_rows=1), dict(restart_marker_blocks=3)):
                    files.append(_enc(img, quality=q, subsampling=sub, **kw))
    files.append(_enc(np.asarray(Image.fromarray(imgs[0]).convert("L")), quality=80))
    files.append(_enc(np.asarray(Image.fromarray(imgs[4]).convert("L")), quality=80, restart_marker_rows=2))
    outs, status = ctx.decode_images(files)
    assert all(s == 0 for s in status)
    for k, (f, t) in enumerate(zip(files, outs)):
        got = t.cpu().numpy()
        assert np.array_equal(got, _pil(f)), k
        if k % 7 == 0:
            assert np.array_equal(got, O.jpeg_decode(f)), k


def test_decode_statuses(ctx):
    from retto_b200 import _lib
    from tools.synth import gen_page
    img = gen_page(7, 200, 300, n_lines=(3, 5))[0]
    good = _enc(img, quality=85, restart_marker_rows=1)
    prog = _enc(img, quality=85, progressive=True)
    png = io.BytesIO()
    from PIL import Image
    Image.fromarray(img).save(png, "PNG")
    broken = bytearray(good)
    pos = [i for i in range(600, len(broken) - 1) if broken[i] == 0xFF and 0xD0 <= broken[i + 1] <= 0xD7]
    broken[pos[2] + 1] = 0x00          # turn one restart marker into a stuffed byte: the marker count no longer matches DRI
    outs, status = ctx.decode_images([good, prog, png.getvalue(), bytes(broken), good[:300]])
    assert status[0] == 0 and np.array_equal(outs[0].cpu().numpy(), _pil(good))
    assert status[1] == _lib.ERR_UNSUPPORTED and status[2] == _lib.ERR_UNSUPPORTED
    assert status[3] == _lib.ERR_DECODE and status[4] == _lib.ERR_DECODE


def test_damaged_streams_terminate_and_do_not_hurt_their_neighbours(ctx):
    """bytes flipped inside the entropy-coded data, files cut short (also as the LAST file of the batch: its bit readers run off the
    end of the data) — with and without restart markers, i.e. through K-J2 and through the self-synchronising K-J2s: every decode
    returns (status 0 or ERR_DECODE, never a hang or a fault), and the intact files of the same batch still decode bit-exactly"""
    import torch
    from retto_b200 import _lib
    from tools.synth import gen_page
    rng = np.random.default_rng(11)
    img = gen_page(12, 480, 640, n_lines=(5, 9))[0]
    goods = [_enc(img, quality=85), _enc(img, quality=85, restart_marker_rows=1), _enc(img, quality=85, subsampling=0, restart_marker_blocks=4)]
    files = []
    for g in goods:
        sos = g.index(b"\xff\xda")
        for k in range(6):
            b = bytearray(g)
            for p in rng.integers(sos + 20, len(b) - 2, size=int(rng.integers(1, 24))):
                b[p] = int(rng.integers(0, 256))
            files.append(bytes(b))
        files.append(g[:sos + 14 + (len(g) - sos) // 3] + b"\xff\xd9")     # cut after a third of the scan, EOI appended
        files.append(g[:len(g) * 2 // 3])                                       # cut without EOI
    batch = [goods[0]] + files + [goods[1], goods[2][:len(goods[2]) // 2]]      # the last file of the batch is a truncated one
    outs, status = ctx.decode_images(batch)
    torch.cuda.synchronize()
    assert all(s in (0, _lib.ERR_DECODE) for s in status), status
    assert status[0] == 0 and np.array_equal(outs[0].cpu().numpy(), _pil(goods[0]))
    assert status[-2] == 0 and np.array_equal(outs[-2].cpu().numpy(), _pil(goods[1]))
    # the context is still healthy
    outs2, status2 = ctx.decode_images(goods)
    assert status2 == [0, 0, 0] and all(np.array_equal(o.cpu().numpy(), _pil(g)) for o, g in zip(outs2, goods))


def _worker():
    from retto_b200.session import CallableWorker
    from tools.demo_worker import StatelessWorker
    w = StatelessWorker()
    return w, CallableWorker(w.det, w.cls, w.rec)


def test_session_on_file_bytes_equals_session_on_pixels(ctx, synth_dict):
    """RettoSession::run's own input is the file (session.rs:75-79): bytes in == the same results as the decoded page in, and
    both equal the oracle pipeline on the oracle-decoded page"""
    from oracle import oracle as O
    from oracle.pipeline import run_page
    from retto_b200.session import RettoSession
    from tools.synth import gen_page
    ctx.dict_load(synth_dict)
    pages = [gen_page(40 + i, h, w, n_lines=(6, 12))[0] for i, (h, w) in enumerate([(900, 1200), (1280, 1280), (500, 640), (2300, 1700)])]
    files = [_enc(p, quality=90, subsampling=2, restart_marker_rows=1) for p in pages[:3]] + [_enc(pages[3], quality=85, subsampling=0)]
    w, cw = _worker()
    sess = RettoSession(worker=cw, ctx=ctx)
    a = sess.run_pages(files)
    b = sess.run_pages([_pil(f) for f in files])
    for k, (x, y) in enumerate(zip(a, b)):
        assert x.status == 0 and len(x.det_result) == len(y.det_result) > 0
        for i in range(len(x.det_result)):
            assert np.array_equal(x.det_result[i].boxes, y.det_result[i].boxes) and x.det_result[i].score == y.det_result[i].score
            assert x.cls_result[i].label == y.cls_result[i].label and x.rec_result[i].text == y.rec_result[i].text
        ref = run_page(O.jpeg_decode(files[k]), w, synth_dict)
        assert len(ref["boxes"]) == len(x.det_result)
        for i in range(len(ref["boxes"])):
            assert np.array_equal(x.det_result[i].boxes, ref["boxes"][i]) and x.rec_result[i].text == ref["rec"][i][0]


def test_encoded_pages_on_two_lanes_equal_one_unit(ctx, synth_dict):
    """file bytes in, units of 4 pages software-pipelined on two lanes (entropy phase once per call on the copy stream, pixel phase
    per unit on the lane's stream) == the same call as one unit; files with and without restart markers in one batch"""
    from retto_b200.session import RettoSession
    from tools.synth import gen_page
    ctx.dict_load(synth_dict)
    pages = [gen_page(60 + i, 480 + 32 * (i % 3), 640 + 48 * (i % 4), n_lines=(4, 8))[0] for i in range(14)]
    files = [_enc(p, quality=88, subsampling=2, **(dict(restart_marker_rows=1) if i % 2 else dict())) for i, p in enumerate(pages)]
    w, cw = _worker()
    sess = RettoSession(worker=cw, ctx=ctx)
    try:
        ctx.set_pipeline(1, 1 << 20)
        one = sess.run_pages(files)
        ctx.set_pipeline(2, 4)
        piped = sess.run_pages(files)
    finally:
        ctx.set_pipeline(0, 0)
    assert len(one) == len(piped) == 14 and sum(len(r.det_result) for r in one) > 14
    for x, y in zip(one, piped):
        assert x.status == y.status == 0 and len(x.det_result) == len(y.det_result)
        for i in range(len(x.det_result)):
            assert np.array_equal(x.det_result[i].boxes, y.det_result[i].boxes) and x.det_result[i].score == y.det_result[i].score
            assert x.cls_result[i].label == y.cls_result[i].label and x.rec_result[i].text == y.rec_result[i].text


def test_unsupported_file_fails_loudly(ctx, synth_dict):
    from retto_b200 import _lib
    from retto_b200._lib import RettoB200Error
    from retto_b200.session import RettoSession
    from tools.synth import gen_page
    ctx.dict_load(synth_dict)
    _, cw = _worker()
    sess = RettoSession(worker=cw, ctx=ctx)
    with pytest.raises(RettoB200Error) as e:
        sess.run(_enc(gen_page(3, 300, 400, n_lines=(3, 5))[0], progressive=True))
    assert e.value.status == _lib.ERR_UNSUPPORTED


def test_run_stream_is_incremental(ctx, synth_dict):
    """session.rs:98,101,104: Det is delivered before worker.cls runs, Cls before worker.rec, Rec last"""
    from retto_b200.session import CallableWorker, RettoSession
    from tools.demo_worker import StatelessWorker
    from tools.synth import gen_page
    ctx.dict_load(synth_dict)
    w = StatelessWorker()
    events = []

    def tap(name, f):
        def g(x):
            events.append(name)
            return f(x)
        return g

    sess = RettoSession(worker=CallableWorker(tap("fwd_det", w.det), tap("fwd_cls", w.cls), tap("fwd_rec", w.rec)), ctx=ctx)
    got = {}

    def sender(item):
        events.append("send_" + item[0])
        got[item[0]] = item[1]

    img = gen_page(44, 800, 1000, n_lines=(5, 9))[0]
    sess.run_stream(img, sender)
    order = [e for i, e in enumerate(events) if i == 0 or events[i - 1] != e]     # collapse the per-batch repeats
    assert order == ["fwd_det", "send_Det", "fwd_cls", "send_Cls", "fwd_rec", "send_Rec"], order
    ref = sess.run(img)
    assert len(got["Det"]) == len(ref.det_result) > 0
    assert all(np.array_equal(a.boxes, b.boxes) for a, b in zip(got["Det"], ref.det_result))
    assert [c.label for c in got["Cls"]] == [c.label for c in ref.cls_result]
    assert [r.text for r in got["Rec"]] == [r.text for r in ref.rec_result]
    # an empty page still reports its three (empty) stages
    events.clear()
    sess.run_stream(np.full((300, 400, 3), 255, np.uint8), sender)
    assert [e for e in events if e.startswith("send_")] == ["send_Det", "send_Cls", "send_Rec"] and got["Det"] == []


def _multi_case(devices, synth_dict):
    from retto_b200.api import Context
    from retto_b200.session import CallableWorker, RettoSession, run_pages_multi
    from tools.demo_worker import StatelessWorker
    from tools.synth import gen_page
    rng = np.random.default_rng(5)
    pages = []
    for i in range(14):
        h, w = int(rng.integers(400, 1500)), int(rng.integers(400, 1500))
        pages.append(gen_page(60 + i, h, w, n_lines=(3, 8))[0])
    files = [_enc(p, quality=88, restart_marker_rows=1) for p in pages]
    w = StatelessWorker()
    sessions = []
    try:
        for d in devices:
            c = Context(d)
            c.dict_load(synth_dict)
            sessions.append(RettoSession(worker=CallableWorker(w.det, w.cls, w.rec), ctx=c))
        for inputs in (pages, files):
            one = sessions[0].run_pages(inputs)
            many = run_pages_multi(sessions, inputs, chunk_pages=3)
            assert len(one) == len(many) == len(pages)
            for a, b in zip(one, many):
                assert a.status == b.status == 0 and len(a.det_result) == len(b.det_result)
                for x, y in zip(a.det_result, b.det_result):
                    assert np.array_equal(x.boxes, y.boxes) and x.score == y.score
                assert [r.text for r in a.rec_result] == [r.text for r in b.rec_result]
                assert [c.label for c in a.cls_result] == [c.label for c in b.cls_result]
        assert sum(len(r.det_result) for r in one) > 30
    finally:
        for s in sessions:
            s.ctx.close()


def test_run_pages_multi_two_contexts_one_gpu(synth_dict):
    """the shared-cursor driver with two contexts (two host threads) on ONE device: same results, page order kept"""
    _multi_case([0, 0], synth_dict)


def test_run_pages_multi_two_gpus(synth_dict):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _multi_case([0, 1], synth_dict)


def test_contexts_on_two_devices_from_one_thread(synth_dict):
    """every C-ABI entry makes its context's device current; one thread driving two devices in turn"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from retto_b200.api import Context
    from tools.synth import gen_probmap
    p = gen_probmap(3, 256, 256, k_range=(3, 6))
    c0, c1 = Context(0), Context(1)
    try:
        outs = []
        for rep in range(2):
            for c, dev in ((c0, 0), (c1, 1)):
                torch.cuda.set_device(1 - dev)          # the caller's current device is the OTHER one
                g = torch.from_numpy(p).to(f"cuda:{dev}")
                torch.cuda.synchronize(dev)
                outs.append(c.det_postprocess([g], [p.shape]))
                assert torch.cuda.current_device() == 1 - dev
        assert all(np.array_equal(o.boxes, outs[0].boxes) for o in outs) and len(outs[0].boxes) > 0
    finally:
        c0.close()
        c1.close()
        torch.cuda.set_device(0)


_ENTROPY_PATHS_SCRIPT = r'''
import io, sys
import numpy as np
sys.path.insert(0, ".")
import cv2
from PIL import Image, ImageFile
ImageFile.MAXBLOCK = 1 << 24          # optimised tables on noise images need a larger encoder buffer
from retto_b200.api import Context
from tools.synth import gen_page
def enc(a, **kw):
    b = io.BytesIO(); Image.fromarray(a).save(b, "JPEG", **kw); return b.getvalue()
rng = np.random.default_rng(3)
blank = np.full((640, 800, 3), 255, np.uint8); blank[300:340, 100:700] = 0          # long runs of DC-only MCUs around one bar
imgs = [gen_page(8, 1280, 1280)[0], gen_page(9, 700, 2000, n_lines=(3, 6))[0], blank,
        cv2.GaussianBlur(rng.integers(0, 256, (512, 768, 3), dtype=np.uint8), (0, 0), 2), rng.integers(0, 256, (300, 200, 3), dtype=np.uint8),
        rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)]
files, f420 = [], []
for img in imgs:
    for sub in (0, 1, 2):
        for kw in (dict(), dict(optimize=True), dict(restart_marker_rows=1), dict(restart_marker_rows=7), dict(restart_marker_blocks=5)):
            files.append(enc(img, quality=88, subsampling=sub, **kw))
            if sub == 2:
                f420.append(files[-1])
files.append(enc(np.asarray(Image.fromarray(imgs[0]).convert("L")), quality=92))
ctx = Context(0)
for batch in (files, f420):      # mixed samplings (generic colour kernel) / 4:2:0 pages of different sizes (flat-grid two-row kernel)
    outs, status = ctx.decode_images(batch)
    assert all(s == 0 for s in status), status
    for k, (f, t) in enumerate(zip(batch, outs)):
        ref = np.asarray(Image.open(io.BytesIO(f)).convert("RGB"))
        assert np.array_equal(t.cpu().numpy(), ref), k
print("OK", len(files))
'''


@pytest.mark.parametrize("env", [{"RETTO_B200_JPEG_SUB": "1"}, {"RETTO_B200_JPEG_NOSUB": "1"}, {"RETTO_B200_JPEG_NOSUB": "1", "RETTO_B200_JPEG_DENSE": "1"}, {}],
                         ids=["sub-sequences everywhere", "intervals only (tiered)", "intervals only (dense)", "default: per file"])
def test_entropy_paths_agree_with_libjpeg(env):
    """the three Huffman layouts — K-J2 tiered / dense (one thread per restart interval) and K-J2s (self-synchronising sub-sequences of
    long intervals) — are chosen per file from its mean interval length; the switches are read once per process, so each forced
    configuration decodes the same 91 files (with / without restart markers, three samplings, optimised tables, blank stretches,
    2000-px-wide rows, greyscale) in its own interpreter and compares with Pillow"""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", _ENTROPY_PATHS_SCRIPT], cwd=root, env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK 91" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
