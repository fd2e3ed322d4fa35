"""The JPEG oracle (oracle/jpeg_oracle.cpp: a restatement of libjpeg-turbo's default decode path) PINNED against the two real
libjpeg-turbo builds on this box — Pillow's and OpenCV's — on every sub-sampling, odd sizes, restart intervals, optimised Huffman
tables and quality levels; plus the host-only header parse of the C ABI (retto_b200_image_info)."""
import io

import numpy as np
import pytest

from oracle import oracle as O


def _enc(a, **kw):
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(a).save(b, "JPEG", **kw)
    return b.getvalue()


def _images():
    import cv2
    from tools.synth import gen_page
    rng = np.random.default_rng(0)
    return {
        "page": gen_page(5, 333, 517, n_lines=(4, 8))[0],
        "noise": rng.integers(0, 256, (123, 77, 3), dtype=np.uint8),
        "smooth": cv2.GaussianBlur(rng.integers(0, 256, (200, 310, 3), dtype=np.uint8), (0, 0), 3),
        "tiny": rng.integers(0, 256, (9, 5, 3), dtype=np.uint8),
        "w4": rng.integers(0, 256, (20, 4, 3), dtype=np.uint8),     # down-sampled width 2: libjpeg replicates instead of the triangle filter
        "w2": rng.integers(0, 256, (3, 2, 3), dtype=np.uint8),
    }


VARIANTS = [dict(), dict(optimize=True), dict(restart_marker_rows=1), dict(restart_marker_blocks=3)]


@pytest.mark.parametrize("sub", [0, 1, 2])
def test_oracle_equals_libjpeg_turbo(sub):
    from PIL import Image
    import cv2
    n = 0
    for name, img in _images().items():
        for q in (30, 75, 90, 100):
            for kw in VARIANTS:
                data = _enc(img, quality=q, subsampling=sub, **kw)
                ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
                ref2 = cv2.cvtColor(cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
                got = O.jpeg_decode(data)
                assert np.array_equal(ref, ref2)
                assert np.array_equal(got, ref), (name, sub, q, kw)
                n += 1
    assert n == 96


def test_oracle_grayscale_and_unsupported():
    from PIL import Image
    img = _images()["page"]
    g = _enc(np.asarray(Image.fromarray(img).convert("L")), quality=80)
    assert np.array_equal(O.jpeg_decode(g), np.asarray(Image.open(io.BytesIO(g)).convert("RGB")))
    with pytest.raises(ValueError) as e:
        O.jpeg_decode(_enc(img, quality=80, progressive=True))
    assert e.value.args[0] == 2
    with pytest.raises(ValueError):
        O.jpeg_decode(b"\x89PNG\r\n\x1a\n" + b"\0" * 64)


def test_image_info_header_parse():
    """retto_b200_image_info is host-only (marker parse), so it runs without a GPU"""
    from retto_b200 import _lib
    from retto_b200.api import image_info
    img = _images()["page"]
    for sub, code in ((0, 0), (1, 1), (2, 2)):
        for kw, ri in ((dict(), 0), (dict(restart_marker_rows=1), None), (dict(restart_marker_blocks=3), 3)):
            info = image_info(_enc(img, quality=85, subsampling=sub, **kw))
            assert info.status == 0 and (info.h, info.w) == img.shape[:2] and info.format == 1 and info.components == 3
            assert info.subsampling == code
            if ri is None:
                assert info.restart_interval == (img.shape[1] + (16 if sub else 8) - 1) // (16 if sub else 8)   # one MCU row
            else:
                assert info.restart_interval == ri
    from PIL import Image
    g = image_info(_enc(np.asarray(Image.fromarray(img).convert("L")), quality=80))
    assert g.status == 0 and g.components == 1 and g.subsampling == 3
    assert image_info(_enc(img, quality=80, progressive=True)).status == _lib.ERR_UNSUPPORTED
    assert image_info(b"\x89PNG\r\n\x1a\n" + b"\0" * 64).status == _lib.ERR_UNSUPPORTED
    assert image_info(b"garbage that is no image at all").status == _lib.ERR_DECODE
    data = _enc(img, quality=85)
    assert image_info(data[:200]).status == _lib.ERR_DECODE      # truncated inside the tables


def test_image_info_survives_mutated_headers():
    """the marker parser reads attacker-controlled lengths: 3000 mutations of valid files (flipped bytes / cut / grown inside the
    header region, segment lengths forged) must each give a status — never a crash, never a read past the buffer (the buffers sit at
    the END of an mmap'd page run followed by a PROT_NONE guard page, so an over-read segfaults the test process)"""
    import ctypes as C
    import mmap
    from retto_b200 import _lib
    img = _images()["page"]
    seeds = [_enc(img, quality=85), _enc(img, quality=60, subsampling=0, restart_marker_blocks=3, optimize=True),
             _enc(np.asarray(__import__("PIL.Image", fromlist=["Image"]).fromarray(img).convert("L")), quality=80)]
    L = _lib.lib()
    libc = C.CDLL(None, use_errno=True)
    libc.mprotect.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    PAGE = mmap.PAGESIZE
    span = 64 * PAGE
    mm = mmap.mmap(-1, span + PAGE)
    base = C.addressof(C.c_char.from_buffer(mm))
    assert libc.mprotect(base + span, PAGE, 0) == 0            # guard page behind the buffer
    rng = np.random.default_rng(5)
    seen = set()
    info = _lib.ImageInfo()
    for it in range(3000):
        src = bytearray(seeds[it % len(seeds)])
        hdr = src.index(b"\xff\xda") + 14
        kind = it % 5
        if kind == 0:
            for p in rng.integers(2, hdr, size=int(rng.integers(1, 6))):
                src[p] = int(rng.integers(0, 256))
        elif kind == 1:
            src = src[:int(rng.integers(1, hdr + 40))]
        elif kind == 2:                                       # forge a segment length
            pos = [i for i in range(2, hdr - 3) if src[i] == 0xFF and src[i + 1] in (0xDB, 0xC4, 0xC0, 0xDD, 0xDA, 0xE0)]
            p = pos[int(rng.integers(0, len(pos)))]
            src[p + 2], src[p + 3] = int(rng.integers(0, 256)), int(rng.integers(0, 256))
        elif kind == 3:
            p = int(rng.integers(2, hdr))
            src[p:p] = bytes(rng.integers(0, 256, size=int(rng.integers(1, 9)), dtype=np.uint8))
        else:
            p = int(rng.integers(2, hdr))
            del src[p:p + int(rng.integers(1, 9))]
        n = min(len(src), span)
        C.memmove(base + span - n, bytes(src[:n]), n)          # the file ends exactly at the guard page
        st = L.retto_b200_image_info(C.c_void_p(base + span - n), n, C.byref(info))
        assert st in (_lib.OK, _lib.ERR_DECODE, _lib.ERR_UNSUPPORTED), st
        seen.add(st)
        if st == _lib.OK:
            assert 0 < info.h <= 65535 and 0 < info.w <= 65535 and info.components in (1, 3)
    assert seen == {_lib.OK, _lib.ERR_DECODE, _lib.ERR_UNSUPPORTED}
    libc.mprotect(base + span, PAGE, 3)
