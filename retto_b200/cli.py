"""retto-cli entry point (retto-cli/src/main.rs:18-95) over the B200 path: same flags, same log lines, plus
`--device b200`, `--gpus N` and `--batch-pages`.

    python -m retto_b200.cli -i <file or dir> --device b200 [--gpus N] [--worker module:factory]

The reference decodes every file with `image::load_from_memory(..).to_rgb8()` (image_helper.rs:34-44) and runs
`session.run` on it one at a time (main.rs:80-86).  Here baseline JPEG files are handed to the device as file bytes and decoded
there (csrc/jpeg_decode.cu, bit-exact with libjpeg-turbo); every other format is decoded on the host (Pillow; PNG is lossless, so
those pages are identical to the reference's).  Pages are batched and, with `--gpus N`, spread over N contexts by
retto_b200_run_pages_multi (one host thread per GPU, shared cursor over the pages sorted by H*W; results in file order).  The DBNet / classifier / SVTR forward passes are NOT part of this package: they come
from a worker object with `det/cls/rec` methods taking and returning lists of CUDA tensors (retto_b200.session.
RettoWorker).  `--worker module:factory` names a callable returning one (called with device_id and the CLI namespace);
`--worker standin` runs the torch stand-in networks (retto_b200.standin: the seam exercised end to end, not OCR); without the
flag an onnxruntime IoBinding worker (retto_b200.ort_worker, untested here: no onnxruntime in this image) is built from the model paths.
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys
import time
from typing import List, Sequence

import numpy as np


def build_parser() -> argparse.ArgumentParser:
    # program name "ratio-cli" (sic, main.rs:19)
    ap = argparse.ArgumentParser(prog="ratio-cli")
    ap.add_argument("--det-model-path", default="ch_PP-OCRv4_det_infer.onnx")
    ap.add_argument("--cls-model-path", default="ch_ppocr_mobile_v2.0_cls_infer.onnx")
    ap.add_argument("--rec-model-path", default="ch_PP-OCRv4_rec_infer.onnx")
    ap.add_argument("--rec-keys-path", default="ppocr_keys_v1.txt",
                    help="dictionary file; the reference accepts and ignores this flag (main.rs:27-28 vs :67-70), here it is read")
    ap.add_argument("-i", "--images", required=True, help="image file or directory (walked recursively)")
    ap.add_argument("--device", choices=["cpu", "cuda", "direct-ml", "b200"], default="b200")
    ap.add_argument("--device-id", type=int, default=0)
    ap.add_argument("--use-hf-hub", default="true", help="accepted for compatibility; there is no network access here")
    ap.add_argument("--gpus", type=int, default=1, help="page-shard the files over this many GPUs (b200 only)")
    ap.add_argument("--batch-pages", type=int, default=64, help="pages per retto_b200_run_pages call")
    ap.add_argument("--worker", default=None, help="module:factory returning an object with det/cls/rec (lists of CUDA tensors)")
    ap.add_argument("--json", action="store_true", help="also print one serde-shaped JSON object per image (fe/index.ts:5-42)")
    return ap


def find_files(root: str) -> List[str]:
    """WalkDir::new(images) filtered to files (main.rs:72-77); sorted so that the report order is reproducible"""
    if os.path.isfile(root):
        return [root]
    out = []
    for d, _, fs in os.walk(root):
        for f in fs:
            out.append(os.path.join(d, f))
    return sorted(out)


def decode_rgb8(path: str) -> np.ndarray:
    """ImageHelper::new_from_raw_img_flow (image_helper.rs:34-44): any supported format -> RGB8 HWC"""
    from PIL import Image
    with Image.open(path) as im:
        return np.ascontiguousarray(np.asarray(im.convert("RGB"), dtype=np.uint8))


class StandInWorker:
    """`--worker standin`: the torch stand-in networks of retto_b200.standin (random-init, PP-OCRv4-mobile I/O shape) run on the
    context's own tensors.  It exercises the whole path end to end on a machine without onnxruntime / weights — the boxes and
    strings it produces are NOT OCR results."""

    def __init__(self, device_id: int, n_classes: int):
        import torch
        from .standin import StandInNets
        self.torch = torch
        self.nets = StandInNets(torch, f"cuda:{device_id}", n_classes)

    def det(self, xs):
        with self.torch.no_grad():
            return self.nets.run(0, xs)

    def cls(self, xs):
        with self.torch.no_grad():
            return self.nets.run(1, xs)

    def rec(self, xs):
        with self.torch.no_grad():
            return self.nets.run(2, xs)


def make_worker(args, device_id: int, n_classes: int):
    if args.worker == "standin":
        return StandInWorker(device_id, n_classes)
    if args.worker:
        mod, _, fn = args.worker.partition(":")
        return getattr(importlib.import_module(mod), fn or "make_worker")(device_id, args)
    from .ort_worker import OrtIoBindingWorker
    return OrtIoBindingWorker(device_id, args.det_model_path, args.cls_model_path, args.rec_model_path)


def _f32(v) -> str:
    """Rust's `{:?}` of an f32: shortest digits that round-trip, always with a decimal point or exponent"""
    x = np.float32(v)
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "inf" if x > 0 else "-inf"
    a = abs(float(x))
    if a != 0.0 and (a < 1e-5 or a >= 1e16):
        m, e = np.format_float_scientific(x, unique=True, trim="-", exp_digits=1).split("e")
        return "%se%d" % (m, int(e))
    return np.format_float_positional(x, unique=True, trim="0")


def _str_debug(t: str) -> str:
    """Rust's `{:?}` of a str (char::escape_debug): quotes, backslash and control characters escaped, printable Unicode kept"""
    out = []
    for ch in t:
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif ch == "\0":
            out.append("\\0")
        elif ch.isprintable():
            out.append(ch)
        else:
            out.append("\\u{%x}" % ord(ch))
    return '"' + "".join(out) + '"'


def fmt_debug(res) -> Sequence[str]:
    """the three tracing lines of RettoSession::run (session.rs:114-122) in the reference's Debug shapes: PointBox has a custom impl
    (points.rs:70-82: tl / tr / br / bl), coordinates are OrderedFloat<f32>, labels u16, strings escape_debug"""
    def pt(p):
        return "Point { x: OrderedFloat(%s), y: OrderedFloat(%s) }" % (_f32(p[0]), _f32(p[1]))
    det = ", ".join("DetProcessorInnerResult { boxes: PointBox { tl: %s, tr: %s, br: %s, bl: %s }, score: %s }" % (
        pt(d.boxes[0]), pt(d.boxes[1]), pt(d.boxes[2]), pt(d.boxes[3]), _f32(d.score)) for d in res.det_result)
    cls = ", ".join("ClsProcessorSingleResult { label: ClsPostProcessLabel { label: %d, score: %s } }" % (c.label, _f32(c.score)) for c in res.cls_result)
    rec = ", ".join("RecProcessorSingleResult { text: %s, score: %s }" % (_str_debug(r.text), _f32(r.score)) for r in res.rec_result)
    return ("Det result: DetProcessorResult([%s])" % det, "Cls result: ClsProcessorResult([%s])" % cls, "Rec result: RecProcessorResult([%s])" % rec)


def run(args) -> int:
    import json
    from . import _lib
    from .api import image_info
    from .session import RecProcessorConfig, RettoSession, RettoSessionConfig, run_pages_multi
    if args.device != "b200":
        raise SystemExit(f"retto_b200.cli: --device {args.device} is the reference's own ORT path; this package implements --device b200 only")
    files = find_files(args.images)
    print(f"Found {len(files)} files, processing...")
    if not files:
        return 0
    dict_text = open(args.rec_keys_path, encoding="utf-8").read() if os.path.exists(args.rec_keys_path) else None
    if dict_text is None:
        raise SystemExit(f"retto_b200.cli: dictionary {args.rec_keys_path} not found (--rec-keys-path)")
    t0 = time.perf_counter()
    # the reference hands the file bytes to session.run (main.rs:83-84).  Baseline JPEG files go to the device as they are (decoded
    # by csrc/jpeg_decode.cu); every other format the `image` crate reads is decoded on the host and passed as RGB8
    blobs = [open(f, "rb").read() for f in files]
    on_dev = [i for i, b in enumerate(blobs) if image_info(b).status == _lib.OK]
    on_host = [i for i in range(len(files)) if i not in set(on_dev)]
    inputs = {i: blobs[i] for i in on_dev}
    for i in on_host:
        inputs[i] = decode_rgb8(files[i])
    n_gpus = max(1, args.gpus)
    sessions = []
    results = [None] * len(files)
    try:
        for rank in range(n_gpus):
            dev = args.device_id + rank
            cfg = RettoSessionConfig(rec_processor_config=RecProcessorConfig(character_source=dict_text), device_id=dev)
            s = RettoSession(cfg)
            s.worker = make_worker(args, dev, s.ctx.dict_size)
            sessions.append(s)
        for group in (on_dev, on_host):     # one call may not mix encoded and decoded pages
            for b0 in range(0, len(group), max(1, args.batch_pages) * n_gpus):
                chunk = group[b0:b0 + max(1, args.batch_pages) * n_gpus]
                batch = [inputs[i] for i in chunk]
                # several GPUs: retto_b200_run_pages_multi (one host thread per context, shared cursor over the LPT-sorted pages);
                # a worker exception on any rank is re-raised here
                out = run_pages_multi(sessions, batch) if n_gpus > 1 else sessions[0].run_pages(batch)
                for i, r in zip(chunk, out):
                    results[i] = r
    finally:
        for s in sessions:
            s.ctx.close()
    dt = time.perf_counter() - t0
    bad = 0
    for f, r in zip(files, results):
        for line in fmt_debug(r):
            print(line)
        if r.status != _lib.OK:   # soft per-page statuses (capacity, degenerate quad): reported, and the run exits non-zero
            bad += 1
            print(f"WARN {f}: page status {r.status} ({_lib.STATUS_NAMES.get(r.status, '?')}): results of this page are incomplete", file=sys.stderr)
        if args.json:
            print(json.dumps({"file": f, **r.to_json()}, ensure_ascii=False))
    print("Successfully processed %d images, avg time: %.2fms" % (len(files) - bad, 1000.0 * dt / len(files)))
    if on_dev:
        print(f"({len(on_dev)} JPEG files decoded on the device, {len(on_host)} files decoded on the host)", file=sys.stderr)
    return 1 if bad else 0


def main(argv=None) -> int:
    return run(build_parser().parse_args(argv))


if __name__ == "__main__":
    sys.exit(main())
