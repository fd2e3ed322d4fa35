"""retto-cli entry point (retto-cli/src/main.rs:18-95) over the B200 path: same flags, same log lines, plus
`--device b200`, `--gpus N` and `--batch-pages`.

    python -m retto_b200.cli -i <file or dir> --device b200 [--gpus N] [--worker module:factory]

The reference decodes every file with `image::load_from_memory(..).to_rgb8()` (image_helper.rs:34-44) and runs
`session.run` on it one at a time (main.rs:80-86).  Here the files are decoded on the host (Pillow; PNG is lossless, so
the pages are identical to the reference's; JPEG decoders differ in IDCT rounding — decode is outside the device
path), batched, and sharded per page over the GPUs (one process-local context per GPU, LPT by H*W,
results reported in file order).  The DBNet / classifier / SVTR forward passes are NOT part of this package: they come
from a worker object with `det/cls/rec` methods taking and returning lists of CUDA tensors (retto_b200.session.
RettoWorker).  `--worker module:factory` names a callable returning one (called with device_id and the CLI namespace);
without it an onnxruntime worker bound through IoBinding is built from the three model paths — onnxruntime is not in this
image, so that branch raises a clear error here.
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


class OrtIoBindingWorker:
    """RettoOrtWorker (worker/ort_worker.rs:188-221) with device-resident inputs/outputs bound through IoBinding."""

    def __init__(self, device_id: int, det: str, cls: str, rec: str):
        try:
            import onnxruntime as ort  # noqa: F401
        except ImportError as e:  # pragma: no cover - onnxruntime is not in this image
            raise SystemExit("retto_b200.cli: onnxruntime is not installed; pass --worker module:factory for the forward passes") from e
        import onnxruntime as ort
        import torch
        self.torch, self.device_id = torch, device_id
        prov = [("CUDAExecutionProvider", {"device_id": device_id})]
        self.sess = [ort.InferenceSession(p, providers=prov) for p in (det, cls, rec)]

    def _run(self, k: int, xs):
        """one session.run per tensor, input and output bound to device memory (no host copies, unlike the six
        copies per stage of ort_worker.rs:188-221); the output buffer is a torch tensor of the stage's known shape"""
        torch = self.torch
        s = self.sess[k]
        outs = []
        for x in xs:
            x = x.contiguous()
            if k == 0:
                shape = (1, 1, x.shape[2], x.shape[3])
            elif k == 1:
                shape = (x.shape[0], 2)
            else:
                shape = (x.shape[0], x.shape[3] // 8, int(s.get_outputs()[0].shape[-1]))
            y = torch.empty(shape, dtype=torch.float32, device=x.device)
            b = s.io_binding()
            b.bind_input(s.get_inputs()[0].name, "cuda", self.device_id, np.float32, tuple(x.shape), x.data_ptr())
            b.bind_output(s.get_outputs()[0].name, "cuda", self.device_id, np.float32, shape, y.data_ptr())
            s.run_with_iobinding(b)
            outs.append(y)
        return outs

    def det(self, xs):
        return self._run(0, xs)

    def cls(self, xs):
        return self._run(1, xs)

    def rec(self, xs):
        return self._run(2, xs)


def make_worker(args, device_id: int):
    if args.worker:
        mod, _, fn = args.worker.partition(":")
        return getattr(importlib.import_module(mod), fn or "make_worker")(device_id, args)
    return OrtIoBindingWorker(device_id, args.det_model_path, args.cls_model_path, args.rec_model_path)


def fmt_debug(res) -> Sequence[str]:
    """the three tracing lines of RettoSession::run (session.rs:114-122), Debug-formatted like the reference's structs"""
    det = ", ".join("DetProcessorInnerResult { boxes: PointBox { inner: [%s] }, score: %r }" % (
        ", ".join("Point { x: %.1f, y: %.1f }" % (float(p[0]), float(p[1])) for p in d.boxes), float(np.float32(d.score))) for d in res.det_result)
    cls = ", ".join("ClsProcessorSingleResult { label: ClsPostProcessLabel { label: %d, score: %r } }" % (c.label, float(np.float32(c.score))) for c in res.cls_result)
    rec = ", ".join("RecProcessorSingleResult { text: %s, score: %r }" % ('"' + r.text.replace('"', '\\"') + '"', float(np.float32(r.score))) for r in res.rec_result)
    return ("Det result: DetProcessorResult([%s])" % det, "Cls result: ClsProcessorResult([%s])" % cls, "Rec result: RecProcessorResult([%s])" % rec)


def run(args) -> int:
    import json
    from .session import RecProcessorConfig, RettoSession, RettoSessionConfig
    from .shard import shard_indices
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
    pages = [decode_rgb8(f) for f in files]
    n_gpus = max(1, args.gpus)
    shards = shard_indices([p.shape[0] * p.shape[1] for p in pages], n_gpus)
    results = [None] * len(files)

    def work(rank: int):
        idx = shards[rank]
        if not idx:
            return
        dev = args.device_id + rank
        cfg = RettoSessionConfig(rec_processor_config=RecProcessorConfig(character_source=dict_text), device_id=dev)
        sess = RettoSession(cfg, worker=make_worker(args, dev))
        for b0 in range(0, len(idx), max(1, args.batch_pages)):
            chunk = idx[b0:b0 + max(1, args.batch_pages)]
            for i, r in zip(chunk, sess.run_pages([pages[i] for i in chunk])):
                results[i] = r
        sess.ctx.close()

    if n_gpus == 1:
        work(0)
    else:
        import threading
        ths = [threading.Thread(target=work, args=(r,)) for r in range(n_gpus)]   # the C calls release the GIL
        [t.start() for t in ths]
        [t.join() for t in ths]
    dt = time.perf_counter() - t0
    for f, r in zip(files, results):
        for line in fmt_debug(r):
            print(line)
        if args.json:
            print(json.dumps({"file": f, **r.to_json()}, ensure_ascii=False))
    print("Successfully processed %d images, avg time: %.2fms" % (len(files), 1000.0 * dt / len(files)))
    return 0


def main(argv=None) -> int:
    return run(build_parser().parse_args(argv))


if __name__ == "__main__":
    sys.exit(main())
