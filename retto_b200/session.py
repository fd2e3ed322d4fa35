"""Host-side mirror of retto-core's session surface for the B200 path (Python because no Rust toolchain
exists in this image; the orchestration itself is C++ inside libretto_b200.so, `retto_b200_run_pages`).

Names, field meanings and defaults follow the reference:
  RettoSession::{new, run, run_stream}      retto-core/src/session.rs:62,108,133
  RettoSessionConfig                         session.rs:17-40
  Det/Cls/RecProcessorConfig                 det_processor.rs:44-93, cls_processor.rs:12-36, rec_processor.rs:100-136
  RettoWorkerResult / RettoWorkerStageResult session.rs:42-56
  RettoInnerWorker::{det, cls, rec}          worker.rs:69-73   (here: device-resident tensors, batched)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import FORWARD_FN, STAGE_FN, Page, Results, RettoB200Error, StageResult, Tensor
from .api import Context, default_config


# ---- configs (defaults verbatim from the reference) -----------------------------------------------
@dataclass
class DetProcessorConfig:
    limit_side_len: int = 736
    limit_type: str = "Min"          # LimitType::Min | Max
    mean: Sequence[float] = (0.5, 0.5, 0.5)
    std: Sequence[float] = (0.5, 0.5, 0.5)
    scale: float = float(np.float32(1.0) / np.float32(255.0))
    threch: float = 0.3              # sic (det_processor.rs:56)
    box_thresh: float = 0.5
    max_candidates: int = 1000       # never read by the reference
    unclip_ratio: float = 1.6
    use_dilation: bool = True        # never read by the reference; dilation_kernel drives it
    score_mode: str = "Fast"         # never read by the reference
    min_mini_box_size: int = 3
    dilation_kernel: Optional[Sequence[Sequence[int]]] = ((1, 1), (1, 1))


@dataclass
class ClsProcessorConfig:
    image_shape: Sequence[int] = (3, 48, 192)
    batch_num: int = 6
    thresh: float = 0.9
    label: Sequence[int] = (0, 180)


@dataclass
class RecProcessorConfig:
    character_source: Optional[str] = None   # dictionary TEXT (RecCharacterDictProvider::OutSide(Blob)); path loading is the caller's
    image_shape: Sequence[int] = (3, 48, 320)
    batch_num: int = 6


@dataclass
class RettoSessionConfig:
    worker_config: object = None
    max_side_len: int = 2000
    min_side_len: int = 30
    det_processor_config: DetProcessorConfig = field(default_factory=DetProcessorConfig)
    cls_processor_config: ClsProcessorConfig = field(default_factory=ClsProcessorConfig)
    rec_processor_config: RecProcessorConfig = field(default_factory=RecProcessorConfig)
    device_id: int = 0

    def to_c(self):
        c = default_config()
        d, k, r = self.det_processor_config, self.cls_processor_config, self.rec_processor_config
        c.max_side_len, c.min_side_len = self.max_side_len, self.min_side_len
        c.det_limit_side_len = d.limit_side_len
        c.det_limit_type = 1 if str(d.limit_type).lower() == "max" else 0
        for i in range(3):
            c.det_mean[i], c.det_std[i] = d.mean[i], d.std[i]
        c.det_scale = d.scale
        c.det_thresh, c.det_box_thresh = d.threch, d.box_thresh
        c.det_max_candidates, c.det_unclip_ratio = d.max_candidates, d.unclip_ratio
        c.det_use_dilation, c.det_min_mini_box_size = int(d.use_dilation), d.min_mini_box_size
        if d.dilation_kernel is None:
            c.det_dilation_2x2 = 0
        else:
            k2 = np.asarray(d.dilation_kernel)
            if k2.shape != (2, 2) or not (k2 != 0).all():
                raise RettoB200Error(_lib.ERR_UNSUPPORTED, "only the reference's default 2x2 all-ones dilation kernel (or None) is implemented")
            c.det_dilation_2x2 = 1
        for i in range(3):
            c.cls_image_shape[i], c.rec_image_shape[i] = k.image_shape[i], r.image_shape[i]
        c.cls_batch_num, c.cls_thresh = k.batch_num, k.thresh
        c.cls_label[0], c.cls_label[1] = k.label[0], k.label[1]
        c.rec_batch_num = r.batch_num
        return c


# ---- results (det_processor.rs:104-113, cls_processor.rs:43-66, rec_processor.rs:150-160) ---------------
@dataclass
class DetProcessorInnerResult:
    boxes: np.ndarray   # [4,2] f32: tl, tr, br, bl (PointBox)
    score: float


@dataclass
class ClsPostProcessLabel:
    label: int
    score: float


@dataclass
class RecProcessorSingleResult:
    text: str
    score: float


@dataclass
class RettoWorkerResult:
    det_result: List[DetProcessorInnerResult]
    cls_result: List[ClsPostProcessLabel]
    rec_result: List[RecProcessorSingleResult]
    status: int = 0

    def to_json(self):
        """serde JSON shape consumed by retto-wasm/fe/index.ts:5-42"""
        return {
            "det_result": [{"boxes": {"inner": [{"x": float(p[0]), "y": float(p[1])} for p in d.boxes]}, "score": float(d.score)} for d in self.det_result],
            "cls_result": [{"label": {"label": int(c.label), "score": float(c.score)}} for c in self.cls_result],
            "rec_result": [{"text": r.text, "score": float(r.score)} for r in self.rec_result],
        }


# ---- device tensor plumbing -----------------------------------------------------------------------------------
class _DevArray:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}


def _wrap(t: Tensor, device):
    import torch
    shape = [t.shape[i] for i in range(t.ndim)]
    if int(np.prod(shape)) == 0:
        return torch.empty(shape, dtype=torch.float32, device=device)
    return torch.as_tensor(_DevArray(t.d_data, shape), device=device)


class RettoWorker:
    """RettoInnerWorker (worker.rs:69-73) with device-resident, batched tensors.  Each method receives a
    list of torch CUDA tensors (views of the context's buffers, valid during the call) and returns a list of
    contiguous f32 CUDA tensors: det [1,3,H,W]->[1,1,H,W]; cls [n,3,48,192]->[n,2]; rec [n,3,48,W]->[n,T,C]."""

    def det(self, xs):
        raise NotImplementedError

    def cls(self, xs):
        raise NotImplementedError

    def rec(self, xs):
        raise NotImplementedError


class CallableWorker(RettoWorker):
    """Worker from three numpy callables (tests): f(np.ndarray) -> np.ndarray, applied per tensor on the host."""

    def __init__(self, det: Callable, cls: Callable, rec: Callable):
        self._f = (det, cls, rec)
        self.seen = ([], [], [])

    def _run(self, k, xs):
        import torch
        outs = []
        for x in xs:
            a = x.detach().cpu().numpy()
            self.seen[k].append(a)
            outs.append(torch.from_numpy(np.ascontiguousarray(self._f[k](a), dtype=np.float32)).to(x.device))
        return outs

    def det(self, xs):
        return self._run(0, xs)

    def cls(self, xs):
        return self._run(1, xs)

    def rec(self, xs):
        return self._run(2, xs)


class RettoSession:
    """RettoSession<W> (session.rs:9-13).  `run` takes a decoded RGB page (HWC u8): image decode
    (image_helper.rs:34-44) is out of scope of the device path."""

    def __init__(self, cfg: Optional[RettoSessionConfig] = None, worker: Optional[RettoWorker] = None, ctx: Optional[Context] = None):
        self.config = cfg or RettoSessionConfig()
        self.worker = worker
        self.ctx = ctx or Context(self.config.device_id, self.config.to_c())
        if self.config.rec_processor_config.character_source is not None:
            self.ctx.dict_load(self.config.rec_processor_config.character_source)   # RecCharacter::new (session.rs:65-66)
        self._keep = None
        self._err = None
        self._cb = FORWARD_FN(self._forward)

    # forward seam: stage 0 det, 1 cls, 2 rec
    def _forward(self, user, stage, n, inputs, outputs, stream):
        import torch
        try:
            dev = f"cuda:{self.ctx.device_id}"
            # the tensors of this call are ordered on `stream` (the lane of the unit being processed, include/retto_b200.h)
            ts = self.ctx.torch_stream() if (not stream or stream == self.ctx.stream_ptr) else torch.cuda.ExternalStream(int(stream), device=dev)
            with torch.cuda.stream(ts):
                xs = [_wrap(inputs[i], dev) for i in range(n)]
                ys = (self.worker.det, self.worker.cls, self.worker.rec)[stage](xs)
                if len(ys) != n:
                    raise ValueError(f"worker returned {len(ys)} tensors for {n} inputs")
                keep = []
                for i, y in enumerate(ys):
                    y = y.contiguous()
                    if y.dtype != torch.float32:
                        y = y.float()
                    keep.append(y)
                    outputs[i].d_data = y.data_ptr()
                    outputs[i].ndim = y.dim()
                    for k in range(y.dim()):
                        outputs[i].shape[k] = y.shape[k]
                lst = self._keep.setdefault((stage, int(stream or 0)), [])   # a lane has one unit in flight: keep its outputs alive
                lst.append(keep)
                del lst[:-2]
            return 0
        except Exception as e:  # surfaced as ERR_WORKER
            self._err = e
            return 1

    def _pages(self, images: Sequence, on_device: bool):
        """retto_b200_page array for a list of pages: HWC u8 numpy arrays (host RGB), torch CUDA tensors (on_device) or `bytes`
        objects = the image FILE as RettoSession::run receives it (session.rs:75-79), decoded on the device"""
        n = len(images)
        pages = (Page * max(n, 1))()
        hold = []
        for i, im in enumerate(images):
            if isinstance(im, (bytes, bytearray, memoryview)):
                a = np.frombuffer(im, np.uint8)
                hold.append(a)
                pages[i] = Page(a.ctypes.data, 0, 0, _lib.PAGE_HOST_ENCODED, len(a))
            elif on_device:
                assert im.is_cuda and im.is_contiguous()
                pages[i] = Page(im.data_ptr(), im.shape[0], im.shape[1], _lib.PAGE_DEVICE_RGB, 0)
            else:
                a = np.ascontiguousarray(im, dtype=np.uint8)
                hold.append(a)
                pages[i] = Page(a.ctypes.data, a.shape[0], a.shape[1], _lib.PAGE_HOST_RGB, 0)
        return pages, hold

    @staticmethod
    def _collect(res, n) -> List[RettoWorkerResult]:
        out = []
        text = C.string_at(res.text, res.text_offsets[res.n_lines]) if res.n_lines else b""
        for p in range(n):
            pr = res.pages[p]
            det, cls, rec = [], [], []
            for k in range(pr.first_line, pr.first_line + pr.n_lines):
                b = res.boxes[k]
                det.append(DetProcessorInnerResult(np.array(b.xy[:], np.float32).reshape(4, 2), float(np.float32(b.score))))
                cls.append(ClsPostProcessLabel(int(res.cls[k].label), float(np.float32(res.cls[k].score))))
                rec.append(RecProcessorSingleResult(text[res.text_offsets[k]:res.text_offsets[k + 1]].decode("utf-8"), float(np.float32(res.rec_scores[k]))))
            out.append(RettoWorkerResult(det, cls, rec, pr.status))
        return out

    def run_pages(self, images: Sequence, on_device: bool = False) -> List[RettoWorkerResult]:
        """process_pipeline for a batch of pages (see _pages for the accepted page kinds)."""
        n = len(images)
        pages, hold = self._pages(images, on_device)
        res = Results()
        self._keep = {}
        self._err = None
        if on_device:
            # device pages were produced on torch's current stream; the context works on its own (see Context._ordered)
            with self.ctx._ordered():
                st = self.ctx._L.retto_b200_run_pages(self.ctx.handle, pages, n, self._cb, None, C.byref(res))
        else:
            st = self.ctx._L.retto_b200_run_pages(self.ctx.handle, pages, n, self._cb, None, C.byref(res))
        if self._err is not None:
            raise self._err
        if st not in (_lib.OK, _lib.ERR_DEGENERATE_QUAD, _lib.ERR_CAPACITY) or (st != _lib.OK and res.n_pages == 0):
            self.ctx._check(st)
        del hold
        return self._collect(res, n)

    def run(self, image) -> RettoWorkerResult:
        """RettoSession::run (session.rs:108-131); `image` may be the file bytes (the reference's own input) or decoded RGB"""
        return self.run_pages([image])[0]

    def run_stream(self, image, sender: Callable):
        """RettoSession::run_stream (session.rs:133-143): each stage result is sent as soon as it exists — Det before the cls forward
        runs, Cls before the rec forward, Rec at the end — through retto_b200_set_stage_callback."""
        err = []

        def on_stage(user, rp):
            try:
                r = rp.contents
                if r.n_pages < 1:
                    return
                pr = r.pages[0]
                rng = range(pr.first_line, pr.first_line + pr.n_lines)
                if r.stage == 0:
                    sender(("Det", [DetProcessorInnerResult(np.array(r.boxes[k].xy[:], np.float32).reshape(4, 2), float(np.float32(r.boxes[k].score))) for k in rng]))
                elif r.stage == 1:
                    sender(("Cls", [ClsPostProcessLabel(int(r.cls[k].label), float(np.float32(r.cls[k].score))) for k in rng]))
                else:
                    text = C.string_at(r.text, r.text_offsets[r.n_lines]) if r.n_lines else b""
                    sender(("Rec", [RecProcessorSingleResult(text[r.text_offsets[k]:r.text_offsets[k + 1]].decode("utf-8"), float(np.float32(r.rec_scores[k]))) for k in rng]))
            except Exception as e:  # noqa
                err.append(e)

        cb = STAGE_FN(on_stage)
        self.ctx._check(self.ctx._L.retto_b200_set_stage_callback(self.ctx.handle, cb, None))
        try:
            self.run(image)
        finally:
            self.ctx._L.retto_b200_set_stage_callback(self.ctx.handle, C.cast(None, STAGE_FN), None)
        if err:
            raise err[0]


def run_pages_multi(sessions: Sequence[RettoSession], images: Sequence, chunk_pages: int = 0) -> List[RettoWorkerResult]:
    """One batch over several sessions — one per GPU of the box (retto_b200_run_pages_multi): pages are sharded per image, the
    sessions' host threads pull LPT-sorted chunks from a shared cursor, results come back in page order.  Every session's worker is
    called from its own thread."""
    n_ctx, n = len(sessions), len(images)
    s0 = sessions[0]
    pages, hold = s0._pages(images, False)
    ctxs = (C.c_void_p * n_ctx)(*[s.ctx.handle for s in sessions])
    handles = {int(s.ctx.handle.value): s for s in sessions}
    for s in sessions:
        s._keep, s._err = {}, None

    def fwd(user, stage, k, inputs, outputs, stream):
        return handles[int(user)]._forward(None, stage, k, inputs, outputs, stream)

    cb = FORWARD_FN(fwd)
    users = (C.c_void_p * n_ctx)(*[s.ctx.handle for s in sessions])
    res = Results()
    st = s0.ctx._L.retto_b200_run_pages_multi(ctxs, n_ctx, pages, n, int(chunk_pages), cb, users, C.byref(res))
    for s in sessions:
        if s._err is not None:
            raise s._err
    if st not in (_lib.OK, _lib.ERR_DEGENERATE_QUAD, _lib.ERR_CAPACITY) or (st != _lib.OK and res.n_pages == 0):
        s0.ctx._check(st)
    del hold
    return RettoSession._collect(res, n)
