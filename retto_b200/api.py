"""Thin Python surface over the C ABI: a `Context` object whose methods take torch CUDA tensors (torch
is used only for device memory and streams) and call libretto_b200.so.  All pixel/logit work runs in
the library's CUDA kernels; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (Batch, Box, ClsResult, Config, CropInfo, CropJob, DetPostDesc, DetPreDesc, LineJob, LogitsDesc, ResizeDesc,
                   RettoB200Error)


def default_config() -> Config:
    cfg = Config()
    _lib.lib().retto_b200_config_default(C.byref(cfg))
    return cfg


def resize_both_plan(h: int, w: int, max_side_len: int = 2000, min_side_len: int = 30):
    dims = (C.c_int32 * 4)()
    n = C.c_int32()
    st = _lib.lib().retto_b200_resize_both_plan(h, w, max_side_len, min_side_len, dims, C.byref(n))
    if st:
        raise RettoB200Error(st, "resize_both_plan")
    return [(dims[2 * i], dims[2 * i + 1]) for i in range(n.value)]


def resize_either_plan(h: int, w: int, limit_type: int = 0, limit_len: int = 736):
    oh, ow = C.c_int32(), C.c_int32()
    st = _lib.lib().retto_b200_resize_either_plan(h, w, limit_type, limit_len, C.byref(oh), C.byref(ow))
    if st:
        raise RettoB200Error(st, "resize_either_plan")
    return oh.value, ow.value


def image_info(data: bytes) -> _lib.ImageInfo:
    """header parse of an image file (host only): dims, sub-sampling, restart interval, status"""
    buf = np.frombuffer(data, np.uint8)
    info = _lib.ImageInfo()
    _lib.lib().retto_b200_image_info(buf.ctypes.data, len(buf), C.byref(info))
    return info


@dataclass
class DetPostOut:
    page_status: np.ndarray   # [n] int32
    offsets: np.ndarray       # [n+1] int32
    boxes: np.ndarray         # [total,4,2] f32
    scores: np.ndarray        # [total] f32

    def page(self, p: int):
        a, b = int(self.offsets[p]), int(self.offsets[p + 1])
        return self.boxes[a:b], self.scores[a:b]


class Context:
    """One GPU + one stream (not thread-safe, like `RettoSession::run(&mut self)`, session.rs:108)."""

    def __init__(self, device_id: int = 0, cfg: Optional[Config] = None):
        self._L = _lib.lib()
        self.cfg = cfg if cfg is not None else default_config()
        h = C.c_void_p()
        st = self._L.retto_b200_create(device_id, C.byref(self.cfg), C.byref(h))
        if st:
            raise RettoB200Error(st, "retto_b200_create failed (is a CUDA device visible?)")
        self._h = h
        self.device_id = device_id
        self._torch_stream = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.retto_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int):
        if st:
            raise RettoB200Error(st, self._L.retto_b200_last_error(self._h).decode("utf-8", "replace"))

    @property
    def handle(self):
        return self._h

    @property
    def stream_ptr(self) -> int:
        return int(self._L.retto_b200_stream(self._h) or 0)

    def torch_stream(self):
        """The context's CUDA stream as a torch ExternalStream (so torch work can be ordered with ours)."""
        import torch
        if self._torch_stream is None:
            self._torch_stream = torch.cuda.ExternalStream(self.stream_ptr, device=f"cuda:{self.device_id}")
        return self._torch_stream

    def sync(self):
        self._check(self._L.retto_b200_sync(self._h))

    def _ordered(self):
        """Stream ordering for calls that read / write torch tensors: the context enqueues on its OWN non-blocking stream, torch
        allocates and fills tensors on its current stream.  Entering makes the context's stream wait for everything queued on
        torch's current stream (the inputs are ready when our kernels run); leaving makes torch's current stream wait for the
        context's (the outputs are ready for whatever torch does next) — no host synchronisation either way."""
        import contextlib
        import torch

        @contextlib.contextmanager
        def cm():
            with torch.cuda.device(self.device_id):
                cur, mine = torch.cuda.current_stream(), self.torch_stream()
                mine.wait_stream(cur)
                try:
                    yield
                finally:
                    cur.wait_stream(mine)
        return cm()

    def set_pipeline(self, lanes: int = 0, unit_pages: int = 0):
        """run_pages pipeline: lanes 1 or 2 (0 = default 1), pages per unit (0 = default)"""
        self._check(self._L.retto_b200_set_pipeline(self._h, int(lanes), int(unit_pages)))

    @property
    def launch_count(self) -> int:
        return int(self._L.retto_b200_launch_count(self._h))

    def enable_kernel_timing(self, on: bool = True):
        self._check(self._L.retto_b200_enable_kernel_timing(self._h, int(on)))

    def reset_kernel_times(self):
        self._check(self._L.retto_b200_reset_kernel_times(self._h))

    def kernel_times(self):
        """{kernel name: (launch count, total ms)} measured with CUDA events on the context's stream"""
        buf = C.create_string_buffer(1 << 16)
        self._check(self._L.retto_b200_kernel_times(self._h, buf, len(buf)))
        out = {}
        for ln in buf.value.decode().splitlines():
            name, cnt, ms = ln.split("\t")
            out[name] = (int(cnt), float(ms))
        return out

    # ---- image decode ---------------------------------------------------------------------------
    def decode_images(self, files: Sequence[bytes]):
        """ImageHelper::new_from_raw_img_flow (image_helper.rs:34-44) for baseline JPEG files on the device.
        Returns (list of torch uint8 CUDA tensors [h,w,3] or None, list of statuses)."""
        import torch
        n = len(files)
        bufs = [np.frombuffer(f, np.uint8) for f in files]
        enc = (_lib.Encoded * max(n, 1))()
        outs, ptrs = [], (C.c_void_p * max(n, 1))()
        for i, b in enumerate(bufs):
            enc[i] = _lib.Encoded(b.ctypes.data, len(b))
            info = image_info(files[i])
            if info.status == 0:
                t = torch.empty((info.h, info.w, 3), dtype=torch.uint8, device=f"cuda:{self.device_id}")
                outs.append(t)
                ptrs[i] = t.data_ptr()
            else:
                outs.append(None)
                ptrs[i] = None
        status = (C.c_int32 * max(n, 1))()
        with self._ordered():
            self._L.retto_b200_decode_images(self._h, enc, n, ptrs, status)
        return outs, [int(status[i]) for i in range(n)]

    # ---- resizes ------------------------------------------------------------------------------
    def thumbnail(self, srcs: Sequence, out_dims: Sequence):
        """image::imageops::thumbnail on a batch of HWC u8 CUDA tensors."""
        import torch
        outs, descs = [], (ResizeDesc * len(srcs))()
        for i, (s, (oh, ow)) in enumerate(zip(srcs, out_dims)):
            assert s.dtype == torch.uint8 and s.is_cuda and s.is_contiguous() and s.shape[2] == 3
            o = torch.empty((oh, ow, 3), dtype=torch.uint8, device=s.device)
            outs.append(o)
            descs[i] = ResizeDesc(s.data_ptr(), s.shape[0], s.shape[1], o.data_ptr(), oh, ow)
        with self._ordered():
            self._check(self._L.retto_b200_thumbnail(self._h, descs, len(srcs)))
        return outs

    # ---- det preprocess -----------------------------------------------------------------------
    def det_preprocess(self, pages: Sequence):
        """DetProcessor::preprocess for a batch of HWC u8 CUDA pages -> list of [1,3,H',W'] f32 tensors."""
        import torch
        outs, descs = [], (DetPreDesc * len(pages))()
        for i, p in enumerate(pages):
            assert p.dtype == torch.uint8 and p.is_cuda and p.is_contiguous() and p.shape[2] == 3
            oh, ow = resize_either_plan(p.shape[0], p.shape[1], self.cfg.det_limit_type, self.cfg.det_limit_side_len)
            o = torch.empty((1, 3, oh, ow), dtype=torch.float32, device=p.device)
            outs.append(o)
            descs[i] = DetPreDesc(p.data_ptr(), p.shape[0], p.shape[1], o.data_ptr(), oh, ow)
        with self._ordered():
            self._check(self._L.retto_b200_det_preprocess(self._h, descs, len(pages)))
        return outs

    # ---- det postprocess ----------------------------------------------------------------------
    def det_postprocess(self, probs: Sequence, ori_hw: Sequence, max_boxes_total: int = 0) -> DetPostOut:
        """DetProcessor::postprocess for a batch of [H,W] (or [1,1,H,W]) f32 CUDA probability maps."""
        import torch
        n = len(probs)
        descs = (DetPostDesc * max(n, 1))()
        for i, (p, (oh, ow)) in enumerate(zip(probs, ori_hw)):
            assert p.dtype == torch.float32 and p.is_cuda and p.is_contiguous()
            descs[i] = DetPostDesc(p.data_ptr(), p.shape[-2], p.shape[-1], oh, ow)
        cap = max_boxes_total or max(4096, 1024 * n)
        status = np.zeros(max(n, 1), np.int32)
        offs = np.zeros(n + 1, np.int32)
        boxes = (Box * cap)()
        with self._ordered():
            st = self._L.retto_b200_det_postprocess(self._h, descs, n, status.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    offs.ctypes.data_as(C.POINTER(C.c_int32)), boxes, cap)
        self._check(st)
        tot = int(offs[n])
        arr = np.frombuffer(boxes, dtype=np.float32, count=tot * 9).reshape(tot, 9) if tot else np.zeros((0, 9), np.float32)
        return DetPostOut(status[:n].copy(), offs, arr[:, :8].reshape(tot, 4, 2).copy(), arr[:, 8].copy())

    def fetch_bitmap(self, page: int, h: int, w: int) -> np.ndarray:
        out = np.zeros((h, w), np.uint8)
        self._check(self._L.retto_b200_det_post_fetch_bitmap(self._h, page, out.ctypes.data))
        return out

    def fetch_labels(self, page: int, h: int, w: int) -> np.ndarray:
        out = np.zeros((h, w), np.int32)
        self._check(self._L.retto_b200_det_post_fetch_labels(self._h, page, out.ctypes.data))
        return out

    def enable_trace(self, on: bool = True):
        self._check(self._L.retto_b200_det_post_enable_trace(self._h, int(on)))

    def fetch_trace(self, page: int):
        """per-component parity tap: dict(key, status, rect1[n,8], sside1, score, n_holes)"""
        n, nh = C.c_int32(), C.c_int32()
        self._check(self._L.retto_b200_det_post_fetch_trace(self._h, page, C.byref(n), C.byref(nh), None, None, None, None, None, 0))
        k = max(n.value, 1)
        key, st = np.zeros(k, np.int32), np.zeros(k, np.int32)
        rect, ss, sc = np.zeros((k, 8), np.int32), np.zeros(k, np.float32), np.zeros(k, np.float32)
        self._check(self._L.retto_b200_det_post_fetch_trace(self._h, page, C.byref(n), C.byref(nh), key.ctypes.data, st.ctypes.data,
                                                            rect.ctypes.data, ss.ctypes.data, sc.ctypes.data, k))
        m = n.value
        return dict(key=key[:m], status=st[:m], rect1=rect[:m], sside1=ss[:m], score=sc[:m], n_holes=nh.value)

    def scale_and_clip(self, boxes: np.ndarray, bitmap_w, bitmap_h, ori_w, ori_h) -> np.ndarray:
        n = len(boxes)
        arr = (Box * max(n, 1))()
        for i in range(n):
            arr[i].xy[:] = [float(v) for v in np.asarray(boxes[i]).reshape(8)]
        self._check(self._L.retto_b200_scale_and_clip(self._h, arr, n, bitmap_w, bitmap_h, ori_w, ori_h))
        return np.array([[arr[i].xy[k] for k in range(8)] for i in range(n)], np.float32).reshape(n, 4, 2)

    # ---- crops ----------------------------------------------------------------------------------
    def crop_boxes(self, pages: Sequence, page_of_box: Sequence[int], boxes: np.ndarray) -> List[CropInfo]:
        """ImageHelper::get_crop_img for boxes [n,4,2] (each in the coordinates of pages[page_of_box[i]])."""
        n = len(boxes)
        jobs = (CropJob * max(n, 1))()
        for i in range(n):
            p = pages[page_of_box[i]]
            jobs[i].d_page = p.data_ptr()
            jobs[i].page_h, jobs[i].page_w = p.shape[0], p.shape[1]
            jobs[i].box.xy[:] = [float(v) for v in np.asarray(boxes[i]).reshape(8)]
        infos = (CropInfo * max(n, 1))()
        with self._ordered():
            self._check(self._L.retto_b200_crop_boxes(self._h, jobs, n, infos))
        return [infos[i] for i in range(n)]

    def crop_fetch(self, i: int, info: CropInfo) -> np.ndarray:
        out = np.zeros((info.h, info.w, 3), np.uint8)
        self._check(self._L.retto_b200_crop_fetch(self._h, i, out.ctypes.data))
        return out

    # ---- batches --------------------------------------------------------------------------------
    def plan_batches(self, kind: int, infos: Sequence[CropInfo]):
        """ordering/batching rules of Cls/RecProcessor::process for ONE page -> (lines, batches, total_floats)"""
        n = len(infos)
        arr = (CropInfo * max(n, 1))(*infos)
        lines = (LineJob * max(n, 1))()
        batches = (Batch * (n + 1))()
        nb = C.c_int32()
        tot = C.c_uint64()
        st = self._L.retto_b200_plan_batches(C.byref(self.cfg), kind, arr, n, lines, batches, C.byref(nb), C.byref(tot))
        self._check(st)
        return [lines[i] for i in range(n)], [batches[i] for i in range(nb.value)], int(tot.value)

    def build_batches(self, kind: int, lines: Sequence[LineJob], total_floats: int) -> int:
        n = len(lines)
        arr = (LineJob * max(n, 1))(*lines)
        base = C.c_void_p()
        with self._ordered():
            self._check(self._L.retto_b200_build_batches(self._h, kind, arr, n, total_floats, C.byref(base)))
        return int(base.value or 0)

    def cls_postprocess(self, logits, crop_index: Sequence[int]):
        n = len(crop_index)
        idx = (C.c_int32 * max(n, 1))(*crop_index)
        res = (ClsResult * max(n, 1))()
        with self._ordered():
            self._check(self._L.retto_b200_cls_postprocess(self._h, logits.data_ptr(), n, idx, res))
        return [(res[i].label, res[i].score) for i in range(n)]

    # ---- CTC ------------------------------------------------------------------------------------
    def dict_load(self, text: str):
        b = text.encode("utf-8")
        self._check(self._L.retto_b200_dict_load(self._h, b, len(b)))

    @property
    def dict_size(self) -> int:
        return int(self._L.retto_b200_dict_size(self._h))

    def ctc_decode(self, logits_list: Sequence, want_tokens: bool = False):
        """RecProcessor::postprocess + decode for a list of [n,T,C] f32 CUDA tensors.
        returns (texts, scores[, tokens, counts]); raises on NaN logits like the reference panics."""
        import torch
        nd = len(logits_list)
        descs = (LogitsDesc * max(nd, 1))()
        total, max_t, Cc = 0, 1, self.dict_size
        for i, t in enumerate(logits_list):
            assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous() and t.dim() == 3
            Cc = t.shape[2]
            descs[i] = LogitsDesc(t.data_ptr(), t.shape[0], t.shape[1])
            total += t.shape[0]
            max_t = max(max_t, t.shape[1])
        offs = np.zeros(total + 1, np.uint32)
        cap = max(16, total * max_t * 8)
        text = C.create_string_buffer(cap)
        scores = np.zeros(max(total, 1), np.float32)
        tokens = np.zeros((max(total, 1), max_t), np.int32) if want_tokens else None
        counts = np.zeros(max(total, 1), np.int32) if want_tokens else None
        with self._ordered():
            st = self._L.retto_b200_ctc_decode(
                self._h, descs, nd, Cc, offs.ctypes.data_as(C.POINTER(C.c_uint32)), C.cast(text, C.c_void_p), cap,
                scores.ctypes.data_as(C.POINTER(C.c_float)),
                tokens.ctypes.data_as(C.POINTER(C.c_int32)) if want_tokens else None,
                counts.ctypes.data_as(C.POINTER(C.c_int32)) if want_tokens else None, max_t)
        self._check(st)
        raw = text.raw
        texts = [raw[offs[i]:offs[i + 1]].decode("utf-8") for i in range(total)]
        if want_tokens:
            return texts, scores[:total], tokens[:total], counts[:total]
        return texts, scores[:total]

    def ctc_argmax(self, logits):
        """argmax / max over classes only (rec_processor.rs:198-199); outputs stay on the device."""
        import torch
        n, T, Cc = logits.shape
        descs = (LogitsDesc * 1)(LogitsDesc(logits.data_ptr(), n, T))
        idx = torch.empty((n, T), dtype=torch.int32, device=logits.device)
        prob = torch.empty((n, T), dtype=torch.float32, device=logits.device)
        with self._ordered():
            self._check(self._L.retto_b200_ctc_argmax(self._h, descs, 1, Cc, idx.data_ptr(), prob.data_ptr()))
        return idx, prob
