"""ctypes binding of libretto_b200.so (include/retto_b200.h).

There is no CPU fallback: if the shared library is missing or does not load, importing any compute
entry point raises.  `build()` compiles it in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libretto_b200.so")
CSRC = os.path.join(_HERE, "csrc")


class RettoB200Error(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"retto_b200 status {status} ({STATUS_NAMES.get(status, '?')}): {msg}")
        self.status = status


STATUS_NAMES = {
    0: "OK", 1: "ERR_INVALID_ARG", 2: "ERR_CUDA", 3: "ERR_OOM", 4: "ERR_CAPACITY", 5: "ERR_NAN_LOGITS",
    6: "ERR_DEGENERATE_QUAD", 7: "ERR_NO_DICT", 8: "ERR_WORKER", 9: "ERR_UNSUPPORTED", 10: "ERR_DECODE",
}
OK, ERR_INVALID_ARG, ERR_CUDA, ERR_OOM, ERR_CAPACITY, ERR_NAN_LOGITS, ERR_DEGENERATE_QUAD, ERR_NO_DICT, ERR_WORKER, ERR_UNSUPPORTED, ERR_DECODE = range(11)
ABI_VERSION = 2
PAGE_HOST_RGB, PAGE_DEVICE_RGB, PAGE_HOST_ENCODED = 0, 1, 2


class Config(C.Structure):
    _fields_ = [
        ("max_side_len", C.c_int32), ("min_side_len", C.c_int32),
        ("det_limit_side_len", C.c_int32), ("det_limit_type", C.c_int32),
        ("det_mean", C.c_float * 3), ("det_std", C.c_float * 3), ("det_scale", C.c_float),
        ("det_thresh", C.c_float), ("det_box_thresh", C.c_float), ("det_max_candidates", C.c_int32),
        ("det_unclip_ratio", C.c_float), ("det_use_dilation", C.c_int32), ("det_score_mode", C.c_int32),
        ("det_min_mini_box_size", C.c_int32), ("det_dilation_2x2", C.c_int32),
        ("cls_image_shape", C.c_int32 * 3), ("cls_batch_num", C.c_int32), ("cls_thresh", C.c_float), ("cls_label", C.c_int32 * 2),
        ("rec_image_shape", C.c_int32 * 3), ("rec_batch_num", C.c_int32),
        ("max_components_per_page", C.c_int32), ("max_det_side", C.c_int32),
    ]


class ResizeDesc(C.Structure):
    _fields_ = [("d_src", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("d_dst", C.c_void_p), ("out_h", C.c_int32), ("out_w", C.c_int32)]


class DetPreDesc(C.Structure):
    _fields_ = [("d_rgb", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("d_out", C.c_void_p), ("out_h", C.c_int32), ("out_w", C.c_int32)]


class DetPostDesc(C.Structure):
    _fields_ = [("d_prob", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("ori_h", C.c_int32), ("ori_w", C.c_int32)]


class Box(C.Structure):
    _fields_ = [("xy", C.c_float * 8), ("score", C.c_float)]


class CropJob(C.Structure):
    _fields_ = [("d_page", C.c_void_p), ("page_h", C.c_int32), ("page_w", C.c_int32), ("box", Box)]


class CropInfo(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("rotated270", C.c_int32), ("status", C.c_int32), ("offset", C.c_uint64)]


class Batch(C.Structure):
    _fields_ = [("first_line", C.c_int32), ("n", C.c_int32), ("img_w", C.c_int32), ("max_wh_ratio", C.c_float), ("offset", C.c_uint64)]


class LineJob(C.Structure):
    _fields_ = [("crop", C.c_int32), ("img_w", C.c_int32), ("resized_w", C.c_int32), ("dst_offset", C.c_uint64)]


class ClsResult(C.Structure):
    _fields_ = [("label", C.c_int32), ("score", C.c_float)]


class LogitsDesc(C.Structure):
    _fields_ = [("d_logits", C.c_void_p), ("n", C.c_int32), ("t", C.c_int32)]


class Tensor(C.Structure):
    _fields_ = [("d_data", C.c_void_p), ("shape", C.c_int64 * 4), ("ndim", C.c_int32)]


class Page(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("on_device", C.c_int32), ("n_bytes", C.c_uint64)]


class Encoded(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("n_bytes", C.c_uint64)]


class ImageInfo(C.Structure):
    _fields_ = [("h", C.c_int32), ("w", C.c_int32), ("format", C.c_int32), ("components", C.c_int32), ("subsampling", C.c_int32),
                ("restart_interval", C.c_int32), ("status", C.c_int32)]


class PageResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("first_line", C.c_int32), ("n_lines", C.c_int32)]


class Results(C.Structure):
    _fields_ = [
        ("n_pages", C.c_int32), ("pages", C.POINTER(PageResult)), ("n_lines", C.c_int32), ("boxes", C.POINTER(Box)),
        ("cls", C.POINTER(ClsResult)), ("text_offsets", C.POINTER(C.c_uint32)), ("text", C.POINTER(C.c_char)),
        ("rec_scores", C.POINTER(C.c_float)),
    ]


class StageResult(C.Structure):
    _fields_ = [
        ("stage", C.c_int32), ("first_page", C.c_int32), ("n_pages", C.c_int32), ("pages", C.POINTER(PageResult)), ("n_lines", C.c_int32),
        ("boxes", C.POINTER(Box)), ("cls", C.POINTER(ClsResult)), ("text_offsets", C.POINTER(C.c_uint32)), ("text", C.POINTER(C.c_char)),
        ("rec_scores", C.POINTER(C.c_float)),
    ]


STAGE_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(StageResult))
FORWARD_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p)

# every symbol include/retto_b200.h declares (tests/test_abi.py checks the .so exports all of them)
EXPORTS = [
    "retto_b200_config_default", "retto_b200_abi_version", "retto_b200_create", "retto_b200_destroy", "retto_b200_last_error",
    "retto_b200_stream", "retto_b200_sync", "retto_b200_launch_count", "retto_b200_enable_kernel_timing",
    "retto_b200_reset_kernel_times", "retto_b200_kernel_times", "retto_b200_dev_alloc", "retto_b200_dev_free",
    "retto_b200_host_alloc", "retto_b200_host_free", "retto_b200_h2d", "retto_b200_d2h", "retto_b200_resize_both_plan",
    "retto_b200_resize_either_plan", "retto_b200_thumbnail", "retto_b200_det_preprocess", "retto_b200_det_postprocess",
    "retto_b200_det_post_fetch_bitmap", "retto_b200_det_post_fetch_labels", "retto_b200_det_post_enable_trace",
    "retto_b200_det_post_fetch_trace", "retto_b200_scale_and_clip", "retto_b200_crop_boxes",
    "retto_b200_crop_fetch", "retto_b200_plan_batches", "retto_b200_build_batches", "retto_b200_cls_postprocess",
    "retto_b200_dict_load", "retto_b200_dict_size", "retto_b200_ctc_decode", "retto_b200_ctc_argmax", "retto_b200_run_pages", "retto_b200_last_run_stats",
    "retto_b200_set_pipeline", "retto_b200_set_stage_callback", "retto_b200_run_pages_multi", "retto_b200_image_info", "retto_b200_decode_images",
]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libretto_b200.so in-tree (nvcc, -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    cmd = ["make", "-C", CSRC, "-j", "8"] + (["-B"] if force else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libretto_b200.so failed")
    return SO_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA extension.  Raises if it is missing: the product has no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    L = C.CDLL(SO_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
    L.retto_b200_abi_version.restype = i32
    L.retto_b200_config_default.argtypes = [C.POINTER(Config)]
    L.retto_b200_create.argtypes = [i32, C.POINTER(Config), C.POINTER(vp)]
    L.retto_b200_destroy.argtypes = [vp]
    L.retto_b200_last_error.argtypes = [vp]
    L.retto_b200_last_error.restype = C.c_char_p
    L.retto_b200_stream.argtypes = [vp]
    L.retto_b200_stream.restype = vp
    L.retto_b200_sync.argtypes = [vp]
    L.retto_b200_launch_count.argtypes = [vp]
    L.retto_b200_launch_count.restype = u64
    L.retto_b200_enable_kernel_timing.argtypes = [vp, i32]
    L.retto_b200_reset_kernel_times.argtypes = [vp]
    L.retto_b200_kernel_times.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.retto_b200_dev_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.retto_b200_dev_free.argtypes = [vp, vp]
    L.retto_b200_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.retto_b200_host_free.argtypes = [vp, vp]
    L.retto_b200_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.retto_b200_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.retto_b200_resize_both_plan.argtypes = [i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.retto_b200_resize_either_plan.argtypes = [i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.retto_b200_thumbnail.argtypes = [vp, C.POINTER(ResizeDesc), i32]
    L.retto_b200_det_preprocess.argtypes = [vp, C.POINTER(DetPreDesc), i32]
    L.retto_b200_det_postprocess.argtypes = [vp, C.POINTER(DetPostDesc), i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(Box), i32]
    L.retto_b200_det_post_fetch_bitmap.argtypes = [vp, i32, vp]
    L.retto_b200_det_post_fetch_labels.argtypes = [vp, i32, vp]
    L.retto_b200_det_post_enable_trace.argtypes = [vp, i32]
    L.retto_b200_det_post_fetch_trace.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), vp, vp, vp, vp, vp, i32]
    L.retto_b200_scale_and_clip.argtypes = [vp, C.POINTER(Box), i32, C.c_double, C.c_double, C.c_double, C.c_double]
    L.retto_b200_crop_boxes.argtypes = [vp, C.POINTER(CropJob), i32, C.POINTER(CropInfo)]
    L.retto_b200_crop_fetch.argtypes = [vp, i32, vp]
    L.retto_b200_plan_batches.argtypes = [C.POINTER(Config), i32, C.POINTER(CropInfo), i32, C.POINTER(LineJob), C.POINTER(Batch), C.POINTER(i32), C.POINTER(u64)]
    L.retto_b200_build_batches.argtypes = [vp, i32, C.POINTER(LineJob), i32, u64, C.POINTER(vp)]
    L.retto_b200_cls_postprocess.argtypes = [vp, vp, i32, C.POINTER(i32), C.POINTER(ClsResult)]
    L.retto_b200_dict_load.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.retto_b200_dict_size.argtypes = [vp]
    L.retto_b200_dict_size.restype = i32
    L.retto_b200_ctc_decode.argtypes = [vp, C.POINTER(LogitsDesc), i32, i32, C.POINTER(C.c_uint32), vp, C.c_size_t, C.POINTER(C.c_float), C.POINTER(i32), C.POINTER(i32), i32]
    L.retto_b200_ctc_argmax.argtypes = [vp, C.POINTER(LogitsDesc), i32, i32, vp, vp]
    L.retto_b200_last_run_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.retto_b200_set_pipeline.argtypes = [vp, i32, i32]
    L.retto_b200_run_pages.argtypes = [vp, C.POINTER(Page), i32, FORWARD_FN, vp, C.POINTER(Results)]
    L.retto_b200_set_stage_callback.argtypes = [vp, STAGE_FN, vp]
    L.retto_b200_run_pages_multi.argtypes = [C.POINTER(vp), i32, C.POINTER(Page), i32, i32, FORWARD_FN, C.POINTER(vp), C.POINTER(Results)]
    L.retto_b200_image_info.argtypes = [vp, u64, C.POINTER(ImageInfo)]
    L.retto_b200_decode_images.argtypes = [vp, C.POINTER(Encoded), i32, C.POINTER(vp), C.POINTER(i32)]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("retto_b200_abi_version", "retto_b200_dict_size"):
            fn.restype = i32
    _lib = L
    return L
