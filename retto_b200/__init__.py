"""retto_b200 — B200-native (sm_100a) implementation of retto-core's OCR image path.

Public surface: `retto_b200.api.Context` (stage-level calls over the C ABI) and the reference-shaped
host mirror in `retto_b200.session` (RettoSession / processors).  The CUDA extension
(libretto_b200.so) is mandatory: there is no CPU fallback.
"""
from ._lib import RettoB200Error, build  # noqa: F401

__all__ = ["RettoB200Error", "build"]
