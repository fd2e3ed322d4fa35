"""ONNX Runtime worker for the forward seam: RettoOrtWorker (retto-core/src/worker/ort_worker.rs:188-221) with device-resident inputs and
outputs bound through IoBinding — the tensors are the context's own buffers (`retto_b200_forward_fn`), so no page touches host memory
between stages (the reference copies six times per stage).

UNTESTED IN THIS IMAGE: onnxruntime and the PP-OCRv4 .onnx files are not available here, so this module has never
executed; the seam it plugs into is exercised by retto_b200.standin (torch networks of the same I/O shape, bench.py `with_forward`,
`python -m retto_b200.cli --worker standin`).  It is kept as the integration sketch INTEGRATION.md §2 refers to, not counted as a feature."""
from __future__ import annotations

import numpy as np


class OrtIoBindingWorker:
    def __init__(self, device_id: int, det: str, cls: str, rec: str):
        try:
            import onnxruntime as ort
        except ImportError as e:
            raise SystemExit("retto_b200: onnxruntime is not installed; pass --worker module:factory (or --worker standin) for the forward passes") from e
        import torch
        self.torch, self.device_id = torch, device_id
        prov = [("CUDAExecutionProvider", {"device_id": device_id})]
        self.sess = [ort.InferenceSession(p, providers=prov) for p in (det, cls, rec)]

    def _run(self, k: int, xs):
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
