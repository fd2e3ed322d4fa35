import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from retto_b200.api import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def synth_dict():
    from tools.synth import synth_dict_text
    return synth_dict_text()


# ---- libm policy -----------------------------------------------------------------------------------------
# Every ASSERTED comparison runs with the oracle on the host glibc (what the reference's f64 trig calls; the
# oracle then shares no code with the kernels).  Tests may add a DIAGNOSTIC second run with the CUDA path's
# correctly-rounded trig hooked in (oracle.set_libm(1)) and report how many boxes differ between the two
# through `libm_diag`; the tally is printed at the end of the session and written to gpurun_out/.
_LIBM_TALLY = {"boxes_compared": 0, "boxes_differ_glibc_vs_crmath": 0, "corpora": {}}


@pytest.fixture(autouse=True)
def _glibc_oracle():
    from oracle import oracle as O
    O.set_libm(0)
    yield
    O.set_libm(0)


@pytest.fixture
def libm_diag():
    """libm_diag(name, fn, asserted) -> re-runs fn() with the diagnostic trig hooks and counts the boxes that
    differ from `asserted` (a list of per-page box arrays obtained in glibc mode)."""
    from oracle import oracle as O
    import numpy as np

    def run(name, fn, asserted):
        O.set_libm(1)
        try:
            other = fn()
        finally:
            O.set_libm(0)
        n = d = 0
        for a, b in zip(asserted, other):
            a, b = np.asarray(a), np.asarray(b)
            if a.shape != b.shape:
                n += max(len(a), len(b))
                d += max(len(a), len(b))
                continue
            n += len(a)
            d += int((a.reshape(len(a), -1) != b.reshape(len(b), -1)).any(axis=1).sum()) if len(a) else 0
        _LIBM_TALLY["boxes_compared"] += n
        _LIBM_TALLY["boxes_differ_glibc_vs_crmath"] += d
        c = _LIBM_TALLY["corpora"].setdefault(name, [0, 0])
        c[0] += n
        c[1] += d
        return d

    return run


def pytest_terminal_summary(terminalreporter):
    if _LIBM_TALLY["boxes_compared"]:
        import json
        terminalreporter.write_line("libm diagnostic (oracle glibc vs oracle with the CUDA path's CR trig): " + json.dumps(_LIBM_TALLY))
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "libm_diag.json"), "w") as f:
                json.dump(_LIBM_TALLY, f)
        except OSError:
            pass
