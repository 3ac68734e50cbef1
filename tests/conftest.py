import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
MNIST_ONNX = os.path.join(GOLDEN, "mnist-8.onnx")
SYNTH_ONNX = os.path.join(ROOT, "models", "squeezenet1.0-8-synth.onnx")

# north_star tolerance: 1e-4 relative / 1e-5 absolute per output element, identical argmax
RTOL, ATOL = 1e-4, 1e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def assert_close(got, want, what=""):
    got = np.asarray(got); want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    tol = ATOL + RTOL * np.abs(want)
    err = np.abs(got - want)
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.size} elements out of tolerance; "
                           f"max err/tol {float((err / tol).max()):.3f} at {np.unravel_index(np.argmax(err / tol), err.shape)}")


@pytest.fixture(scope="session")
def ctx():
    from onnx_rusty_inference_engine_b200 import _lib
    return _lib.Context(0)


@pytest.fixture(scope="session")
def synth_onnx():
    from onnx_rusty_inference_engine_b200 import synth
    return synth.ensure_squeezenet(SYNTH_ONNX, seed=0)
