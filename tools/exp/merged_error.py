"""Error of the tcgen05 conv against an fp64 reference for a sweep of reduction lengths (run once per B200_TC_MERGED
setting): prints max err/tol (tol = 1e-5 + 1e-4 |want|) and the mean signed error, on the parity tests' input distribution."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from onnx_rusty_inference_engine_b200 import _lib as L
ctx = L.Context(0)
for C in (16, 32, 48, 64):
    rng = np.random.default_rng(C)
    x = (rng.standard_normal((2, C, 27, 27)) * 3).astype(np.float32)
    w = (rng.uniform(-1, 1, (128, C, 3, 3)) / np.sqrt(C * 9)).astype(np.float32)
    for relu_in in (0, 1):
        xx = np.maximum(x, 0) if relu_in else x
        want = torch.nn.functional.conv2d(torch.from_numpy(xx).double(), torch.from_numpy(w).double(), padding=1).numpy()
        got = L.conv2d(ctx, ctx.tensor(xx), ctx.tensor(w), strides=(1, 1), pads=(1,) * 4).numpy().astype(np.float64)
        err = got - want
        tol = 1e-5 + 1e-4 * np.abs(want)
        print(f"K={C*9:4d} relu_in={relu_in} max err/tol {np.abs(err / tol).max():.3f}  mean err {err.mean():+.3e}  rms err {np.sqrt((err**2).mean()):.3e}  rms want {np.sqrt((want**2).mean()):.2f}")
