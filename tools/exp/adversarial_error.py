"""Experiment (GPU): accuracy of the tcgen05 conv on same-sign inputs.  Run once per setting in a fresh process:
    B200_TC_MERGED=0|1 (force the accumulator mode) or B200_CONV_PATH=1 (CUDA-core fp32 kernel) python tools/exp/adversarial_error.py
Prints max err/tol (tol = 1e-5 + 1e-4 |want|, want = fp64 conv) per distribution and reduction length, 3x3 / pad 1, M = 128."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from onnx_rusty_inference_engine_b200 import _lib as L
ctx = L.Context(0)
tag = f"merged={os.environ.get('B200_TC_MERGED', 'auto')} conv_path={os.environ.get('B200_CONV_PATH', '0')}"
for C in (16, 32, 48, 64):
    rng = np.random.default_rng(C)
    base = rng.standard_normal((3, C, 27, 27))
    w = (rng.uniform(-1, 1, (128, C, 3, 3)) / np.sqrt(C * 9)).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, (128,)).astype(np.float32)
    dists = {"normal3": base * 3, "relu_pos": np.maximum(base * 3 + 2.0, 0), "offset_pos": rng.uniform(5.0, 15.0, base.shape),
             "relu_n10": np.maximum(base * 10, 0)}
    for name, x in dists.items():
        x = x.astype(np.float32)
        want = torch.nn.functional.conv2d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), torch.from_numpy(b).double(), padding=1).numpy()
        got = L.conv2d(ctx, ctx.tensor(x), ctx.tensor(w), bias=ctx.tensor(b), strides=(1, 1), pads=(1,) * 4).numpy().astype(np.float64)
        err = got - want
        tol = 1e-5 + 1e-4 * np.abs(want)
        print(f"{tag} K={C*9:4d} {name:10s} max err/tol {np.abs(err / tol).max():.3f}  frac>0.5 {float((np.abs(err/tol) > 0.5).mean()):.2e}  rms err {np.sqrt((err**2).mean()):.3e}  rms want {np.sqrt((want**2).mean()):.2f}")
