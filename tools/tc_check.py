"""Diagnostic (GPU): tcgen05 3xTF32 conv vs fp64 torch reference, per shape, with error statistics.
Run in a fresh process per path: B200_CONV_PATH=1 selects the CUDA-core kernel for comparison."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from onnx_rusty_inference_engine_b200 import _lib as L

ctx = L.Context(0)
CASES = [  # N, C, H, W, M, k, stride, pad
    (1, 32, 8, 16, 16, 1, 1, 0),      # exactly one tile, one k-block
    (2, 96, 9, 11, 16, 1, 1, 0),      # squeeze-like, 3 k-blocks, ragged P
    (2, 16, 13, 13, 64, 3, 1, 1),     # expand3x3, K=144 (4.5 k-blocks), taps inside a k-block
    (1, 64, 13, 13, 1000, 1, 1, 0),   # conv10-like: 4 n-tiles of 256, ragged M
    (2, 4, 37, 41, 96, 7, 2, 0),      # conv1-like (C padded to 4): K=196
    (3, 48, 27, 27, 192, 3, 1, 1),    # fire6 expand3x3: C=48 (tap boundary inside chunks of a k-block)
    (2, 8, 14, 14, 16, 5, 1, 2),      # MNIST conv2
    (4, 512, 13, 13, 64, 1, 1, 0),    # fire9 squeeze: 16 k-blocks -> pipeline wraps
    (8, 128, 54, 54, 32, 1, 1, 0),    # many tiles
]
worst = 0.0
for (N, C, H, W, M, k, s, p) in CASES:
    rng = np.random.default_rng(C * 7 + M)
    x = (rng.standard_normal((N, C, H, W)) * 3).astype(np.float32)
    w = (rng.uniform(-1, 1, (M, C, k, k)) / np.sqrt(C * k * k)).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, (M,)).astype(np.float32)
    want = F.conv2d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), torch.from_numpy(b).double(), stride=s, padding=p).numpy()
    tx, tw, tb = ctx.tensor(x), ctx.tensor(w), ctx.tensor(b)
    t0 = time.time()
    y = L.conv2d(ctx, tx, tw, bias=tb, strides=(s, s), pads=(p, p, p, p)).numpy()
    dt = time.time() - t0
    err = np.abs(y - want)
    tol = 1e-5 + 1e-4 * np.abs(want)
    ratio = float((err / tol).max())
    worst = max(worst, ratio)
    print(f"case N={N} C={C} {H}x{W} M={M} k={k} s={s} p={p}: max|err|={err.max():.3e} max err/tol={ratio:.4f} "
          f"mean|y|={np.abs(want).mean():.3f} nan={int(np.isnan(y).sum())} ({dt*1e3:.1f} ms)", flush=True)
    if ratio > 1:
        bad = np.argwhere(err > tol)
        print("   first bad idx:", bad[:5].tolist(), " n_bad", len(bad), "of", err.size)
        i = tuple(bad[0]); print("   got", y[i], "want", want[i])
print("WORST err/tol:", worst)
sys.exit(0 if worst <= 1 else 1)
