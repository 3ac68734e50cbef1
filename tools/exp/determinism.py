"""Stress (GPU): run each conv layer many times on the same input and compare the outputs bit for bit."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from onnx_rusty_inference_engine_b200 import _lib as L

LAYERS = {  # name: (N, C, H, W, M, k, stride, pad)
    "conv1": (64, 4, 224, 224, 96, 7, 2, 0),
    "f2_sq": (256, 96, 54, 54, 16, 1, 1, 0),
    "f2_fused": (128, 16, 54, 54, 128, 3, 1, 1),
    "f4_e1": (256, 32, 54, 54, 128, 1, 1, 0),
    "f4_e3": (128, 32, 54, 54, 128, 3, 1, 1),
    "f6_e1": (256, 48, 27, 27, 192, 1, 1, 0),
    "f6_e3": (256, 48, 27, 27, 192, 3, 1, 1),
    "f8_sq": (256, 384, 27, 27, 64, 1, 1, 0),
    "f8_e1": (256, 64, 27, 27, 256, 1, 1, 0),
    "f8_e3": (256, 64, 27, 27, 256, 3, 1, 1),
    "f9_sq": (256, 512, 13, 13, 64, 1, 1, 0),
    "conv10": (256, 512, 13, 13, 1000, 1, 1, 0),
}
names = sys.argv[1:] or list(LAYERS)
reps = int(os.environ.get("REPS", "25"))
torch.cuda.set_device(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
ctx = L.Context(0, s.cuda_stream)
bad_total = 0
for nm in names:
    N, C, H, W, M, k, st, p = LAYERS[nm]
    rng = np.random.default_rng(1)
    x = ctx.tensor((rng.standard_normal((N, C, H, W)) * 3).astype(np.float32))
    w = ctx.tensor((rng.standard_normal((M, C, k, k)) * 0.05).astype(np.float32))
    b = ctx.tensor(rng.standard_normal((M,)).astype(np.float32))
    y = L.conv2d(ctx, x, w, bias=b, strides=(st, st), pads=(p,) * 4, fuse_relu=True)
    ref = y.numpy().copy()
    bad = 0
    for it in range(reps):
        L.conv2d(ctx, x, w, bias=b, strides=(st, st), pads=(p,) * 4, fuse_relu=True, y=y)
        got = y.numpy()
        if not np.array_equal(got, ref):
            d = np.argwhere(got != ref)
            bad += 1
            if bad <= 6:
                Ho, Wo = got.shape[2], got.shape[3]
                pix = d[:, 0] * Ho * Wo + d[:, 2] * Wo + d[:, 3]
                tiles = sorted(set((pix // 128).tolist())); rows = sorted(set((pix % 128).tolist())); chans = sorted(set(d[:, 1].tolist()))
                err = np.abs(got - ref)[tuple(d.T)]
                print(f"  {nm} rep {it}: {len(d)} elements differ; tiles {tiles[:8]} rows {rows[:40]} channels {chans[:40]} max|diff| {err.max():.3g} ref~{np.abs(ref[tuple(d.T)]).mean():.3g}", flush=True)
    print(f"{nm:10s} {'OK' if bad == 0 else 'NONDETERMINISTIC'} ({bad}/{reps} runs differ)", flush=True)
    bad_total += bad
sys.exit(1 if bad_total else 0)
