"""Micro-benchmark (GPU): one conv layer through the per-op C ABI, CUDA-event timed (torch events on a shared stream).
usage: tc_bench.py [name ...]   names from LAYERS; default all.  B200_CONV_PATH=1 -> CUDA-core kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_rusty_inference_engine_b200 import _lib as L

LAYERS = {  # name: (N, C, H, W, M, k, stride, pad)
    "conv1": (256, 4, 224, 224, 96, 7, 2, 0),
    "f2_sq": (256, 96, 54, 54, 16, 1, 1, 0),
    "f2_e1": (256, 16, 54, 54, 64, 1, 1, 0),
    "f2_e3": (256, 16, 54, 54, 64, 3, 1, 1),
    "f4_e1": (256, 32, 54, 54, 128, 1, 1, 0),
    "f4_e3": (256, 32, 54, 54, 128, 3, 1, 1),
    "f8_sq": (256, 384, 27, 27, 64, 1, 1, 0),
    "f8_e3": (256, 64, 27, 27, 256, 3, 1, 1),
    "conv10": (256, 512, 13, 13, 1000, 1, 1, 0),
    "f2_fused": (256, 16, 54, 54, 128, 3, 1, 1),   # fire2 expand1x1 + expand3x3 as one launch (planner fusion)
    "f5_e1": (256, 32, 27, 27, 128, 1, 1, 0),
    "f6_e1": (256, 48, 27, 27, 192, 1, 1, 0),
    "f6_e3": (256, 48, 27, 27, 192, 3, 1, 1),
    "f8_e1": (256, 64, 27, 27, 256, 1, 1, 0),
    "f9_e1": (256, 64, 13, 13, 256, 1, 1, 0),
    "f9_e3": (256, 64, 13, 13, 256, 3, 1, 1),
    "f9_sq": (256, 512, 13, 13, 64, 1, 1, 0),
    "tiny1": (1, 32, 8, 16, 128, 1, 1, 0),     # one tile, one k-block: the kernel's fixed cost
    "tiny148": (148, 32, 8, 16, 128, 1, 1, 0),  # one tile per SM
}
names = sys.argv[1:] or [n for n in LAYERS if not n.startswith('tiny')]
torch.cuda.set_device(0)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
ctx = L.Context(0, s.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for nm in names:
    N, C, H, W, M, k, st, p = LAYERS[nm]
    rng = np.random.default_rng(0)
    x = L.DeviceTensor.alloc(ctx, (N, C, H, W))      # contents irrelevant for timing (uninitialised memory)
    w = ctx.tensor(rng.standard_normal((M, C, k, k)).astype(np.float32) * 0.05)
    b = ctx.tensor(rng.standard_normal((M,)).astype(np.float32))
    y = L.conv2d(ctx, x, w, bias=b, strides=(st, st), pads=(p,) * 4, fuse_relu=True)
    Ho, Wo = y.shape[2], y.shape[3]
    flops = 2.0 * N * Ho * Wo * M * C * k * k
    byts = 4.0 * (N * C * H * W + N * M * Ho * Wo + M * C * k * k)
    ts = []
    for it in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        L.conv2d(ctx, x, w, bias=b, strides=(st, st), pads=(p,) * 4, fuse_relu=True, y=y)
        e1.record(s)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{nm:8s} {ms:8.3f} ms  {flops/ms/1e9:8.2f} TFLOP/s  {byts/ms/1e6:8.1f} GB/s", flush=True)
