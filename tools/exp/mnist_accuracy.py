"""Experiment (GPU): error of the MNIST-8 paths (fused two-launch path, node-by-node plan, CUDA-core conv path) against
the oracle (fp32, the reference's summation order) and against an fp64 evaluation of the same graph, on N(0,10^2) inputs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from onnx_rusty_inference_engine_b200 import synth
from onnx_rusty_inference_engine_b200.inference_engine import Engine
from oracle import onnx_wire as ow, ref_model as rm
ONNX = "tests/golden/mnist-8.onnx"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
xs = synth.synthetic_batch(n, chw=(1, 28, 28), seed=7)
m = ow.load_model(ONNX)
def arr(name):
    return torch.from_numpy(np.asarray(m.initializer(name).array(), dtype=np.float64))
try:
    w1, a1, w2, a2, wm, bm = arr("Parameter5"), arr("Parameter6"), arr("Parameter87"), arr("Parameter88"), arr("Parameter193"), arr("Parameter194")
    x = torch.from_numpy(xs).double()
    y = F.max_pool2d(F.relu(F.conv2d(x, w1, padding=2) + a1[None]), 2, 2)
    y = F.max_pool2d(F.relu(F.conv2d(y, w2, padding=2) + a2[None]), 3, 3)
    ref64 = (y.reshape(n, 256) @ wm.reshape(256, 10) + bm).numpy()
except Exception as e:   # the oracle's tensor container may differ: fall back to the oracle only
    print("fp64 reference unavailable:", e); ref64 = None
want = rm.run_batch(m, xs, ["Input3", "Parameter193"], threads=16)
def stats(name, got, ref):
    r = np.abs(got - ref) / (1e-5 + 1e-4 * np.abs(ref))
    print(f"{name:34s} max err/tol {r.max():.3f}  >1: {int((r > 1).sum())}/{r.size}  p99.9 {np.quantile(r, 0.999):.3f}  rms abs err {np.sqrt(((got - ref) ** 2).mean()):.3e}")
if ref64 is not None:
    stats("oracle vs fp64", want, ref64)
eng = Engine(ONNX)
for label, opts in (("fused", {}), ("node plan (tcgen05 conv2)", {"fused_cnn": 0}), ("node plan, CUDA-core convs", {"fused_cnn": 0, "conv_path": 1})):
    for k, v in opts.items():
        eng.model.set_option(k, v)
    got = eng(xs)
    stats(label + " vs oracle", got, want)
    if ref64 is not None:
        stats(label + " vs fp64", got, ref64)
