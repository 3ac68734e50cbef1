"""Small end-to-end runs for compute-sanitizer (one tool per gpurun call): the fused MNIST path on a ragged batch, the
node-by-node MNIST plan, SqueezeNet on 2 images (incl. the finite guard's fallback plan), Concat on every axis, MatMul."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from onnx_rusty_inference_engine_b200 import _lib as L, synth
from onnx_rusty_inference_engine_b200.inference_engine import Engine
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ctx = L.Context(0)
eng = Engine(os.path.join(root, "tests", "golden", "mnist-8.onnx"), ctx=ctx)
x = synth.synthetic_batch(19, chw=(1, 28, 28), seed=1)
a = eng(x)
eng.model.set_option("fused_cnn", 0)
b = eng(x)
print("mnist fused vs node plan max abs diff", float(np.abs(a - b).max()))
sq = Engine(synth.ensure_squeezenet(os.path.join(root, "models", "squeezenet1.0-8-synth.onnx"), seed=0), ctx=ctx)
xs = synth.synthetic_batch(2, seed=3)
y = sq(xs)
xs[1, 0, 5, 5] = np.inf
y2 = sq(xs)
print("squeezenet rows sum", y.sum(1), "fallback finite", bool(np.isfinite(y2[0]).all()))
rng = np.random.default_rng(0)
for ax in range(4):
    sa, sb = [2, 6, 5, 7], [2, 6, 5, 7]; sb[ax] = 3
    p, q = rng.standard_normal(sa, dtype=np.float32), rng.standard_normal(sb, dtype=np.float32)
    assert np.array_equal(L.concat(ctx, ctx.tensor(p), ctx.tensor(q), axis=ax).numpy(), np.concatenate([p, q], ax))
m = L.matmul(ctx, ctx.tensor(rng.standard_normal((7, 256), dtype=np.float32)), ctx.tensor(rng.standard_normal((256, 10), dtype=np.float32))).numpy()
print("ok", m.shape)
