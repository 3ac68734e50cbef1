"""Config 5 (BASELINE.json configs[4]) on one GPU: MNIST-8 at a large synthetic batch, device-resident, CUDA events.
usage: mnist_bench.py [batch=8192] [iters=20]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_rusty_inference_engine_b200.inference_engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.cuda.set_device(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
eng = Engine(os.path.join(root, "tests", "golden", "mnist-8.onnx"), device=0, stream=s.cuda_stream)
if os.environ.get("MNIST_CONV_PATH"):
    eng.model.set_option("conv_path", int(os.environ["MNIST_CONV_PATH"]))
x = torch.randn((B, 1, 28, 28), device="cuda") * 10
out = torch.empty((B, 10), device="cuda")
for _ in range(3):
    eng.run_torch(x, out)
s.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(iters):
    eng.run_torch(x, out)
e1.record(s); e1.synchronize()
ms = e0.elapsed_time(e1) / iters
prof = eng.model.profile(B, iters=3, flush_l2=True)
print(json.dumps({"batch": B, "ms_per_batch": ms, "images_per_s": B / ms * 1e3, "launches": eng.model.launches_per_run(B),
                  "per_launch": [(p["name"], p["kind"], round(p["ms"], 4)) for p in prof]}))
