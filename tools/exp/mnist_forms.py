"""Experiment (GPU): MNIST-8 batch 65,536 in its three plans (fused_cnn = 2 one launch, 1 two launches, 0 node by node)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from onnx_rusty_inference_engine_b200.inference_engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
torch.cuda.set_device(0); s = torch.cuda.Stream(); torch.cuda.set_stream(s)
eng = Engine("tests/golden/mnist-8.onnx", device=0, stream=s.cuda_stream)
x = torch.randn((B, 1, 28, 28), device="cuda") * 10; out = torch.empty((B, 10), device="cuda")
for mode in (2, 1, 0):
    eng.model.set_option("fused_cnn", mode)
    for _ in range(3): eng.run_torch(x, out)
    s.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(20): eng.run_torch(x, out)
    e1.record(s); e1.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"fused_cnn={mode}: {ms:.4f} ms  {B / ms / 1e3:.1f} M img/s  launches {eng.model.launches_per_run(B)}  "
          + str([(p['name'][:14], round(p['ms'], 4)) for p in eng.model.profile(B, iters=3, flush_l2=True)]))
