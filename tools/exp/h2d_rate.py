"""Experiment (GPU): pinned host -> device copy rate for one bench batch (154 MB), alone and under a running step."""
import torch, time
x = torch.empty(256 * 3 * 224 * 224, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
s = torch.cuda.Stream()
for n in (1, 8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s):
        for _ in range(n): d.copy_(x, non_blocking=True)
    s.synchronize(); dt = time.perf_counter() - t0
    print(f"{n} copies of {x.numel()*4/1e6:.0f} MB: {n*x.numel()*4/dt/1e9:.1f} GB/s ({dt/n*1e3:.2f} ms each)")
