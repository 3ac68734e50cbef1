"""Multi-GPU shard equivalence on the CUDA path (sharded == single-GPU, bitwise, on the same images).

  * b200_model_run_sharded (C ABI, no torch / NCCL) at n = 1 == b200_model_run            -- any GPU box
  * b200_model_run_sharded over 2 devices, equal (8 = 4+4) and unequal (5 = 3+2) shards    -- needs >= 2 GPUs
  * sharding.run_sharded(Engine.run_torch, ...) under torch.distributed / NCCL, 2 ranks,
    engine on its OWN stream (not torch's): exercises the stream ordering in run_torch      -- needs >= 2 GPUs
The 2-GPU tests skip on a one-GPU box; run them with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`.
"""
import os
import socket

import numpy as np
import pytest

from conftest import MNIST_ONNX, assert_close

pytestmark = pytest.mark.gpu


def _n_gpus():
    from onnx_rusty_inference_engine_b200 import _lib as L
    return L.lib().b200_device_count()


def test_run_sharded_one_device_equals_run(synth_onnx):
    from onnx_rusty_inference_engine_b200 import _lib as L, synth
    ctx = L.Context(0)
    m = L.Model(ctx, synth_onnx)
    xs = synth.synthetic_batch(5, seed=31)
    assert np.array_equal(L.run_sharded([m], xs), m.run(xs))
    with pytest.raises(L.B200Error, match="share a context"):
        L.run_sharded([m, m], xs)
    mn = L.Model(ctx, MNIST_ONNX)
    with pytest.raises(L.B200Error, match="different input shape"):
        L.run_sharded([m, mn], xs)


@pytest.mark.parametrize("n_images", [8, 5, 1])
def test_run_sharded_two_devices_bitwise(synth_onnx, n_images):
    """Same kernels, same per-image reduction order, batch-position-invariant tiles => identical bits."""
    from onnx_rusty_inference_engine_b200 import _lib as L, synth
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    models = [L.Model(L.Context(d), synth_onnx) for d in range(2)]
    xs = synth.synthetic_batch(n_images, seed=40 + n_images)
    single = models[0].run(xs)
    sharded = L.run_sharded(models, xs)
    assert np.array_equal(sharded, single), "2-GPU sharded result differs from the 1-GPU result"
    assert np.array_equal(models[1].run(xs), single), "device 1 alone differs from device 0"


def test_run_sharded_two_devices_vs_oracle(synth_onnx):
    from onnx_rusty_inference_engine_b200 import _lib as L, synth
    from oracle import onnx_wire as ow, ref_model as rm
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    models = [L.Model(L.Context(d), synth_onnx) for d in range(2)]
    xs = synth.synthetic_batch(4, seed=77)
    got = L.run_sharded(models, xs)
    want = rm.run_batch(ow.load_model(synth_onnx), xs, threads=4)
    assert_close(got, want, "2-GPU sharded vs oracle")
    assert (got.argmax(1) == want.argmax(1)).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, onnx_path, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from onnx_rusty_inference_engine_b200 import sharding, synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    eng = Engine(onnx_path, device=rank)            # the engine's own (non-blocking) stream, NOT torch's
    xs = torch.from_numpy(synth.synthetic_batch(n, seed=55)).cuda()
    outs = []
    for _ in range(3):                              # repeated: an ordering bug shows as a stale / half-written gather
        outs.append(sharding.run_sharded(eng.run_torch, xs, dst=None).cpu().numpy())
    single = eng.run_torch(xs).cpu().numpy() if rank == 0 else None
    if rank == 0:
        q.put((outs, single))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [8, 5])
def test_torch_nccl_sharded_equals_single_gpu(synth_onnx, n_images):
    import torch.multiprocessing as mp
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, synth_onnx, n_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs, single = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for o in outs:
        assert np.array_equal(o, single), "NCCL-gathered sharded logits differ from the 1-GPU run"
