"""Multi-rank host logic on CPU (gloo, world_size 2): contiguous batch split, per-rank run, logits gather.
The per-rank runner is the CPU oracle here (tests may use it); on GPUs it is Engine.run_torch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import MNIST_ONNX
from onnx_rusty_inference_engine_b200 import sharding, synth


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 5, 256, 2048, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.shard_bounds(2048, 8, 3) == (768, 1024)
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import onnx_wire as ow, ref_model as rm
    model = ow.load_model(MNIST_ONNX)
    xs = torch.from_numpy(synth.synthetic_batch(n, chw=(1, 28, 28), seed=3))

    def run_local(shard):
        if shard.shape[0] == 0:
            return torch.empty((0, 10))
        return torch.from_numpy(rm.run_batch(model, shard.numpy(), ["Input3"]))

    out0 = sharding.run_sharded(run_local, xs, dst=0)
    out_all = sharding.run_sharded(run_local, xs, dst=None)
    if rank == 0:
        q.put((out0.numpy(), out_all.numpy()))
    else:
        assert out0 is None and out_all.shape == (n, 10)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 5])
def test_sharded_equals_single_process(n):
    """Sharded result == single-process result, bitwise, for equal (8 = 4+4) and unequal (5 = 3+2) shards."""
    from oracle import onnx_wire as ow, ref_model as rm
    want = rm.run_batch(ow.load_model(MNIST_ONNX), synth.synthetic_batch(n, chw=(1, 28, 28), seed=3), ["Input3"])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got0, got_all = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got0, want) and np.array_equal(got_all, want)
