"""The reference's executor surface (src/inference_engine/model_inference.rs) over the B200 backend.

Two ways to run a model, both entirely on the device:

  * `inference(model, input_data, input_tensor_name)` -- the reference's node-by-node walk
    (model_inference.rs:29-120, dispatch :128-162): one C-ABI operator call per node, activations stay in HBM
    in a `Store`.  Kept for drop-in parity with the per-op interface and used by the op-level tests.
  * `Engine` -- the graph-level path (b200_model_*): weights uploaded once, Conv+Add+Relu / Concat / Dropout /
    Reshape fused or elided, the node sequence replayed as one CUDA graph.  This is what bench.py times.

The reference returns () and prints; both paths here also return the result array.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from . import inference_fp32_ops as ops
from . import onnx_proto as P
from ._lib import B200Error

_default_ctx: Optional[L.Context] = None


def default_context() -> L.Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = L.Context(0)
    return _default_ctx


def manage_input_data(store: ops.Store, model, input_data: np.ndarray, input_tensor_name: Sequence[str]) -> None:
    """manage_input_data, utils.rs:29-45: names that are initializers are skipped (:35); every other name
    receives the same input_data shaped by graph.input's static dims (:36-40)."""
    g = model.graph
    for name in input_tensor_name:
        if ops.already_into_initializer(g.initializer, name):
            continue
        vi = next((v for v in g.input if v.name == name), None)
        if vi is None:
            raise B200Error(-1, f"input {name} not found in graph.input (utils.rs:36)")
        dims = P.value_info_dims(vi)
        if len(dims) != 4 or any(d < 0 for d in dims):
            raise B200Error(-2, f"input {name}: need 4 static dims (utils.rs:36-40, :67)")
        flat = np.asarray(input_data, dtype=np.float32).reshape(-1)
        if flat.size != int(np.prod(dims)):
            raise B200Error(-1, "input length != static model shape (utils.rs:40 from_shape_vec unwrap)")
        store[name] = (None, store.ctx.tensor(flat.reshape(dims)))


def node_inference(node, store: ops.Store, model, verbose: bool = False) -> None:
    """node_inference, model_inference.rs:128-162."""
    if verbose:
        print(f"INFERENCE ON INPUT(s) {list(node.input)} ON {node.op_type} OPERATION done by MAIN PROCESS")
    g = model.graph
    op = node.op_type
    if op == "Conv":
        ops.convolution(store, node, g.input, g.initializer)
    elif op == "Relu":
        ops.relu(store, node)
    elif op == "MaxPool":
        ops.max_pool(store, node, g.input, g.initializer)
    elif op == "Concat":
        ops.concatenation(store, node)
    elif op == "Dropout":
        ops.drop_out(store, node)
    elif op == "GlobalAveragePool":
        ops.global_average_pool(store, node)
    elif op == "Softmax":
        ops.softmax(store, node)
    elif op == "Reshape":
        ops.reshape(store, node, g.input, g.initializer)
    elif op == "Add":
        ops.add(store, node, g.input, g.initializer)
    elif op == "MatMul":
        ops.mul(store, node)
    else:
        raise B200Error(-2, f"INFERENCE OPERATION '{op}' NOT FOUND FOR NODE {node.name}")  # model_inference.rs:158


def inference(model, input_data, input_tensor_name: Sequence[str], ctx: Optional[L.Context] = None,
              verbose: bool = False) -> np.ndarray:
    """inference(), model_inference.rs:29.  `model` is a decoded ModelProto (onnx_proto.load_model).
    Nodes run in file order on one stream; the reference's branch threads (multithreading/*.rs) only change
    which host thread issues a node, never the result."""
    ctx = ctx or default_context()
    store = ops.Store(ctx)
    manage_input_data(store, model, input_data, input_tensor_name)
    for node in model.graph.node:
        node_inference(node, store, model, verbose)
    res = store.last_result
    if res is None:
        name = model.graph.output[0].name if model.graph.output else model.graph.node[-1].output[0]
        s2, s4 = store[name]
        res = s2 if s2 is not None else s4
    out = res.numpy()
    return out.reshape(out.shape[0], -1)


class Engine:
    """Graph-level executor for one GPU (wraps b200_model)."""

    def __init__(self, onnx_file, device: int = 0, stream: Optional[int] = None, ctx: Optional[L.Context] = None):
        self.ctx = ctx or L.Context(device, stream)
        self.model = L.Model(self.ctx, onnx_file)
        self.in_chw = self.model.in_chw
        self.out_per_image = self.model.out_per_image

    def __call__(self, x: np.ndarray) -> np.ndarray:
        return self.model.run(x)

    def run_torch(self, x, out=None):
        """x: contiguous float32 CUDA tensor [N,C,H,W] on this engine's device; returns [N, out_per_image].
        Ordered against torch: when the engine's stream is not torch's current stream, the engine's stream first waits
        for what torch has enqueued (x is ready), and torch's current stream then waits for the run (out is ready for
        the next torch op or NCCL collective); both tensors are marked as used on the engine's stream so the caching
        allocator does not hand their memory out early.  With the same stream on both sides this is plain stream order."""
        import torch
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        n = x.shape[0]
        if out is None:
            out = torch.empty((n, self.out_per_image), dtype=torch.float32, device=x.device)
        cur = torch.cuda.current_stream(x.device)
        mine = self.ctx.stream
        if cur.cuda_stream == mine:
            self.model.run_raw(x.data_ptr(), n, out.data_ptr(), device=True)
            return out
        es = torch.cuda.ExternalStream(mine, device=x.device)
        es.wait_stream(cur)
        self.model.run_raw(x.data_ptr(), n, out.data_ptr(), device=True)
        cur.wait_stream(es)
        x.record_stream(es)
        out.record_stream(es)
        return out

    def run_pinned_async(self, x_host, out_host) -> None:
        """Pipelined host-to-host (b200_model_run_async): returns immediately; call sync() before reading out_host."""
        self.model.run_async_raw(x_host.data_ptr(), x_host.shape[0], out_host.data_ptr())

    def sync(self) -> None:
        self.model.sync()

    def run_pinned(self, x_host, out_host) -> None:
        """Host-to-host through the C ABI with caller-provided (ideally pinned) torch CPU tensors."""
        self.model.run_raw(x_host.data_ptr(), x_host.shape[0], out_host.data_ptr(), device=False)
