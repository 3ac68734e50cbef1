"""CPU oracle: the reference's executor (inference / node_inference / manage_input_data).

TEST INFRASTRUCTURE ONLY (see oracle/ref_ops.py).  Restates
  inference()          model_inference.rs:29-120   (file-order walk; the branch threads of
                                                    multithreading/*.rs only change WHO runs a node,
                                                    never the result, so the walk here is sequential)
  node_inference()     model_inference.rs:128-162
  manage_input_data()  utils.rs:29-45
The reference returns () and only prints; the oracle returns the tensor the reference prints
(Softmax result, softmax_op.rs:41 / the final 2-D Add, add_op.rs:104) or, failing that, the
store entry named by graph.output[0].

Batch-N = N independent batch-1 runs (the reference is batch-1 only).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence

import numpy as np

from . import onnx_wire as ow
from . import ref_ops as R


def manage_input_data(store: R.Store, model: ow.Model, input_data: np.ndarray,
                      input_tensor_name: Sequence[str]) -> None:
    for name in input_tensor_name:
        if R.already_into_initializer(model.initializers, name):
            continue  # utils.rs:35
        vi = model.input_info(name)
        if vi is None or len(vi.dims) != 4 or any(isinstance(d, str) for d in vi.dims):
            raise R.RefPanic(f"input {name}: need 4 static dims in graph.input (utils.rs:36-40, :67)")
        dims = [int(d) for d in vi.dims]
        flat = np.asarray(input_data, dtype=np.float32).reshape(-1)
        if flat.size != int(np.prod(dims)):
            raise R.RefPanic("from_shape_vec: input length != static model shape (utils.rs:40 unwrap)")
        store[name] = (None, flat.reshape(dims).copy())


def node_inference(node: ow.Node, store: R.Store, model: ow.Model) -> Optional[np.ndarray]:
    op = node.op_type
    mi, init = model.inputs, model.initializers
    if op == "Conv":
        R.convolution(store, node, mi, init)
    elif op == "Relu":
        R.relu(store, node)
    elif op == "MaxPool":
        R.max_pool(store, node, mi, init)
    elif op == "Concat":
        R.concatenation(store, node)
    elif op == "Dropout":
        R.drop_out(store, node)
    elif op == "GlobalAveragePool":
        R.global_average_pool(store, node)
    elif op == "Softmax":
        return R.softmax(store, node)
    elif op == "Reshape":
        R.reshape(store, node, mi, init)
    elif op == "Add":
        R.add(store, node, mi, init)
    elif op == "MatMul":
        R.mul(store, node)
    else:
        raise R.RefPanic(f"INFERENCE OPERATION '{op}' NOT FOUND FOR NODE {node.name}")
    return None


def inference(model: ow.Model, input_data: np.ndarray, input_tensor_name: Sequence[str],
              return_store: bool = False):
    """One batch-1 inference; returns the final tensor as a flat-per-image 2-D array [1, K]."""
    store: R.Store = {}
    manage_input_data(store, model, input_data, input_tensor_name)
    result = None
    for node in model.nodes:
        r = node_inference(node, store, model)
        if r is not None:
            result = r
    if result is None:
        name = model.outputs[0].name if model.outputs else model.nodes[-1].output[0]
        a2, a4 = store[name]
        result = a2 if a2 is not None else a4.reshape(a4.shape[0], -1)
    if return_store:
        return result, store
    return result


def default_input_names(model: ow.Model) -> List[str]:
    return [vi.name for vi in model.inputs if not R.already_into_initializer(model.initializers, vi.name)]


def run_batch(model: ow.Model, x: np.ndarray, input_tensor_name: Optional[Sequence[str]] = None,
              threads: int = 1) -> np.ndarray:
    """x: [N, C, H, W].  N independent batch-1 reference runs; ctypes releases the GIL inside the C ops,
    so `threads` images run concurrently on separate host cores."""
    names = list(input_tensor_name) if input_tensor_name else default_input_names(model)
    R.lib()
    xs = [np.ascontiguousarray(x[i], dtype=np.float32) for i in range(x.shape[0])]

    def one(img):
        return inference(model, img, names)[0]

    if threads <= 1 or len(xs) == 1:
        outs = [one(i) for i in xs]
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            outs = list(ex.map(one, xs))
    return np.stack(outs, axis=0)
