"""The reference's op-dispatch surface (src/inference_fp32_ops/mod.rs:1-10) over the B200 backend.

Same function names, argument order and error behaviour as the Rust functions; the store keeps the reference's
shape -- name -> (2-D slot, 4-D slot), model_inference.rs:30-32 -- but its values are handles to HBM-resident
tensors (`DeviceTensor`) instead of host `ndarray`s, so activations never leave the device between nodes.
Where the reference panics, these raise `B200Error` (code EINVAL / EUNSUPPORTED); nothing computes on the host.

  convolution(store, node, model_inputs, model_initializers)          convolution_op.rs:94
  max_pool(store, node, model_inputs, model_initializers)             max_pool_op.rs:65
  reshape(store, node, model_inputs, model_initializers)              reshape_op.rs:16
  add(store, node, model_inputs, model_initializers)                  add_op.rs:16
  relu / concatenation / drop_out / global_average_pool / softmax / mul (store, node)
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib as L
from . import onnx_proto as P
from ._lib import B200Error, DeviceTensor

Slot = Tuple[Optional[DeviceTensor], Optional[DeviceTensor]]


class Store(dict):
    """name -> (Array2 slot, Array4 slot) of device tensors, plus the context they live on and a
    per-inference cache of uploaded initializers (the reference re-decodes them on every use, utils.rs:113)."""

    def __init__(self, ctx: L.Context):
        super().__init__()
        self.ctx = ctx
        self.consts: Dict[str, DeviceTensor] = {}
        self.last_result: Optional[DeviceTensor] = None


def _fail(msg: str, code: int = -1):
    raise B200Error(code, msg)


def already_into_initializer(model_initializers, name: str) -> bool:
    """utils.rs:14-21."""
    return any(t.name == name for t in model_initializers)


def _initializer_array(i: int, node, model_inputs, model_initializers) -> np.ndarray:
    """get_stored_tensor (utils.rs:113-185): rank from graph.input (utils.rs:122), else TensorProto.dims."""
    name = node.input[i]
    t = next((t for t in model_initializers if t.name == name), None)
    if t is None:
        _fail(f"initializer {name} not found")
    dims = None
    for vi in model_inputs:
        if vi.name == name:
            dims = P.value_info_dims(vi)
            if any(d < 0 for d in dims):
                _fail("DimParam in an initializer shape (utils.rs:67)", -2)
    a = P.tensor_to_numpy(t)
    if dims is None:
        dims = list(a.shape)
    if not 1 <= len(dims) <= 4:
        _fail(f"initializer {name}: rank {len(dims)} unsupported (utils.rs:146-184)", -2)
    return a.reshape(dims)


def get_stored_tensor(i: int, node, model_inputs, model_initializers, store: Store) -> DeviceTensor:
    """Device-resident get_stored_tensor: uploaded once per store, not once per use."""
    name = node.input[i]
    if name not in store.consts:
        a = _initializer_array(i, node, model_inputs, model_initializers)
        if a.dtype != np.float32:
            _fail(f"initializer {name} is not f32", -2)
        store.consts[name] = store.ctx.tensor(a)
    return store.consts[name]


def _slot4(store: Store, name: str, what: str) -> DeviceTensor:
    if name not in store or store[name][1] is None:
        _fail(f"{what}: {name} has no 4-D slot in the store")
    return store[name][1]


def _attr_str(a) -> str:
    return a.s.decode("utf-8") if isinstance(a.s, (bytes, bytearray)) else str(a.s)


def convolution(store: Store, node, model_inputs, model_initializers) -> None:
    """convolution(), convolution_op.rs:94-193 -> b200_conv2d."""
    ctx = store.ctx
    x = store[node.input[0]][1] if node.input[0] in store else get_stored_tensor(0, node, model_inputs, model_initializers, store)
    w = store[node.input[1]][1] if node.input[1] in store else get_stored_tensor(1, node, model_inputs, model_initializers, store)
    if x is None or w is None:
        _fail("Conv operands must be rank 4 (convolution_op.rs:101,111 unwrap)")
    bias = None
    if len(node.input) > 2:
        bias = get_stored_tensor(2, node, model_inputs, model_initializers, store)
    strides: List[int] = [0, 0]
    pads: List[int] = [0, 0, 0, 0]
    dil = [0, 0]
    group = 0
    auto_pad = L.PAD_VALID  # convolution_op.rs:134
    for a in node.attribute:
        if a.name == "auto_pad":
            table = {"SAME_UPPER": L.PAD_SAME_UPPER, "SAME_LOWER": L.PAD_SAME_LOWER, "VALID": L.PAD_VALID,
                     "NOT_SET": L.PAD_NOTSET}  # sic, convolution_op.rs:143
            s = _attr_str(a)
            if s not in table:
                _fail(f"Convolution Auto Pad specified not found: {s}", -2)
            auto_pad = table[s]
        elif a.name == "dilations":
            dil = list(a.ints[:2])
        elif a.name == "group":
            group = a.i
        elif a.name == "kernel_shape":
            pass  # ignored, taken from the weight (convolution_op.rs:151)
        elif a.name == "pads":
            pads = list(a.ints[:4])
        elif a.name == "strides":
            strides = list(a.ints[:2])
        else:
            _fail(f"ATTRIBUTE NAME FOR CONVOLUTION NOT FOUND, {a.name}", -2)
    y = L.conv2d(ctx, x, w, bias=bias, strides=strides, pads=pads, dilations=dil, group=group, auto_pad=auto_pad)
    store[node.output[0]] = (None, y)


def max_pool(store: Store, node, model_inputs, model_initializers) -> None:
    """max_pool(), max_pool_op.rs:65-129 -> b200_maxpool2d."""
    x = store[node.input[0]][1] if node.input[0] in store else get_stored_tensor(0, node, model_inputs, model_initializers, store)
    if x is None:
        _fail("MaxPool input must be rank 4")
    kernel = [0, 0]
    strides = [0, 0]
    pads = [0, 0, 0, 0]
    auto_pad = L.PAD_VALID  # max_pool_op.rs:88
    for a in node.attribute:
        if a.name == "auto_pad":
            table = {"SAME_UPPER": L.PAD_SAME_UPPER, "SAME_LOWER": L.PAD_SAME_LOWER, "VALID": L.PAD_VALID,
                     "NOTSET": L.PAD_NOTSET}  # sic, max_pool_op.rs:96
            s = _attr_str(a)
            if s not in table:
                _fail(f"MaxPool Auto Pad specified not found: {s}", -2)
            auto_pad = table[s]
        elif a.name == "kernel_shape":
            kernel = list(a.ints[:2])
        elif a.name == "pads":
            pads = list(a.ints[:4])
        elif a.name == "storage_order":
            pass
        elif a.name == "strides":
            strides = list(a.ints[:2])
        else:
            _fail(f"ATTRIBUTE NAME FOR MAX POOL NOT FOUND, {a.name}", -2)
    y = L.maxpool2d(store.ctx, x, kernel=kernel, strides=strides, pads=pads, auto_pad=auto_pad)
    store[node.output[0]] = (None, y)


def relu(store: Store, node) -> None:
    """relu(), relu_op.rs:11-33."""
    x = _slot4(store, node.input[0], "Relu")
    store[node.output[0]] = (None, L.relu(store.ctx, x))


def add(store: Store, node, model_inputs, model_initializers) -> None:
    """add(), add_op.rs:16-107: input 2 must be an initializer; rank-4 + [C,1,1] or rank-2 + rank-2."""
    if already_into_initializer(model_initializers, node.input[0]):
        x = get_stored_tensor(0, node, model_inputs, model_initializers, store)
        if x.rank not in (2, 4):
            _fail("Cannot retrieve input 1 for Add operation from initializers")
    else:
        if node.input[0] not in store:
            _fail("Add input 1 missing from the store (add_op.rs:41 unwrap)")
        s2, s4 = store[node.input[0]]
        x = s2 if s2 is not None else s4
        if x is None:
            _fail("Cannot retrieve input 1 for Add operation from hashmap input/output")
    if not already_into_initializer(model_initializers, node.input[1]):
        _fail("Cannot retrieve input 2 for Add operation", -2)
    b = get_stored_tensor(1, node, model_inputs, model_initializers, store)
    y = L.add(store.ctx, x, b)
    if x.rank == 4:
        store[node.output[0]] = (None, y)
    else:
        store[node.output[0]] = (y, None)
        store.last_result = y  # the reference prints this one (add_op.rs:91-105)


def mul(store: Store, node) -> None:
    """mul() (ONNX MatMul), mul_op.rs:11-32: both operands from the 2-D slot."""
    ops = []
    for nm in node.input[:2]:
        if nm not in store or store[nm][0] is None:
            _fail(f"MatMul operand {nm} has no 2-D slot (mul_op.rs:17,19 unwrap)")
        ops.append(store[nm][0])
    store[node.output[0]] = (L.matmul(store.ctx, ops[0], ops[1]), None)


def reshape(store: Store, node, model_inputs, model_initializers) -> None:
    """reshape(), reshape_op.rs:16-92: shape must be an int64 initializer; output is always 2-D."""
    if already_into_initializer(model_initializers, node.input[0]):
        data = get_stored_tensor(0, node, model_inputs, model_initializers, store)
    else:
        data = _slot4(store, node.input[0], "Reshape")
    if data.rank != 4:
        _fail("Reshape data must be rank 4 (reshape_op.rs:27,30 unwrap)")
    if not already_into_initializer(model_initializers, node.input[1]):
        _fail("Unable to retrieve Shape for Reshape operation", -2)
    shape = _initializer_array(1, node, model_inputs, model_initializers)
    if shape.dtype != np.int64 or shape.ndim != 1:
        _fail("Reshape shape must be a rank-1 int64 initializer (utils.rs:170-183)")
    store[node.output[0]] = (L.reshape(store.ctx, data, [int(v) for v in shape]), None)


def concatenation(store: Store, node) -> None:
    """concatenation(), concatenate_op.rs:11-41: exactly two 4-D inputs, `axis` is the only attribute."""
    a = _slot4(store, node.input[0], "Concat")
    b = _slot4(store, node.input[1], "Concat")
    axis = 1
    for at in node.attribute:
        if at.name == "axis":
            axis = at.i
        else:
            _fail(f"ATTRIBUTE NAME FOR CONCATENATE NOT FOUND, {at.name}", -2)
    store[node.output[0]] = (None, L.concat(store.ctx, a, b, axis=axis))


def drop_out(store: Store, node) -> None:
    """drop_out(), dropout_op.rs:12-50: `ratio` is the only attribute; identity at inference."""
    x = _slot4(store, node.input[0], "Dropout")
    ratio = 0.5
    for at in node.attribute:
        if at.name == "ratio":
            ratio = at.f
        else:
            _fail(f"ATTRIBUTE NAME FOR DROP OUT NOT FOUND, {at.name}", -2)
    store[node.output[0]] = (None, L.dropout(store.ctx, x, ratio))


def global_average_pool(store: Store, node) -> None:
    """global_average_pool(), global_average_pool_op.rs:11-52."""
    x = _slot4(store, node.input[0], "GlobalAveragePool")
    store[node.output[0]] = (None, L.global_avgpool(store.ctx, x))


def softmax(store: Store, node) -> None:
    """softmax(), softmax_op.rs:13-57.  The reference prints the result and does not insert it in the store
    (softmax_op.rs:30-41); here it is kept in store.last_result so that callers can read it back."""
    x = _slot4(store, node.input[0], "Softmax")
    store.last_result = L.softmax(store.ctx, x)
