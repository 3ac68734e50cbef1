"""CPU oracle: the reference's ten fp32 operators with the reference's store semantics.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs.  The product path never imports this module.

Every op has the reference's shape `op(store, node[, model_inputs, model_initializers])`
(src/inference_fp32_ops/*.rs) and works on the reference's store type
(model_inference.rs:30-32): name -> (Array2 | None, Array4 | None).  The arithmetic is in
oracle/ref_ops.c (C restatement, loaded through ctypes); attribute parsing, error
behaviour ("panic" -> RefPanic) and store bookkeeping are restated here.

The reference is batch-1 only; arrays in the store always have a leading dim of 1.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import onnx_wire as ow

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libref_ops.so")

PAD_VALID, PAD_SAME_UPPER, PAD_SAME_LOWER, PAD_NOTSET = 0, 1, 2, 3


class RefPanic(RuntimeError):
    """Raised where the reference would panic!/unwrap()-abort."""


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ref_ops.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        f32p = ctypes.POINTER(ctypes.c_float)
        szp = ctypes.POINTER(ctypes.c_size_t)
        llp = ctypes.POINTER(ctypes.c_longlong)
        sz = ctypes.c_size_t
        _lib.ref_conv2d.argtypes = [f32p, sz, sz, sz, f32p, sz, sz, sz, f32p, ctypes.c_int, llp, llp,
                                    f32p, sz, szp, szp]
        _lib.ref_conv2d.restype = ctypes.c_int
        _lib.ref_maxpool2d.argtypes = [f32p, sz, sz, sz, sz, sz, ctypes.c_int, llp, llp, f32p, sz, szp, szp]
        _lib.ref_maxpool2d.restype = ctypes.c_int
        _lib.ref_relu.argtypes = [f32p, sz, f32p]
        _lib.ref_add_channel.argtypes = [f32p, sz, sz, f32p, f32p]
        _lib.ref_add_same.argtypes = [f32p, f32p, sz, f32p]
        _lib.ref_matmul.argtypes = [f32p, f32p, sz, sz, sz, f32p]
        _lib.ref_global_avgpool.argtypes = [f32p, sz, sz, f32p]
        _lib.ref_softmax_row.argtypes = [f32p, sz, f32p]
    return _lib


def _p(a: Optional[np.ndarray]):
    if a is None:
        return ctypes.POINTER(ctypes.c_float)()
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ll(vals) -> ctypes.Array:
    v = [int(x) for x in vals]
    return (ctypes.c_longlong * len(v))(*v)


Store = Dict[str, Tuple[Optional[np.ndarray], Optional[np.ndarray]]]


# ----------------------------------------------------------------------------- utils.rs
def already_into_initializer(model_initializers: List[ow.Tensor], name: str) -> bool:
    """utils.rs:14-21."""
    return any(t.name == name for t in model_initializers)


def get_stored_tensor(i: int, node: ow.Node, model_inputs: List[ow.ValueInfo],
                      model_initializers: List[ow.Tensor]) -> np.ndarray:
    """utils.rs:113-185: decode initializer node.input[i]; the rank comes from graph.input
    (utils.rs:122), falling back to TensorProto.dims when the initializer is not listed there."""
    name = node.input[i]
    t = next((t for t in model_initializers if t.name == name), None)
    if t is None:
        raise RefPanic(f"initializer {name} not found")
    dims = None
    for vi in model_inputs:
        if vi.name == name:
            if any(isinstance(d, str) for d in vi.dims):
                raise RefPanic("DimParam in initializer shape (utils.rs:67)")
            dims = [int(d) for d in vi.dims]
    if dims is None:
        dims = list(t.dims)
    if not 1 <= len(dims) <= 4:
        raise RefPanic(f"initializer {name}: rank {len(dims)} unsupported (utils.rs:146-184)")
    flat = t.array().reshape(-1)
    return flat.reshape(dims)


# ----------------------------------------------------------------------------- operators
def _attr_s(a: ow.Attribute) -> str:
    return a.s.decode("utf-8")


def conv2d_image(x: np.ndarray, w: np.ndarray, bias: Optional[np.ndarray], auto_pad: int, pads, strides):
    """One image [C,H,W] through ref_conv2d (convolution_op.rs:224-517)."""
    x = _c(x); w = _c(w)
    C, H, W = x.shape
    M, Cw, kh, kw = w.shape
    if C != Cw:
        raise RefPanic("assert C_in == w.shape[0]*group (convolution_op.rs:252)")
    b = _c(bias) if bias is not None else None
    if b is not None and b.shape[0] != M:
        raise RefPanic("Bias array has the wrong shape (convolution_op.rs:711)")
    ho, wo = ctypes.c_size_t(), ctypes.c_size_t()
    padv = _ll(pads if len(pads) == 4 else [0, 0, 0, 0])
    st = _ll(strides)
    L = lib()
    if L.ref_conv2d(_p(x), C, H, W, _p(w), M, kh, kw, _p(b), auto_pad, padv, st, _p(None), 0,
                    ctypes.byref(ho), ctypes.byref(wo)):
        raise RefPanic("conv2d geometry panics upstream")
    out = np.empty((M, ho.value, wo.value), dtype=np.float32)
    if L.ref_conv2d(_p(x), C, H, W, _p(w), M, kh, kw, _p(b), auto_pad, padv, st, _p(out), out.size,
                    ctypes.byref(ho), ctypes.byref(wo)):
        raise RefPanic("conv2d failed")
    return out


def convolution(store: Store, node: ow.Node, model_inputs, model_initializers) -> None:
    """convolution(), convolution_op.rs:94-193."""
    if node.input[0] in store:
        x = store[node.input[0]][1]
        if x is None:
            raise RefPanic("Conv input is not 4-D (convolution_op.rs:101 unwrap)")
    else:
        x = get_stored_tensor(0, node, model_inputs, model_initializers)
    if node.input[1] in store:
        w = store[node.input[1]][1]
    else:
        w = get_stored_tensor(1, node, model_inputs, model_initializers)
    if x.ndim != 4 or w.ndim != 4:
        raise RefPanic("Conv operands must be rank 4")
    bias = None
    if len(node.input) > 2:
        bias = get_stored_tensor(2, node, model_inputs, model_initializers)
        if bias.ndim != 1:
            raise RefPanic("Conv bias must be rank 1 (convolution_op.rs:129 unwrap)")
    strides: List[int] = []
    pads: List[int] = []
    auto_pad = PAD_VALID
    group = 1
    dil = None
    for a in node.attribute:
        if a.name == "auto_pad":
            s = _attr_s(a)
            table = {"SAME_UPPER": PAD_SAME_UPPER, "SAME_LOWER": PAD_SAME_LOWER, "VALID": PAD_VALID,
                     "NOT_SET": PAD_NOTSET}  # note: "NOT_SET", convolution_op.rs:143
            if s not in table:
                raise RefPanic(f"Convolution Auto Pad specified not found: {s}")
            auto_pad = table[s]
        elif a.name == "dilations":
            dil = list(a.ints[:2])
        elif a.name == "group":
            group = a.i
        elif a.name == "kernel_shape":
            pass
        elif a.name == "pads":
            pads = list(a.ints)
        elif a.name == "strides":
            strides = list(a.ints)
        else:
            raise RefPanic(f"ATTRIBUTE NAME FOR CONVOLUTION NOT FOUND, {a.name}")
    if pads and any(p > 0 for p in pads[:4]):
        auto_pad = PAD_NOTSET  # convolution_op.rs:169-173
    if len(strides) < 2:
        raise RefPanic("strides attribute is required (convolution_op.rs:285 unwrap)")
    if group != 1 or (dil is not None and (dil[0] != 1 or dil[1] != 1)):
        raise RefPanic("group>1 / dilation>1 are broken upstream and out of scope (SURVEY.md #12)")
    if auto_pad == PAD_NOTSET and len(pads) < 4:
        raise RefPanic("pads required under NOTSET (convolution_op.rs:267 unwrap)")
    out = conv2d_image(x[0], w, bias, auto_pad, pads, strides)
    store[node.output[0]] = (None, out[None])


def maxpool_image(x: np.ndarray, kernel, auto_pad: int, pads, strides) -> np.ndarray:
    x = _c(x)
    C, H, W = x.shape
    ho, wo = ctypes.c_size_t(), ctypes.c_size_t()
    padv = _ll(pads if len(pads) == 4 else [0, 0, 0, 0])
    st = _ll(strides)
    L = lib()
    if L.ref_maxpool2d(_p(x), C, H, W, int(kernel[0]), int(kernel[1]), auto_pad, padv, st, _p(None), 0,
                       ctypes.byref(ho), ctypes.byref(wo)):
        raise RefPanic("max_pool2d geometry panics upstream")
    out = np.empty((C, ho.value, wo.value), dtype=np.float32)
    if L.ref_maxpool2d(_p(x), C, H, W, int(kernel[0]), int(kernel[1]), auto_pad, padv, st, _p(out), out.size,
                       ctypes.byref(ho), ctypes.byref(wo)):
        raise RefPanic("max_pool2d failed")
    return out


def max_pool(store: Store, node: ow.Node, model_inputs, model_initializers) -> None:
    """max_pool(), max_pool_op.rs:65-129."""
    if node.input[0] in store:
        x = store[node.input[0]][1]
    else:
        x = get_stored_tensor(0, node, model_inputs, model_initializers)
    if x is None or x.ndim != 4:
        raise RefPanic("MaxPool input must be rank 4")
    kernel = None
    strides: List[int] = []
    pads: List[int] = []
    auto_pad = PAD_VALID
    for a in node.attribute:
        if a.name == "auto_pad":
            s = _attr_s(a)
            table = {"SAME_UPPER": PAD_SAME_UPPER, "SAME_LOWER": PAD_SAME_LOWER, "VALID": PAD_VALID,
                     "NOTSET": PAD_NOTSET}  # note: "NOTSET", max_pool_op.rs:96
            if s not in table:
                raise RefPanic(f"MaxPool Auto Pad specified not found: {s}")
            auto_pad = table[s]
        elif a.name == "kernel_shape":
            kernel = list(a.ints[:2])
        elif a.name == "pads":
            pads = list(a.ints)
        elif a.name == "storage_order":
            pass
        elif a.name == "strides":
            strides = list(a.ints)
        else:
            raise RefPanic(f"ATTRIBUTE NAME FOR MAX POOL NOT FOUND, {a.name}")
    if kernel is None:
        raise RefPanic("kernel_shape is required (max_pool_op.rs:100)")
    if len(strides) < 2:
        raise RefPanic("strides attribute is required (max_pool_op.rs:207 unwrap)")
    if auto_pad == PAD_NOTSET and len(pads) < 4:
        raise RefPanic("pads required under NOTSET (max_pool_op.rs:189 unwrap)")
    out = maxpool_image(x[0], kernel, auto_pad, pads, strides)
    store[node.output[0]] = (None, out[None])


def relu(store: Store, node: ow.Node) -> None:
    """relu(), relu_op.rs:11-33: 4-D slot only."""
    x = store[node.input[0]][1]
    if x is None:
        raise RefPanic("Relu input has no 4-D slot (relu_op.rs:16 unwrap)")
    x = _c(x)
    out = np.empty_like(x)
    lib().ref_relu(_p(x), x.size, _p(out))
    store[node.output[0]] = (None, out)


def add(store: Store, node: ow.Node, model_inputs, model_initializers) -> None:
    """add(), add_op.rs:16-107."""
    a4 = a2 = None
    if already_into_initializer(model_initializers, node.input[0]):
        t = get_stored_tensor(0, node, model_inputs, model_initializers)
        if t.ndim == 4:
            a4 = t
        elif t.ndim == 2:
            a2 = t
        else:
            raise RefPanic("Cannot retrieve input 1 for Add operation from initializers")
    else:
        if node.input[0] not in store:
            raise RefPanic("Add input 1 missing (add_op.rs:41 unwrap)")
        s2, s4 = store[node.input[0]]
        if s2 is not None:
            a2 = s2
        elif s4 is not None:
            a4 = s4
        else:
            raise RefPanic("Cannot retrieve input 1 for Add operation from hashmap input/output")
    if not already_into_initializer(model_initializers, node.input[1]):
        raise RefPanic("Cannot retrieve input 2 for Add operation")
    b = get_stored_tensor(1, node, model_inputs, model_initializers)
    if b.ndim not in (2, 3):
        raise RefPanic("Cannot retrieve input 2 for Add operation from initializes")
    if a4 is not None:
        if b.ndim != 3 or b.shape[0] != a4.shape[1] or b.shape[1] != 1 or b.shape[2] != 1:
            # ndarray would broadcast other [C|1,H|1,W|1] shapes too; neither model needs them.
            raise RefPanic("Add: only [C,1,1] broadcast is supported")
        x = _c(a4)
        out = np.empty_like(x)
        C, HW = x.shape[1], x.shape[2] * x.shape[3]
        bb = _c(b.reshape(-1))
        lib().ref_add_channel(_p(x), C, HW, _p(bb), _p(out))
        store[node.output[0]] = (None, out)
    else:
        x = _c(a2)
        bb = _c(b)
        if bb.shape != x.shape:
            raise RefPanic("Add: 2-D operands must have the same shape")
        out = np.empty_like(x)
        lib().ref_add_same(_p(x), _p(bb), x.size, _p(out))
        store[node.output[0]] = (out, None)


def mul(store: Store, node: ow.Node) -> None:
    """mul() (ONNX MatMul), mul_op.rs:11-32: both operands from the store's 2-D slot."""
    for nm in node.input[:2]:
        if nm not in store or store[nm][0] is None:
            raise RefPanic("MatMul operand has no 2-D slot (mul_op.rs:17/19 unwrap)")
    a = _c(store[node.input[0]][0]); b = _c(store[node.input[1]][0])
    if a.shape[1] != b.shape[0]:
        raise RefPanic("MatMul shape mismatch (ndarray dot panics)")
    out = np.empty((a.shape[0], b.shape[1]), dtype=np.float32)
    lib().ref_matmul(_p(a), _p(b), a.shape[0], a.shape[1], b.shape[1], _p(out))
    store[node.output[0]] = (out, None)


def reshape(store: Store, node: ow.Node, model_inputs, model_initializers) -> None:
    """reshape(), reshape_op.rs:16-92: output is always 2-D (shape[0], shape[1]); 0 copies the input dim."""
    if already_into_initializer(model_initializers, node.input[0]):
        data = get_stored_tensor(0, node, model_inputs, model_initializers)
    else:
        data = store[node.input[0]][1]
    if data is None or data.ndim != 4:
        raise RefPanic("Reshape data must be rank 4 (reshape_op.rs:27/30 unwrap)")
    if not already_into_initializer(model_initializers, node.input[1]):
        raise RefPanic("Unable to retrieve Shape for Reshape operation")
    shape = get_stored_tensor(1, node, model_inputs, model_initializers)
    if shape.dtype != np.int64 or shape.ndim != 1 or shape.shape[0] < 2:
        raise RefPanic("Reshape shape must be a rank-1 int64 initializer with >= 2 entries")
    new_shape = [int(v) for v in shape]
    for i, v in enumerate(new_shape):
        if v == 0:
            new_shape[i] = data.shape[i]
    d1, d2 = new_shape[0], new_shape[1]
    if d1 < 0 or d2 < 0 or d1 * d2 != data.size:
        raise RefPanic("Reshape: from_shape_vec fails (reshape_op.rs:90 unwrap)")
    store[node.output[0]] = (_c(data).reshape(d1, d2).copy(), None)


def concatenation(store: Store, node: ow.Node) -> None:
    """concatenation(), concatenate_op.rs:11-41: exactly two 4-D inputs."""
    a = store[node.input[0]][1]; b = store[node.input[1]][1]
    if a is None or b is None:
        raise RefPanic("Concat inputs must be 4-D (concatenate_op.rs:16/18 unwrap)")
    axis = 1
    for at in node.attribute:
        if at.name == "axis":
            axis = at.i
        else:
            raise RefPanic(f"ATTRIBUTE NAME FOR CONCATENATE NOT FOUND, {at.name}")
    store[node.output[0]] = (None, np.concatenate([a, b], axis=axis))


def drop_out(store: Store, node: ow.Node) -> None:
    """drop_out(), dropout_op.rs:12-50: identity at inference (:66-71)."""
    x = store[node.input[0]][1]
    if x is None:
        raise RefPanic("Dropout input must be 4-D")
    for at in node.attribute:
        if at.name != "ratio":
            raise RefPanic(f"ATTRIBUTE NAME FOR DROP OUT NOT FOUND, {at.name}")
    store[node.output[0]] = (None, x.copy())


def global_average_pool(store: Store, node: ow.Node) -> None:
    """global_average_pool(), global_average_pool_op.rs:11-52."""
    x = store[node.input[0]][1]
    if x is None:
        raise RefPanic("GlobalAveragePool input must be 4-D")
    x = _c(x)
    out = np.empty((1, x.shape[1], 1, 1), dtype=np.float32)
    lib().ref_global_avgpool(_p(x[0]), x.shape[1], x.shape[2] * x.shape[3], _p(out))
    store[node.output[0]] = (None, out)


def softmax(store: Store, node: ow.Node) -> np.ndarray:
    """softmax(), softmax_op.rs:13-57.  The reference prints the result and does NOT insert it in
    the store (:30-41); the oracle returns it so that callers can compare."""
    x = store[node.input[0]][1]
    if x is None:
        raise RefPanic("Softmax input must be 4-D")
    x = _c(x).reshape(x.shape[0], -1)
    out = np.empty_like(x)
    for r in range(x.shape[0]):
        lib().ref_softmax_row(_p(x[r]), x.shape[1], _p(out[r]))
    return out
