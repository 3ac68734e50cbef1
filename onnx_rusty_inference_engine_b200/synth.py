"""Synthetic SqueezeNet1.0-8 ONNX generator (seeded weights) -- BASELINE.json configs 3 and 4.

`models/squeezenet1.0-8.onnx` is not shipped with the reference (.MISSING_LARGE_BLOBS), so the benchmark model
is generated: the 66-node zoo topology the reference's operators hard-code (two-input Concat
concatenate_op.rs:15-18, Dropout with `ratio` only dropout_op.rs:24, Softmax tail softmax_op.rs:41):

    Conv(7x7/2) Relu MaxPool | 8 x Fire = [Conv s1x1, Relu, Conv e1x1, Relu, Conv e3x3 pad 1, Relu, Concat]
    with MaxPool after fire4 and fire8 | Dropout Conv(conv10 1x1) Relu GlobalAveragePool Softmax

Attributes are chosen so that the reference's semantics and ONNX's coincide (shape caveat):
every Conv carries kernel_shape/strides/pads; every MaxPool carries kernel_shape=[3,3], strides=[2,2], pads; the
pool after fire4 carries auto_pad="NOTSET" + pads=[0,0,1,1] (the reference honours pads only under NOTSET,
max_pool_op.rs:96,188-201), giving the canonical 224 -> 109 -> 54 -> 27 -> 13 geometry, 1.6378 GFLOP / image.
All initializers are also listed in graph.input (that is where the reference takes their shapes, utils.rs:122).

Weights: Kaiming-uniform U(+-sqrt(6/fan_in)) for convs, N(0, 0.01^2) for conv10, biases U(-0.1, 0.1).
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np

from . import onnx_proto as P

FIRE = [  # (name, squeeze, expand)  torchvision squeezenet1_0 channel plan
    ("fire2", 16, 64), ("fire3", 16, 64), ("fire4", 32, 128), ("fire5", 32, 128),
    ("fire6", 48, 192), ("fire7", 48, 192), ("fire8", 64, 256), ("fire9", 64, 256),
]
POOL_AFTER = {"fire4", "fire8"}
NUM_CLASSES = 1000


def squeezenet_layers() -> List[Tuple[str, int, int, int, int, int, int]]:
    """(name, C_in, M, k, stride, pad, H_out) for the 26 convs at 224x224 input -- used for FLOP accounting."""
    layers = []
    h = (224 - 7) // 2 + 1  # 109
    layers.append(("conv1", 3, 96, 7, 2, 0, h))
    h = (h - 3) // 2 + 1    # 54
    c = 96
    for name, s, e in FIRE:
        layers.append((f"{name}/squeeze1x1", c, s, 1, 1, 0, h))
        layers.append((f"{name}/expand1x1", s, e, 1, 1, 0, h))
        layers.append((f"{name}/expand3x3", s, e, 3, 1, 1, h))
        c = 2 * e
        if name == "fire4":
            h = (h - 3 + 1) // 2 + 1  # 27 (pads [0,0,1,1] under NOTSET)
        elif name == "fire8":
            h = (h - 3) // 2 + 1      # 13
    layers.append(("conv10", c, NUM_CLASSES, 1, 1, 0, h))
    return layers


def conv_flops_per_image() -> float:
    return float(sum(2 * m * c * k * k * h * h for _, c, m, k, _, _, h in squeezenet_layers()))


def build_squeezenet(seed: int = 0, raw: bool = True) -> bytes:
    rng = np.random.default_rng(seed)
    nodes: List[Dict] = []
    inits: List[Dict] = []
    inputs: List[Dict] = [P.make_value_info("data_0", [1, 3, 224, 224])]

    def add_init(name: str, arr: np.ndarray) -> None:
        inits.append(P.make_tensor(name, arr.astype(np.float32), raw=raw))
        inputs.append(P.make_value_info(name, arr.shape))

    def conv(name: str, x: str, cin: int, cout: int, k: int, stride: int, pad: int, final: bool = False) -> str:
        fan_in = cin * k * k
        if final:
            w = rng.normal(0.0, 0.01, size=(cout, cin, k, k))
        else:
            b = np.sqrt(6.0 / fan_in)
            w = rng.uniform(-b, b, size=(cout, cin, k, k))
        bias = rng.uniform(-0.1, 0.1, size=(cout,))
        add_init(f"{name}_w_0", w)
        add_init(f"{name}_b_0", bias)
        y = f"{name}_1"
        nodes.append(P.make_node("Conv", [x, f"{name}_w_0", f"{name}_b_0"], [y], name=name,
                                 kernel_shape=[k, k], strides=[stride, stride], pads=[pad] * 4))
        r = f"{name}_2"
        nodes.append(P.make_node("Relu", [y], [r], name=f"{name}_relu"))
        return r

    def pool(name: str, x: str, ceil_like: bool) -> str:
        y = f"{name}_1"
        if ceil_like:
            nodes.append(P.make_node("MaxPool", [x], [y], name=name, kernel_shape=[3, 3], strides=[2, 2],
                                     pads=[0, 0, 1, 1], auto_pad="NOTSET"))
        else:
            nodes.append(P.make_node("MaxPool", [x], [y], name=name, kernel_shape=[3, 3], strides=[2, 2],
                                     pads=[0, 0, 0, 0]))
        return y

    x = conv("conv1", "data_0", 3, 96, 7, 2, 0)
    x = pool("pool1", x, False)
    c = 96
    for name, s, e in FIRE:
        sq = conv(f"{name}_squeeze1x1", x, c, s, 1, 1, 0)
        e1 = conv(f"{name}_expand1x1", sq, s, e, 1, 1, 0)
        e3 = conv(f"{name}_expand3x3", sq, s, e, 3, 1, 1)
        x = f"{name}_concat_1"
        nodes.append(P.make_node("Concat", [e1, e3], [x], name=f"{name}_concat", axis=1))
        c = 2 * e
        if name in POOL_AFTER:
            x = pool(f"pool_{name}", x, ceil_like=(name == "fire4"))
    nodes.append(P.make_node("Dropout", [x], ["fire9_dropout_1"], name="drop9", ratio=0.5))
    x = conv("conv10", "fire9_dropout_1", c, NUM_CLASSES, 1, 1, 0, final=True)
    nodes.append(P.make_node("GlobalAveragePool", [x], ["pool10_1"], name="pool10"))
    nodes.append(P.make_node("Softmax", ["pool10_1"], ["softmaxout_1"], name="softmax"))
    graph = {"node": nodes, "name": "squeezenet1.0-synth", "initializer": inits, "input": inputs,
             "output": [P.make_value_info("softmaxout_1", [1, NUM_CLASSES, 1, 1])]}
    model = {"ir_version": 3, "producer_name": "b200-synth", "graph": graph,
             "opset_import": [{"domain": "", "version": 8}]}
    return P.encode("ModelProto", model)


def build_fire(cin: int, squeeze: int, expand: int, hw: int, seed: int = 0) -> bytes:
    """One Fire module on its own (test model): data_0 [1, cin, hw, hw] -> squeeze 1x1 -> {expand 1x1, expand 3x3 pad 1}
    -> Concat, which is the graph output.  Weights scaled so that activations stay O(input)."""
    rng = np.random.default_rng(seed)
    nodes: List[Dict] = []
    inits: List[Dict] = []
    inputs: List[Dict] = [P.make_value_info("data_0", [1, cin, hw, hw])]

    def conv(name: str, x: str, ci: int, co: int, k: int, pad: int) -> str:
        b = np.sqrt(6.0 / (ci * k * k))
        for nm, arr in ((f"{name}_w_0", rng.uniform(-b, b, size=(co, ci, k, k))), (f"{name}_b_0", rng.uniform(-0.1, 0.1, size=(co,)))):
            inits.append(P.make_tensor(nm, arr.astype(np.float32), raw=True))
            inputs.append(P.make_value_info(nm, arr.shape))
        nodes.append(P.make_node("Conv", [x, f"{name}_w_0", f"{name}_b_0"], [f"{name}_1"], name=name,
                                 kernel_shape=[k, k], strides=[1, 1], pads=[pad] * 4))
        nodes.append(P.make_node("Relu", [f"{name}_1"], [f"{name}_2"], name=f"{name}_relu"))
        return f"{name}_2"

    sq = conv("fire_squeeze1x1", "data_0", cin, squeeze, 1, 0)
    e1 = conv("fire_expand1x1", sq, squeeze, expand, 1, 0)
    e3 = conv("fire_expand3x3", sq, squeeze, expand, 3, 1)
    nodes.append(P.make_node("Concat", [e1, e3], ["fire_concat_1"], name="fire_concat", axis=1))
    graph = {"node": nodes, "name": "fire-synth", "initializer": inits, "input": inputs,
             "output": [P.make_value_info("fire_concat_1", [1, 2 * expand, hw, hw])]}
    model = {"ir_version": 3, "producer_name": "b200-synth", "graph": graph, "opset_import": [{"domain": "", "version": 8}]}
    return P.encode("ModelProto", model)


def build_pool_squeeze(cin: int, squeeze: int, hw: int, pads=(0, 0, 0, 0), seed: int = 0, tail: bool = False) -> bytes:
    """MaxPool 3x3 / 2 followed by a pointwise squeeze convolution (test model for the pool fusion): data_0 [1, cin, hw, hw]
    -> MaxPool(pads, auto_pad NOTSET: the reference honours pads only then, max_pool_op.rs:96) -> Conv 1x1 + Relu, which is
    the graph output (tail=True: one more 1x1 Conv + Relu behind it, so the fused launch's output is an arena value)."""
    rng = np.random.default_rng(seed)
    nodes: List[Dict] = []
    inits: List[Dict] = []
    inputs: List[Dict] = [P.make_value_info("data_0", [1, cin, hw, hw])]
    nodes.append(P.make_node("MaxPool", ["data_0"], ["pool_1"], name="pool", kernel_shape=[3, 3], strides=[2, 2],
                             pads=list(pads), auto_pad="NOTSET"))
    hp = (hw + pads[0] + pads[2] - 3) // 2 + 1

    def conv(name: str, x: str, ci: int, co: int) -> str:
        b = np.sqrt(6.0 / ci)
        for nm, arr in ((f"{name}_w_0", rng.uniform(-b, b, size=(co, ci, 1, 1))), (f"{name}_b_0", rng.uniform(-0.1, 0.1, size=(co,)))):
            inits.append(P.make_tensor(nm, arr.astype(np.float32), raw=True))
            inputs.append(P.make_value_info(nm, arr.shape))
        nodes.append(P.make_node("Conv", [x, f"{name}_w_0", f"{name}_b_0"], [f"{name}_1"], name=name,
                                 kernel_shape=[1, 1], strides=[1, 1], pads=[0, 0, 0, 0]))
        nodes.append(P.make_node("Relu", [f"{name}_1"], [f"{name}_2"], name=f"{name}_relu"))
        return f"{name}_2"

    y = conv("squeeze1x1", "pool_1", cin, squeeze)
    co = squeeze
    if tail:
        y = conv("tail1x1", y, squeeze, 24)
        co = 24
    graph = {"node": nodes, "name": "pool-squeeze-synth", "initializer": inits, "input": inputs,
             "output": [P.make_value_info(y, [1, co, hp, hp])]}
    model = {"ir_version": 3, "producer_name": "b200-synth", "graph": graph, "opset_import": [{"domain": "", "version": 8}]}
    return P.encode("ModelProto", model)


def ensure_squeezenet(path: str, seed: int = 0) -> str:
    """Write the synthetic model to `path` if it is not there yet (deterministic for a given seed)."""
    if not os.path.exists(path):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = path + f".tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(build_squeezenet(seed))
        os.replace(tmp, path)
    return path


def synthetic_batch(n: int, chw=(3, 224, 224), seed: int = 1, std: float = 10.0) -> np.ndarray:
    """N(0, std^2) images -- the bundled *_data_0.pb inputs are N(0, ~10^2) noise."""
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(size=(n,) + tuple(chw), dtype=np.float32) * np.float32(std)).astype(np.float32)
