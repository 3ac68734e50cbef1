"""Minimal ONNX / protobuf wire-format reader and writer (pure Python, no `onnx` package).

TEST INFRASTRUCTURE ONLY.  This module belongs to the oracle: it is imported by
`tests/`, by `oracle/ref_model.py`, by the synthetic-model generator and by
`bench.py`'s cpu_baseline leg.  The product path (`libb200rt.so`) has its own,
independent C++ wire reader (`csrc/onnx_wire.cpp`); the two are deliberately
separate implementations so that a parser bug cannot cancel out in parity tests.

The reference parses models with the third-party crates `onnx-protobuf = "0.2.3"` /
`protobuf = "=3.4.0"` (Cargo.toml:16,22; call sites main.rs:29-30 and main.rs:50).
Those crates are not vendored in /root/reference, so the wire format is restated
here from the schema copy the reference ships, `models/onnx.proto`:

  ModelProto   : ir_version=1 (:347) opset_import=8 (:357) graph=7 (:384)
  GraphProto   : node=1 (:445) name=2 initializer=5 (:454) input=11 (:463) output=12 (:464)
  NodeProto    : input=1 output=2 name=3 op_type=4 attribute=5 (:201-214)
  AttributeProto: name=1 (:142) f=2 (:162) i=3 (:163) s=4 (:164) floats=7 ints=8 (:173) type=20 (:159)
  TensorProto  : dims=1 (:527) data_type=2 (:531) float_data=4 (:554) int64_data=7 (:572)
                 name=8 (:575) raw_data=9 (:595)
  ValueInfoProto: name=1 type=2 (:185-188) -> TypeProto.tensor_type=1 (:727)
                 -> elem_type=1 shape=2 (:686-687) -> dim=1 (:674) -> dim_value=1 / dim_param=2 (:664)
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

FLOAT = 1
INT64 = 7

# AttributeProto.AttributeType
AT_FLOAT, AT_INT, AT_STRING, AT_TENSOR, AT_FLOATS, AT_INTS = 1, 2, 3, 4, 6, 7


# --------------------------------------------------------------------------- wire decode
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("varint too long")


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def fields(buf: bytes):
    """Yield (field_number, wire_type, value) for one message. value is int (wt 0/1/5 raw) or bytes (wt 2)."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
            yield fn, wt, v
        elif wt == 1:
            yield fn, wt, buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > n:
                raise ValueError("truncated length-delimited field")
            yield fn, wt, buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            yield fn, wt, buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")


def _packed_varints(b: bytes) -> List[int]:
    out = []
    pos = 0
    while pos < len(b):
        v, pos = _varint(b, pos)
        out.append(_signed64(v))
    return out


# --------------------------------------------------------------------------- data model
@dataclass
class Tensor:
    name: str = ""
    dims: List[int] = field(default_factory=list)
    data_type: int = FLOAT
    # exactly one of these is the storage the file used; `array()` unifies them
    float_data: Optional[np.ndarray] = None
    int64_data: Optional[np.ndarray] = None
    raw_data: Optional[bytes] = None

    def array(self) -> np.ndarray:
        """Decode the way the reference does (utils.rs:124-142): raw_data as LE f32 chunks, else
        float_data, else int64_data.  (For INT64 raw_data we decode as int64, a superset.)"""
        if self.raw_data is not None and len(self.raw_data) > 0:
            if self.data_type == INT64:
                a = np.frombuffer(self.raw_data, dtype="<i8")
            else:
                a = np.frombuffer(self.raw_data, dtype="<f4")
        elif self.float_data is not None and len(self.float_data) > 0:
            a = np.asarray(self.float_data, dtype=np.float32)
        elif self.int64_data is not None and len(self.int64_data) > 0:
            a = np.asarray(self.int64_data, dtype=np.int64)
        else:
            a = np.zeros((0,), dtype=np.float32)
        return a.reshape(self.dims) if self.dims else a


@dataclass
class Attribute:
    name: str = ""
    type: int = 0
    f: float = 0.0
    i: int = 0
    s: bytes = b""
    ints: List[int] = field(default_factory=list)
    floats: List[float] = field(default_factory=list)


@dataclass
class Node:
    op_type: str = ""
    name: str = ""
    input: List[str] = field(default_factory=list)
    output: List[str] = field(default_factory=list)
    attribute: List[Attribute] = field(default_factory=list)

    def attr(self, name: str) -> Optional[Attribute]:
        for a in self.attribute:
            if a.name == name:
                return a
        return None


@dataclass
class ValueInfo:
    name: str = ""
    elem_type: int = 0
    dims: List[object] = field(default_factory=list)  # int, or str for dim_param


@dataclass
class Model:
    ir_version: int = 0
    opset: int = 0
    producer: str = ""
    graph_name: str = ""
    nodes: List[Node] = field(default_factory=list)
    initializers: List[Tensor] = field(default_factory=list)
    inputs: List[ValueInfo] = field(default_factory=list)
    outputs: List[ValueInfo] = field(default_factory=list)

    def initializer(self, name: str) -> Optional[Tensor]:
        for t in self.initializers:
            if t.name == name:
                return t
        return None

    def input_info(self, name: str) -> Optional[ValueInfo]:
        for v in self.inputs:
            if v.name == name:
                return v
        return None


# --------------------------------------------------------------------------- message parsers
def parse_tensor(buf: bytes) -> Tensor:
    t = Tensor()
    fl: List[float] = []
    i64: List[int] = []
    for fn, wt, v in fields(buf):
        if fn == 1:
            t.dims.extend(_packed_varints(v) if wt == 2 else [_signed64(v)])
        elif fn == 2:
            t.data_type = v
        elif fn == 4:
            if wt == 2:
                fl.extend(np.frombuffer(v, dtype="<f4").tolist())
            else:
                fl.append(struct.unpack("<f", v)[0])
        elif fn == 7:
            i64.extend(_packed_varints(v) if wt == 2 else [_signed64(v)])
        elif fn == 8:
            t.name = v.decode("utf-8")
        elif fn == 9:
            t.raw_data = bytes(v)
    if fl:
        t.float_data = np.asarray(fl, dtype=np.float32)
    if i64:
        t.int64_data = np.asarray(i64, dtype=np.int64)
    return t


def parse_attribute(buf: bytes) -> Attribute:
    a = Attribute()
    for fn, wt, v in fields(buf):
        if fn == 1:
            a.name = v.decode("utf-8")
        elif fn == 2:
            a.f = struct.unpack("<f", v)[0]
        elif fn == 3:
            a.i = _signed64(v)
        elif fn == 4:
            a.s = bytes(v)
        elif fn == 7:
            if wt == 2:
                a.floats.extend(np.frombuffer(v, dtype="<f4").tolist())
            else:
                a.floats.append(struct.unpack("<f", v)[0])
        elif fn == 8:
            a.ints.extend(_packed_varints(v) if wt == 2 else [_signed64(v)])
        elif fn == 20:
            a.type = v
    return a


def parse_node(buf: bytes) -> Node:
    n = Node()
    for fn, wt, v in fields(buf):
        if fn == 1:
            n.input.append(v.decode("utf-8"))
        elif fn == 2:
            n.output.append(v.decode("utf-8"))
        elif fn == 3:
            n.name = v.decode("utf-8")
        elif fn == 4:
            n.op_type = v.decode("utf-8")
        elif fn == 5:
            n.attribute.append(parse_attribute(v))
    return n


def parse_value_info(buf: bytes) -> ValueInfo:
    vi = ValueInfo()
    for fn, wt, v in fields(buf):
        if fn == 1:
            vi.name = v.decode("utf-8")
        elif fn == 2:  # TypeProto
            for fn2, _, v2 in fields(v):
                if fn2 == 1:  # tensor_type
                    for fn3, _, v3 in fields(v2):
                        if fn3 == 1:
                            vi.elem_type = v3
                        elif fn3 == 2:  # TensorShapeProto
                            for fn4, _, v4 in fields(v3):
                                if fn4 == 1:  # Dimension
                                    dv: object = 0
                                    for fn5, _, v5 in fields(v4):
                                        if fn5 == 1:
                                            dv = _signed64(v5)
                                        elif fn5 == 2:
                                            dv = v5.decode("utf-8")
                                    vi.dims.append(dv)
    return vi


def parse_model(buf: bytes) -> Model:
    m = Model()
    graph = None
    for fn, wt, v in fields(buf):
        if fn == 1:
            m.ir_version = v
        elif fn == 2:
            m.producer = v.decode("utf-8")
        elif fn == 7:
            graph = v
        elif fn == 8:
            for fn2, _, v2 in fields(v):
                if fn2 == 2:
                    m.opset = v2
    if graph is None:
        raise ValueError("ModelProto has no graph")
    for fn, wt, v in fields(graph):
        if fn == 1:
            m.nodes.append(parse_node(v))
        elif fn == 2:
            m.graph_name = v.decode("utf-8")
        elif fn == 5:
            m.initializers.append(parse_tensor(v))
        elif fn == 11:
            m.inputs.append(parse_value_info(v))
        elif fn == 12:
            m.outputs.append(parse_value_info(v))
    return m


def load_model(path: str) -> Model:
    with open(path, "rb") as f:
        return parse_model(f.read())


def load_tensor_pb(path: str) -> np.ndarray:
    """Serialized TensorProto (.pb), as read by the reference's read_input_data (main.rs:44-53)."""
    with open(path, "rb") as f:
        return parse_tensor(f.read()).array()


# --------------------------------------------------------------------------- wire encode
def _enc_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(fn: int, wt: int) -> bytes:
    return _enc_varint((fn << 3) | wt)


def _ld(fn: int, payload: bytes) -> bytes:
    return _key(fn, 2) + _enc_varint(len(payload)) + payload


def _vi(fn: int, v: int) -> bytes:
    return _key(fn, 0) + _enc_varint(v)


def enc_tensor(name: str, arr: np.ndarray, storage: str = "raw") -> bytes:
    """storage: 'raw' (raw_data, like SqueezeNet zoo files) or 'typed' (float_data/int64_data, like mnist-8)."""
    arr = np.asarray(arr)
    out = b""
    for d in arr.shape:
        out += _vi(1, int(d))
    if arr.dtype == np.int64:
        out += _vi(2, INT64)
        if storage == "raw":
            body = _ld(9, arr.astype("<i8").tobytes())
        else:
            body = _ld(7, b"".join(_enc_varint(int(x)) for x in arr.reshape(-1)))
    else:
        out += _vi(2, FLOAT)
        a32 = arr.astype("<f4")
        body = _ld(9, a32.tobytes()) if storage == "raw" else _ld(4, a32.tobytes())
    return out + _ld(8, name.encode()) + body


def enc_attr_ints(name: str, ints) -> bytes:
    out = _ld(1, name.encode())
    for v in ints:  # unpacked repeated int64, as proto2 writers (CNTK, Caffe2) emit
        out += _vi(8, int(v))
    return out + _vi(20, AT_INTS)


def enc_attr_int(name: str, v: int) -> bytes:
    return _ld(1, name.encode()) + _vi(3, int(v)) + _vi(20, AT_INT)


def enc_attr_float(name: str, v: float) -> bytes:
    return _ld(1, name.encode()) + _key(2, 5) + struct.pack("<f", v) + _vi(20, AT_FLOAT)


def enc_attr_string(name: str, s: str) -> bytes:
    return _ld(1, name.encode()) + _ld(4, s.encode()) + _vi(20, AT_STRING)


def enc_node(op_type: str, inputs, outputs, attrs=(), name: str = "") -> bytes:
    out = b""
    for i in inputs:
        out += _ld(1, i.encode())
    for o in outputs:
        out += _ld(2, o.encode())
    if name:
        out += _ld(3, name.encode())
    out += _ld(4, op_type.encode())
    for a in attrs:
        out += _ld(5, a)
    return out


def enc_value_info(name: str, dims, elem_type: int = FLOAT) -> bytes:
    shape = b""
    for d in dims:
        if isinstance(d, str):
            shape += _ld(1, _ld(2, d.encode()))
        else:
            shape += _ld(1, _vi(1, int(d)))
    tensor_type = _vi(1, elem_type) + _ld(2, shape)
    return _ld(1, name.encode()) + _ld(2, _ld(1, tensor_type))


def enc_model(nodes, initializers, inputs, outputs, graph_name="g", ir_version=3, opset=8,
              producer="b200-oracle") -> bytes:
    g = b""
    for n in nodes:
        g += _ld(1, n)
    g += _ld(2, graph_name.encode())
    for t in initializers:
        g += _ld(5, t)
    for i in inputs:
        g += _ld(11, i)
    for o in outputs:
        g += _ld(12, o)
    m = _vi(1, ir_version) + _ld(2, producer.encode()) + _ld(7, g) + _ld(8, _vi(2, opset))
    return m


def save_tensor_pb(path: str, name: str, arr: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(enc_tensor(name, arr, "raw"))
