"""Host-side ONNX (protobuf wire format) reader/writer for the Python entry points.

Replaces the crates `onnx-protobuf = "0.2.3"` / `protobuf = "=3.4.0"` (Cargo.toml:16,22; call sites
main.rs:29-30, main.rs:50).  The graph-level hot path parses ONNX inside libb200rt.so (csrc/onnx_wire.cpp);
this module exists so that the Python mirror of the reference's per-op interface
(`inference_fp32_ops.convolution(store, node, model_inputs, model_initializers)` ...) has NodeProto /
TensorProto / ValueInfoProto objects to pass around, and so that the synthetic SqueezeNet generator can
write a model file.  Schema-driven: one table per message, field numbers from models/onnx.proto.
"""
from __future__ import annotations

import struct
from types import SimpleNamespace
from typing import Any, Dict, List, Tuple

import numpy as np

# kind: 'v' varint int64, 'f' float32, 's' string, 'b' bytes, 'm:<Msg>' message; '*' suffix = repeated
SCHEMA: Dict[str, Dict[int, Tuple[str, str]]] = {
    "ModelProto": {1: ("ir_version", "v"), 2: ("producer_name", "s"), 7: ("graph", "m:GraphProto"),
                   8: ("opset_import", "m:OperatorSetId*")},                      # onnx.proto:347-384
    "OperatorSetId": {1: ("domain", "s"), 2: ("version", "v")},
    "GraphProto": {1: ("node", "m:NodeProto*"), 2: ("name", "s"), 5: ("initializer", "m:TensorProto*"),
                   11: ("input", "m:ValueInfoProto*"), 12: ("output", "m:ValueInfoProto*")},  # :445-468
    "NodeProto": {1: ("input", "s*"), 2: ("output", "s*"), 3: ("name", "s"), 4: ("op_type", "s"),
                  5: ("attribute", "m:AttributeProto*")},                          # :201-214
    "AttributeProto": {1: ("name", "s"), 2: ("f", "f"), 3: ("i", "v"), 4: ("s", "b"), 7: ("floats", "f*"),
                       8: ("ints", "v*"), 20: ("type", "v")},                      # :142-173
    "TensorProto": {1: ("dims", "v*"), 2: ("data_type", "v"), 4: ("float_data", "f*"), 7: ("int64_data", "v*"),
                    8: ("name", "s"), 9: ("raw_data", "b")},                       # :527-595
    "ValueInfoProto": {1: ("name", "s"), 2: ("type", "m:TypeProto")},              # :185-188
    "TypeProto": {1: ("tensor_type", "m:TypeTensor")},                             # :727
    "TypeTensor": {1: ("elem_type", "v"), 2: ("shape", "m:TensorShapeProto")},     # :686-687
    "TensorShapeProto": {1: ("dim", "m:Dimension*")},                              # :674
    "Dimension": {1: ("dim_value", "v"), 2: ("dim_param", "s")},                   # :664
}
FLOAT, INT64 = 1, 7


def _rd_varint(b: bytes, p: int) -> Tuple[int, int]:
    v = s = 0
    while True:
        c = b[p]; p += 1
        v |= (c & 0x7F) << s
        if c < 0x80:
            return v, p
        s += 7
        if s > 63:
            raise ValueError("varint overflow")


def _sint(v: int) -> int:
    return v - (1 << 64) if v >> 63 else v


def decode(msg: str, buf: bytes) -> SimpleNamespace:
    """Decode one message of type `msg` into a namespace with every schema field present."""
    sch = SCHEMA[msg]
    out: Dict[str, Any] = {}
    for name, kind in sch.values():
        out[name] = [] if kind.endswith("*") else (None if kind.startswith("m:") else {"v": 0, "f": 0.0, "s": "", "b": b""}[kind[0]])
    p, n = 0, len(buf)
    while p < n:
        key, p = _rd_varint(buf, p)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            val, p = _rd_varint(buf, p); raw = None
        elif wt == 1:
            raw = buf[p:p + 8]; p += 8; val = None
        elif wt == 5:
            raw = buf[p:p + 4]; p += 4; val = None
        elif wt == 2:
            ln, p = _rd_varint(buf, p)
            if p + ln > n:
                raise ValueError("truncated field")
            raw = buf[p:p + ln]; p += ln; val = None
        else:
            raise ValueError(f"wire type {wt} unsupported")
        if fn not in sch:
            continue
        name, kind = sch[fn]
        rep = kind.endswith("*")
        k = kind.rstrip("*")
        if k == "v":
            if wt == 2:  # packed
                q, vals = 0, []
                while q < len(raw):
                    x, q = _rd_varint(raw, q); vals.append(_sint(x))
            else:
                vals = [_sint(val)]
        elif k == "f":
            vals = list(np.frombuffer(raw, dtype="<f4")) if wt == 2 else [struct.unpack("<f", raw)[0]]
        elif k == "s":
            vals = [raw.decode("utf-8")]
        elif k == "b":
            vals = [bytes(raw)]
        else:
            vals = [decode(k[2:], raw)]
        if rep:
            out[name].extend(vals)
        else:
            out[name] = vals[-1]
    return SimpleNamespace(**out)


def load_model(path: str) -> SimpleNamespace:
    """ModelProto::parse_from_bytes (main.rs:29-30)."""
    with open(path, "rb") as f:
        m = decode("ModelProto", f.read())
    if m.graph is None:
        raise ValueError("ModelProto has no graph")
    return m


def tensor_to_numpy(t: SimpleNamespace) -> np.ndarray:
    """Initializer decode with the reference's precedence (utils.rs:124-142): raw_data, float_data, int64_data."""
    if t.raw_data:
        a = np.frombuffer(t.raw_data, dtype="<i8" if t.data_type == INT64 else "<f4")
    elif len(t.float_data):
        a = np.asarray(t.float_data, dtype=np.float32)
    elif len(t.int64_data):
        a = np.asarray(t.int64_data, dtype=np.int64)
    else:
        a = np.zeros((0,), np.float32)
    return a.reshape([int(d) for d in t.dims]) if t.dims else a


def value_info_dims(vi: SimpleNamespace) -> List[int]:
    """Static dims of a ValueInfoProto (get_input_data_shape, utils.rs:53-97); -1 for symbolic dims."""
    tt = vi.type.tensor_type if vi.type is not None else None
    if tt is None or tt.shape is None:
        return []
    return [int(d.dim_value) if not d.dim_param else -1 for d in tt.shape.dim]


# ----------------------------------------------------------------------------- encoding (synthetic models, .pb files)
def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    o = bytearray()
    while v >= 0x80:
        o.append((v & 0x7F) | 0x80); v >>= 7
    o.append(v)
    return bytes(o)


def _fld(fn: int, wt: int, payload: bytes) -> bytes:
    return _varint((fn << 3) | wt) + (_varint(len(payload)) + payload if wt == 2 else payload)


def encode(msg: str, obj: Dict[str, Any]) -> bytes:
    """Inverse of decode for dict input; fields are emitted in schema order."""
    out = b""
    for fn, (name, kind) in SCHEMA[msg].items():
        if name not in obj or obj[name] is None:
            continue
        rep = kind.endswith("*")
        k = kind.rstrip("*")
        vals = obj[name] if rep else [obj[name]]
        if k == "f" and rep:
            if len(vals):
                out += _fld(fn, 2, np.asarray(vals, dtype="<f4").tobytes())
            continue
        for v in vals:
            if k == "v":
                out += _fld(fn, 0, _varint(int(v)))
            elif k == "f":
                out += _fld(fn, 5, struct.pack("<f", float(v)))
            elif k == "s":
                out += _fld(fn, 2, v.encode("utf-8"))
            elif k == "b":
                out += _fld(fn, 2, bytes(v))
            else:
                out += _fld(fn, 2, encode(k[2:], v))
    return out


def make_tensor(name: str, arr: np.ndarray, raw: bool = True) -> Dict[str, Any]:
    arr = np.asarray(arr)
    t: Dict[str, Any] = {"dims": list(arr.shape), "name": name}
    if arr.dtype == np.int64:
        t["data_type"] = INT64
        if raw:
            t["raw_data"] = arr.astype("<i8").tobytes()
        else:
            t["int64_data"] = [int(x) for x in arr.reshape(-1)]
    else:
        t["data_type"] = FLOAT
        if raw:
            t["raw_data"] = arr.astype("<f4").tobytes()
        else:
            t["float_data"] = arr.astype(np.float32).reshape(-1)
    return t


def make_value_info(name: str, dims) -> Dict[str, Any]:
    return {"name": name, "type": {"tensor_type": {"elem_type": FLOAT, "shape": {
        "dim": [({"dim_param": d} if isinstance(d, str) else {"dim_value": int(d)}) for d in dims]}}}}


def make_attr(name: str, value) -> Dict[str, Any]:
    if isinstance(value, str):
        return {"name": name, "s": value.encode(), "type": 3}
    if isinstance(value, float):
        return {"name": name, "f": value, "type": 1}
    if isinstance(value, int):
        return {"name": name, "i": value, "type": 2}
    return {"name": name, "ints": [int(v) for v in value], "type": 7}


def make_node(op_type: str, inputs, outputs, name: str = "", **attrs) -> Dict[str, Any]:
    return {"input": list(inputs), "output": list(outputs), "name": name, "op_type": op_type,
            "attribute": [make_attr(k, v) for k, v in attrs.items()]}


def read_tensor_pb(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        return tensor_to_numpy(decode("TensorProto", f.read()))


def write_tensor_pb(path: str, name: str, arr: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(encode("TensorProto", make_tensor(name, arr)))
