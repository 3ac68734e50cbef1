"""ctypes binding of libb200rt.so (include/b200rt.h).  No CPU fallback: if the library is missing the
import fails loudly, and if there is no sm_100 GPU every compute entry point raises B200Error."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200RT_LIB") or os.path.join(_HERE, "lib", "libb200rt.so")   # B200RT_LIB: A/B builds (tools/exp)

PAD_VALID, PAD_SAME_UPPER, PAD_SAME_LOWER, PAD_NOTSET = 0, 1, 2, 3
ERR_NAMES = {-1: "EINVAL", -2: "EUNSUPPORTED", -3: "ECUDA", -4: "ENOMEM", -5: "EPARSE", -6: "ENODEVICE"}


class B200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200rt {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class ConvParams(C.Structure):
    _fields_ = [("strides", C.c_int64 * 2), ("pads", C.c_int64 * 4), ("dilations", C.c_int64 * 2),
                ("group", C.c_int64), ("auto_pad", C.c_int32), ("fuse_relu", C.c_int32)]


class PoolParams(C.Structure):
    _fields_ = [("kernel", C.c_int64 * 2), ("strides", C.c_int64 * 2), ("pads", C.c_int64 * 4),
                ("auto_pad", C.c_int32), ("reserved", C.c_int32)]


_vp = C.c_void_p
_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)

# name -> (restype, argtypes): every symbol include/b200rt.h declares
SIGNATURES = {
    "b200_last_error": (C.c_char_p, []),
    "b200_version": (C.c_char_p, []),
    "b200_device_count": (C.c_int, []),
    "b200_ctx_create": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "b200_ctx_destroy": (C.c_int, [_vp]),
    "b200_sync": (C.c_int, [_vp]),
    "b200_ctx_launch_count": (C.c_int64, [_vp]),
    "b200_ctx_stream": (_vp, [_vp]),
    "b200_tensor_alloc": (C.c_int, [_vp, _i64p, C.c_int, C.POINTER(_vp)]),
    "b200_tensor_upload": (C.c_int, [_vp, _vp, C.c_size_t]),
    "b200_tensor_download": (C.c_int, [_vp, _vp, C.c_size_t]),
    "b200_tensor_rank": (C.c_int, [_vp]),
    "b200_tensor_dims": (C.c_int, [_vp, _i64p]),
    "b200_tensor_view_channels": (C.c_int, [_vp, C.c_int64, C.c_int64, C.POINTER(_vp)]),
    "b200_tensor_free": (C.c_int, [_vp]),
    "b200_conv2d": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.POINTER(ConvParams), C.POINTER(_vp)]),
    "b200_conv2d_out_dims": (C.c_int, [_i64p, _i64p, C.POINTER(ConvParams), _i64p]),
    "b200_maxpool2d": (C.c_int, [_vp, _vp, C.POINTER(PoolParams), C.POINTER(_vp)]),
    "b200_maxpool2d_out_dims": (C.c_int, [_i64p, C.POINTER(PoolParams), _i64p]),
    "b200_relu": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "b200_add": (C.c_int, [_vp, _vp, _vp, C.POINTER(_vp)]),
    "b200_matmul": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "b200_reshape": (C.c_int, [_vp, _vp, _i64p, C.c_int, C.POINTER(_vp)]),
    "b200_concat": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.POINTER(_vp)]),
    "b200_dropout": (C.c_int, [_vp, _vp, C.c_float, C.POINTER(_vp)]),
    "b200_global_avgpool": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "b200_softmax": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "b200_model_load_onnx": (C.c_int, [_vp, _vp, C.c_size_t, C.POINTER(_vp)]),
    "b200_model_load_file": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "b200_model_free": (C.c_int, [_vp]),
    "b200_model_io": (C.c_int, [_vp, _i64p, _i64p]),
    "b200_model_run": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "b200_model_run_device": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "b200_model_run_async": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "b200_model_sync": (C.c_int, [_vp]),
    "b200_model_run_sharded": (C.c_int, [C.POINTER(_vp), C.c_int, _vp, C.c_int64, _vp]),
    "b200_model_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "b200_model_profile": (C.c_int, [_vp, C.c_int64, C.c_int, C.c_int, C.c_char_p, C.c_size_t]),
    "b200_model_launches_per_run": (C.c_int64, [_vp, C.c_int64]),
    "b200_model_arena_bytes": (C.c_int, [_vp, C.c_int64, _i64p, _i64p]),
    "b200_tensorproto_read": (C.c_int, [_vp, C.c_size_t, _vp, C.c_size_t, _i64p, C.POINTER(C.c_int),
                                        C.POINTER(C.c_size_t)]),
}

_lib = None


def lib():
    """Load libb200rt.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here means the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise B200Error(rc, lib().b200_last_error().decode("utf-8", "replace"))


def _i64arr(vals: Sequence[int], n: Optional[int] = None):
    v = [int(x) for x in vals]
    if n is not None:
        v = (v + [0] * n)[:n]
    return (C.c_int64 * len(v))(*v)


class Context:
    """One per GPU (b200_ctx).  `stream` is an optional raw cudaStream_t (int)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._h = _vp()
        check(lib().b200_ctx_create(int(device), _vp(stream) if stream else None, C.byref(self._h)))
        self.device = device

    @property
    def handle(self):
        return self._h

    def sync(self) -> None:
        check(lib().b200_sync(self._h))

    def launch_count(self) -> int:
        return int(lib().b200_ctx_launch_count(self._h))

    @property
    def stream(self) -> int:
        """The context's cudaStream_t as an integer (0 = the legacy default stream)."""
        return int(lib().b200_ctx_stream(self._h) or 0)

    def close(self) -> None:
        if self._h:
            lib().b200_ctx_destroy(self._h)
            self._h = _vp()

    def tensor(self, array: np.ndarray) -> "DeviceTensor":
        a = np.ascontiguousarray(array, dtype=np.float32)
        t = DeviceTensor.alloc(self, a.shape)
        t.upload(a)
        return t


class DeviceTensor:
    """A value of the store: a handle to an HBM-resident fp32 tensor (b200_tensor)."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle

    @staticmethod
    def alloc(ctx: Context, dims: Sequence[int]) -> "DeviceTensor":
        h = _vp()
        check(lib().b200_tensor_alloc(ctx.handle, _i64arr(dims), len(dims), C.byref(h)))
        return DeviceTensor(ctx, h)

    @property
    def handle(self):
        return self._h

    @property
    def rank(self) -> int:
        return int(lib().b200_tensor_rank(self._h))

    @property
    def shape(self):
        d = (C.c_int64 * 4)()
        check(lib().b200_tensor_dims(self._h, d))
        return tuple(int(d[i]) for i in range(self.rank))

    def upload(self, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a, dtype=np.float32)
        check(lib().b200_tensor_upload(self._h, a.ctypes.data_as(_vp), a.size))

    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.float32)
        check(lib().b200_tensor_download(self._h, out.ctypes.data_as(_vp), out.size))
        return out

    def view_channels(self, c_off: int, c_len: int) -> "DeviceTensor":
        h = _vp()
        check(lib().b200_tensor_view_channels(self._h, c_off, c_len, C.byref(h)))
        return DeviceTensor(self.ctx, h)

    def free(self) -> None:
        if self._h:
            lib().b200_tensor_free(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _out(y: Optional[DeviceTensor]):
    return _vp(y.handle.value) if y is not None else _vp()


def _wrap(ctx: Context, y: Optional[DeviceTensor], h) -> DeviceTensor:
    return y if y is not None else DeviceTensor(ctx, h)


def _h(t: Optional[DeviceTensor]):
    return t.handle if t is not None else None


# ----------------------------------------------------------------------------- thin op wrappers
def conv2d(ctx, x, w, bias=None, chan_add=None, strides=(0, 0), pads=(0, 0, 0, 0), dilations=(0, 0), group=0,
           auto_pad=PAD_VALID, fuse_relu=False, y=None) -> DeviceTensor:
    p = ConvParams(_i64arr(strides, 2), _i64arr(pads, 4), _i64arr(dilations, 2), int(group), int(auto_pad),
                   1 if fuse_relu else 0)
    h = _out(y)
    check(lib().b200_conv2d(ctx.handle, x.handle, w.handle, _h(bias), _h(chan_add), C.byref(p), C.byref(h)))
    return _wrap(ctx, y, h)


def maxpool2d(ctx, x, kernel=(0, 0), strides=(0, 0), pads=(0, 0, 0, 0), auto_pad=PAD_VALID, y=None) -> DeviceTensor:
    p = PoolParams(_i64arr(kernel, 2), _i64arr(strides, 2), _i64arr(pads, 4), int(auto_pad), 0)
    h = _out(y)
    check(lib().b200_maxpool2d(ctx.handle, x.handle, C.byref(p), C.byref(h)))
    return _wrap(ctx, y, h)


def _unary(fn_name):
    def f(ctx, x, y=None) -> DeviceTensor:
        h = _out(y)
        check(getattr(lib(), fn_name)(ctx.handle, x.handle, C.byref(h)))
        return _wrap(ctx, y, h)
    return f


relu = _unary("b200_relu")
global_avgpool = _unary("b200_global_avgpool")
softmax = _unary("b200_softmax")


def add(ctx, x, b, y=None) -> DeviceTensor:
    h = _out(y)
    check(lib().b200_add(ctx.handle, x.handle, b.handle, C.byref(h)))
    return _wrap(ctx, y, h)


def matmul(ctx, a, b, bias=None, y=None) -> DeviceTensor:
    h = _out(y)
    check(lib().b200_matmul(ctx.handle, a.handle, b.handle, _h(bias), C.byref(h)))
    return _wrap(ctx, y, h)


def reshape(ctx, x, shape, y=None) -> DeviceTensor:
    h = _out(y)
    s = _i64arr(shape)
    check(lib().b200_reshape(ctx.handle, x.handle, s, len(shape), C.byref(h)))
    return _wrap(ctx, y, h)


def concat(ctx, a, b, axis=1, y=None) -> DeviceTensor:
    h = _out(y)
    check(lib().b200_concat(ctx.handle, a.handle, b.handle, int(axis), C.byref(h)))
    return _wrap(ctx, y, h)


def dropout(ctx, x, ratio=0.5, y=None) -> DeviceTensor:
    h = _out(y)
    check(lib().b200_dropout(ctx.handle, x.handle, float(ratio), C.byref(h)))
    return _wrap(ctx, y, h)


def read_tensor_pb(path: str) -> np.ndarray:
    """read_input_data (main.rs:44-53) through the library's TensorProto reader."""
    with open(path, "rb") as f:
        data = f.read()
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    n = C.c_size_t()
    rank = C.c_int()
    dims = (C.c_int64 * 8)()
    check(lib().b200_tensorproto_read(C.cast(buf, _vp), len(data), None, 0, dims, C.byref(rank), C.byref(n)))
    out = np.empty((n.value,), dtype=np.float32)
    check(lib().b200_tensorproto_read(C.cast(buf, _vp), len(data), out.ctypes.data_as(_vp), out.size, dims,
                                      C.byref(rank), C.byref(n)))
    shape = [int(dims[i]) for i in range(rank.value)]
    return out.reshape(shape) if shape and int(np.prod(shape)) == out.size else out


class Model:
    """b200_model: inference() with the whole node walk resident on the device."""

    def __init__(self, ctx: Context, onnx_path_or_bytes):
        self.ctx = ctx
        self._h = _vp()
        if isinstance(onnx_path_or_bytes, (bytes, bytearray)):
            data = bytes(onnx_path_or_bytes)
            buf = (C.c_char * len(data)).from_buffer_copy(data)
            check(lib().b200_model_load_onnx(ctx.handle, C.cast(buf, _vp), len(data), C.byref(self._h)))
        else:
            check(lib().b200_model_load_file(ctx.handle, str(onnx_path_or_bytes).encode(), C.byref(self._h)))
        chw = (C.c_int64 * 3)()
        opi = C.c_int64()
        check(lib().b200_model_io(self._h, chw, C.byref(opi)))
        self.in_chw = tuple(int(v) for v in chw)
        self.out_per_image = int(opi.value)

    def set_option(self, key: str, value: int) -> None:
        check(lib().b200_model_set_option(self._h, key.encode(), int(value)))

    def run(self, x: np.ndarray) -> np.ndarray:
        """Host-to-host: x [N,C,H,W] float32 -> [N, out_per_image]."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        if tuple(x.shape[1:]) != self.in_chw:
            raise B200Error(-1, f"input shape {x.shape[1:]} != model input {self.in_chw}")
        out = np.empty((n, self.out_per_image), dtype=np.float32)
        check(lib().b200_model_run(self._h, x.ctypes.data_as(_vp), n, out.ctypes.data_as(_vp)))
        return out

    def run_raw(self, in_ptr: int, batch: int, out_ptr: int, device: bool) -> None:
        fn = lib().b200_model_run_device if device else lib().b200_model_run
        check(fn(self._h, _vp(in_ptr), int(batch), _vp(out_ptr)))

    def run_async_raw(self, in_ptr: int, batch: int, out_ptr: int) -> None:
        check(lib().b200_model_run_async(self._h, _vp(in_ptr), int(batch), _vp(out_ptr)))

    def sync(self) -> None:
        check(lib().b200_model_sync(self._h))

    def arena_bytes(self, batch: int):
        """(bytes of the liveness-coloured activation arena, bytes without reuse) for a batch size."""
        a, b = C.c_int64(), C.c_int64()
        check(lib().b200_model_arena_bytes(self._h, int(batch), C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def launches_per_run(self, batch: int) -> int:
        return int(lib().b200_model_launches_per_run(self._h, int(batch)))

    def profile(self, batch: int, iters: int = 5, flush_l2: bool = True):
        import json
        buf = C.create_string_buffer(1 << 20)
        # flush_l2: False / True, or "in_order" (2): the launch list runs in model order, each launch timed in place
        mode = 2 if flush_l2 == "in_order" else (1 if flush_l2 else 0)
        check(lib().b200_model_profile(self._h, int(batch), int(iters), mode, buf, len(buf)))
        return json.loads(buf.value.decode())

    def close(self) -> None:
        if self._h:
            lib().b200_model_free(self._h)
            self._h = _vp()


def run_sharded(models: Sequence["Model"], x: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
    """b200_model_run_sharded: x [N,C,H,W] split contiguously over `models` (one per device), logits in image order."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = x.shape[0]
    if out is None:
        out = np.empty((n, models[0].out_per_image), dtype=np.float32)
    arr = (_vp * len(models))(*[m._h for m in models])
    check(lib().b200_model_run_sharded(arr, len(models), x.ctypes.data_as(_vp), n, out.ctypes.data_as(_vp)))
    return out
