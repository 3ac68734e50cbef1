"""ONNX wire-format readers/writers (no GPU): the package's host-side reader, the oracle's independent reader and
the library's C++ reader must agree on the reference's bundled files and on the synthetic model."""
import ctypes as C
import os

import numpy as np

from conftest import GOLDEN, MNIST_ONNX
from onnx_rusty_inference_engine_b200 import _lib as L
from onnx_rusty_inference_engine_b200 import onnx_proto as P
from onnx_rusty_inference_engine_b200 import synth
from oracle import onnx_wire as ow


def _same_model(pm, om):
    assert [n.op_type for n in pm.graph.node] == [n.op_type for n in om.nodes]
    for a, b in zip(pm.graph.node, om.nodes):
        assert list(a.input) == b.input and list(a.output) == b.output and a.name == b.name
        assert [(x.name, list(x.ints), x.i, bytes(x.s)) for x in a.attribute] == \
               [(x.name, list(x.ints), x.i, bytes(x.s)) for x in b.attribute]
    assert [t.name for t in pm.graph.initializer] == [t.name for t in om.initializers]
    for a, b in zip(pm.graph.initializer, om.initializers):
        assert np.array_equal(P.tensor_to_numpy(a), b.array())
    assert [(v.name, P.value_info_dims(v)) for v in pm.graph.input] == [(v.name, list(v.dims)) for v in om.inputs]


def test_readers_agree_on_mnist():
    pm, om = P.load_model(MNIST_ONNX), ow.load_model(MNIST_ONNX)
    assert len(pm.graph.node) == 12 and pm.ir_version == 3
    _same_model(pm, om)
    # mnist-8 stores weights in float_data, not raw_data (utils.rs:134-137 path)
    assert all(len(t.float_data) or len(t.int64_data) for t in pm.graph.initializer)


def test_synth_model_roundtrip(tmp_path):
    data = synth.build_squeezenet(seed=0)
    assert data == synth.build_squeezenet(seed=0), "generator must be deterministic"
    pm, om = P.decode("ModelProto", data), ow.parse_model(data)
    assert len(pm.graph.node) == 66 and len(pm.graph.initializer) == 52
    _same_model(pm, om)
    assert abs(synth.conv_flops_per_image() - 1.637849152e9) < 1
    ops = [n.op_type for n in pm.graph.node]
    assert ops.count("Conv") == 26 and ops.count("Concat") == 8 and ops.count("MaxPool") == 3
    assert ops[-4:] == ["Conv", "Relu", "GlobalAveragePool", "Softmax"] and "Dropout" in ops
    # float_data storage variant decodes to the same arrays
    pm2 = P.decode("ModelProto", synth.build_squeezenet(seed=0, raw=False))
    for a, b in zip(pm.graph.initializer, pm2.graph.initializer):
        assert np.array_equal(P.tensor_to_numpy(a), P.tensor_to_numpy(b))


def test_tensor_pb_readers(tmp_path):
    for name in ("mnist_data_0.pb", "mnist_output_0.pb"):
        p = os.path.join(GOLDEN, name)
        a, b, c = P.read_tensor_pb(p), ow.load_tensor_pb(p), L.read_tensor_pb(p)   # c: the library's C++ reader
        assert np.array_equal(a, b) and np.array_equal(a.reshape(-1), np.asarray(c).reshape(-1))
    x = np.random.default_rng(0).standard_normal((1, 3, 5, 7)).astype(np.float32)
    q = str(tmp_path / "t.pb")
    P.write_tensor_pb(q, "data_0", x)
    assert np.array_equal(ow.load_tensor_pb(q), x) and np.array_equal(L.read_tensor_pb(q), x)


def test_library_rejects_garbage():
    buf = (C.c_char * 5).from_buffer_copy(b"\xff\xff\xff\xff\xff")
    n = C.c_size_t()
    rc = L.lib().b200_tensorproto_read(C.cast(buf, C.c_void_p), 5, None, 0, None, None, C.byref(n))
    assert rc == -5 and b"TensorProto" in L.lib().b200_last_error()
