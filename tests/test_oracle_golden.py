"""CPU tests of the oracle itself (no GPU): pinned against the reference's bundled golden pair and cross-checked
against torch.nn.functional, including the reference's corner semantics."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, MNIST_ONNX, assert_close
from oracle import onnx_wire as ow
from oracle import ref_model as rm
from oracle import ref_ops as R


def test_mnist_golden_pins_the_oracle():
    """mnist_data_0.pb -> mnist_output_0.pb through models/mnist-8.onnx: Conv (C=1 and C>1, SAME_UPPER 5x5),
    Add (4-D + [C,1,1], 2-D + 2-D), Relu, MaxPool (2x2 s2, 3x3 s3, NOTSET), Reshape (both forms), MatMul."""
    m = ow.load_model(MNIST_ONNX)
    x = ow.load_tensor_pb(os.path.join(GOLDEN, "mnist_data_0.pb"))
    want = ow.load_tensor_pb(os.path.join(GOLDEN, "mnist_output_0.pb")).reshape(1, 10)
    got = rm.inference(m, x, ["Input3", "Parameter193"])   # names as passed in main.rs:14
    assert_close(got, want, "oracle vs mnist golden")
    assert int(got.argmax()) == int(want.argmax()) == 2     # reference prints "Class 3-nth" (1-based)


@pytest.mark.parametrize("C,H,W,M,kh,kw,s,pads", [
    (1, 12, 12, 4, 5, 5, (1, 1), (0, 0, 0, 0)),
    (3, 21, 23, 8, 7, 7, (2, 2), (0, 0, 0, 0)),
    (8, 9, 9, 16, 3, 3, (1, 1), (1, 1, 1, 1)),
    (5, 10, 9, 7, 3, 2, (2, 1), (1, 0, 2, 1)),
    (16, 6, 6, 32, 1, 1, (1, 1), (0, 0, 0, 0)),
])
def test_oracle_conv_is_cross_correlation(C, H, W, M, kh, kw, s, pads):
    rng = np.random.default_rng(C * 100 + M)
    x = rng.standard_normal((C, H, W), dtype=np.float32)
    w = rng.standard_normal((M, C, kh, kw), dtype=np.float32)
    b = rng.standard_normal((M,), dtype=np.float32)
    ap = R.PAD_NOTSET if any(pads) else R.PAD_VALID
    got = R.conv2d_image(x, w, b, ap, pads, s)
    xp = F.pad(torch.from_numpy(x)[None].double(), (pads[1], pads[3], pads[0], pads[2]))
    want = F.conv2d(xp, torch.from_numpy(w).double(), torch.from_numpy(b).double(), stride=s)[0].numpy()
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-4, atol=1e-4)


def test_oracle_same_upper_puts_odd_pad_top_left():
    """get_padding_size returns top/bottom and left/right swapped (convolution_op.rs:547-556)."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 10, 10), dtype=np.float32)
    w = rng.standard_normal((3, 2, 3, 3), dtype=np.float32)
    got = R.conv2d_image(x, w, None, R.PAD_SAME_UPPER, [], (2, 2))       # total pad 1 per axis
    xp = F.pad(torch.from_numpy(x)[None], (1, 0, 1, 0))                  # extra pad at left / top
    want = F.conv2d(xp, torch.from_numpy(w), stride=2)[0].numpy()
    assert got.shape == (3, 5, 5)
    assert np.allclose(got, want, rtol=1e-4, atol=1e-5)
    assert np.array_equal(got, R.conv2d_image(x, w, None, R.PAD_SAME_LOWER, [], (2, 2)))


def test_oracle_maxpool_quirks():
    rng = np.random.default_rng(1)
    x = np.abs(rng.standard_normal((4, 12, 12), dtype=np.float32))
    got = R.maxpool_image(x, (3, 3), R.PAD_NOTSET, (0, 0, 1, 1), (2, 2))
    want = F.max_pool2d(F.pad(torch.from_numpy(x)[None], (0, 1, 0, 1)), 3, 2)[0].numpy()
    assert got.shape == (4, 6, 6) and np.array_equal(got, want)
    # pads ignored without auto_pad=NOTSET (max_pool_op.rs:88,188): 12 -> 5, not 6
    assert R.maxpool_image(x, (3, 3), R.PAD_VALID, (0, 0, 1, 1), (2, 2)).shape == (4, 5, 5)
    # zero-fill, not -inf: all-negative input gives 0.0 where a window touches the padding (max_pool_op.rs:265-276)
    neg = -x - 1
    g = R.maxpool_image(neg, (3, 3), R.PAD_NOTSET, (0, 0, 1, 1), (2, 2))
    assert (g[:, -1, :] == 0).all() and (g[:, :, -1] == 0).all() and (g[:, :-1, :-1] < 0).all()


def test_oracle_store_semantics_and_panics():
    m = ow.load_model(MNIST_ONNX)
    x = ow.load_tensor_pb(os.path.join(GOLDEN, "mnist_data_0.pb"))
    out, store = rm.inference(m, x, ["Input3"], return_store=True)
    assert store["Parameter193_reshape1"][0].shape == (256, 10) and store["Parameter193_reshape1"][1] is None
    assert store["Pooling160_Output_0"][1].shape == (1, 16, 4, 4)
    with pytest.raises(R.RefPanic):   # input length must match the static model shape (utils.rs:40)
        rm.inference(m, np.zeros(2 * 784, np.float32), ["Input3"])
    bad = ow.Node(op_type="Sigmoid", name="s", input=["Input3"], output=["y"])
    with pytest.raises(R.RefPanic, match="NOT FOUND"):
        rm.node_inference(bad, {}, m)
    conv = m.nodes[1]
    conv2 = ow.Node(op_type="Conv", name="c", input=conv.input, output=conv.output,
                    attribute=conv.attribute + [ow.Attribute(name="foo", i=1)])
    st = {"Input3": (None, np.zeros((1, 1, 28, 28), np.float32))}
    with pytest.raises(R.RefPanic, match="ATTRIBUTE NAME FOR CONVOLUTION"):
        rm.node_inference(conv2, st, m)


def test_oracle_softmax_gap_fixtures():
    s = np.array([118.85734, 5640.1426, 2, 3, 1000, 1001, 1002, 1003], np.float32)   # softmax_op.rs:59-67
    out = np.empty_like(s)
    R.lib().ref_softmax_row(R._p(s), 8, R._p(out))
    assert np.isfinite(out).all() and abs(out.sum() - 1) < 1e-6 and out.argmax() == 1
    f = np.tile(np.arange(1, 17, dtype=np.float32).reshape(1, 4, 4), (2, 1, 1))          # global_average_pool_op.rs:54-65
    g = np.empty(2, np.float32)
    R.lib().ref_global_avgpool(R._p(np.ascontiguousarray(f)), 2, 16, R._p(g))
    assert np.array_equal(g, np.array([8.5, 8.5], np.float32))
