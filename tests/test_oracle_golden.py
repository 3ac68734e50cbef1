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


def _torch_onnx_forward(model, x):
    """A second, independent evaluation of an ONNX graph: ONNX operator semantics on torch.nn.functional in fp64.  Only what the
    synthetic SqueezeNet / Fire / pool models use; padded MaxPool cells are ZEROS as in the reference (max_pool_op.rs), which
    equals ONNX's -inf padding wherever the pooled tensor is non-negative (always the case behind a Relu)."""
    env = {t.name: torch.from_numpy(np.asarray(t.array(), dtype=np.float64)) for t in model.initializers}
    data_in = [v.name for v in model.inputs if v.name not in env]
    assert len(data_in) == 1
    env[data_in[0]] = torch.from_numpy(x.astype(np.float64))
    for n in model.nodes:
        a = {at.name: at for at in n.attribute}
        ins = [env[i] for i in n.input]
        if n.op_type == "Conv":
            pads = list(a["pads"].ints) if "pads" in a else [0, 0, 0, 0]
            xp = F.pad(ins[0], (pads[1], pads[3], pads[0], pads[2]))
            y = F.conv2d(xp, ins[1], ins[2] if len(ins) > 2 else None, stride=tuple(a["strides"].ints))
        elif n.op_type == "Relu":
            y = torch.relu(ins[0])
        elif n.op_type == "MaxPool":
            pads = list(a["pads"].ints) if "pads" in a else [0, 0, 0, 0]
            xp = F.pad(ins[0], (pads[1], pads[3], pads[0], pads[2]), value=0.0)
            y = F.max_pool2d(xp, tuple(a["kernel_shape"].ints), tuple(a["strides"].ints))
        elif n.op_type == "Concat":
            y = torch.cat(ins, dim=a["axis"].i)
        elif n.op_type == "Dropout":
            y = ins[0]
        elif n.op_type == "GlobalAveragePool":
            y = ins[0].mean(dim=(2, 3), keepdim=True)
        elif n.op_type == "Softmax":
            ax = a["axis"].i if "axis" in a else 1
            y = torch.softmax(ins[0].flatten(ax), dim=1).reshape(ins[0].shape)
        else:
            raise AssertionError(f"operator {n.op_type} not in the cross-check interpreter")
        env[n.output[0]] = y
    return env[model.outputs[0].name].numpy()


def test_oracle_squeezenet_ops_against_an_independent_evaluation(tmp_path):
    """The reference ships no SqueezeNet blob (.MISSING_LARGE_BLOBS), so Concat / Dropout / GlobalAveragePool / Softmax / the
    strided 7x7 stem / the padded MaxPool are pinned by the restated oracle only.  This pins the restatement itself from a
    second side: the whole synthetic SqueezeNet (66 nodes) and a Fire module through the oracle against ONNX semantics on
    torch.nn.functional in fp64 (the synthetic models are built so that the reference's semantics and ONNX's coincide,
    synth.py).  Not a substitute for the missing golden pair: a restatement error shared with ONNX semantics would pass."""
    from onnx_rusty_inference_engine_b200 import synth
    path = synth.ensure_squeezenet(str(tmp_path / "squeezenet_synth.onnx"), seed=0)
    m = ow.load_model(path)
    xs = synth.synthetic_batch(2, seed=11)
    got = rm.run_batch(m, xs, threads=2).reshape(2, -1)
    want = np.stack([_torch_onnx_forward(m, xs[i:i + 1]).reshape(-1) for i in range(2)])
    assert got.shape == want.shape == (2, 1000)
    assert np.array_equal(got.argmax(1), want.argmax(1))
    assert_close(got, want.astype(np.float32), "oracle vs independent fp64 evaluation (SqueezeNet synth)")
    assert abs(float(got.sum(1).max()) - 1.0) < 1e-4
    fpath = str(tmp_path / "fire.onnx")
    with open(fpath, "wb") as f:
        f.write(synth.build_fire(16, 16, 64, 13, seed=3))
    fm = ow.load_model(fpath)
    fx = synth.synthetic_batch(2, chw=(16, 13, 13), seed=5, std=1.0)
    fgot = rm.run_batch(fm, fx, threads=2)
    fwant = np.stack([_torch_onnx_forward(fm, fx[i:i + 1])[0] for i in range(2)])
    assert_close(fgot.reshape(fwant.shape), fwant.astype(np.float32), "oracle vs independent fp64 evaluation (Fire module)")
    ppath = str(tmp_path / "pool.onnx")
    with open(ppath, "wb") as f:
        f.write(synth.build_pool_squeeze(8, 4, 21, pads=(0, 0, 1, 1), seed=2))
    pm = ow.load_model(ppath)
    px = np.abs(synth.synthetic_batch(2, chw=(8, 21, 21), seed=6, std=1.0))   # non-negative: zero padding == ONNX padding
    pgot = rm.run_batch(pm, px, threads=2)
    pwant = np.stack([_torch_onnx_forward(pm, px[i:i + 1])[0] for i in range(2)])
    assert_close(pgot.reshape(pwant.shape), pwant.astype(np.float32), "oracle vs independent fp64 evaluation (padded MaxPool)")
