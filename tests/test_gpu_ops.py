"""Per-operator parity on the GPU: every C-ABI operator entry point against the CPU oracle
(oracle/ref_ops.py, the restatement of src/inference_fp32_ops/*.rs) on seeded random inputs, the literal inputs of
the reference's dead test_* functions, and the reference's corner semantics.
Tolerance: 1e-4 relative + 1e-5 absolute (north_star); bit-exact where the op does no arithmetic."""
import numpy as np
import pytest

from conftest import assert_close

pytestmark = pytest.mark.gpu

PAD_VALID, PAD_SAME_UPPER, PAD_SAME_LOWER, PAD_NOTSET = 0, 1, 2, 3


def _oracle_conv(x, w, b, auto_pad, pads, strides):
    from oracle import ref_ops as R
    return np.stack([R.conv2d_image(x[i], w, b, auto_pad, pads, strides) for i in range(x.shape[0])])


def _oracle_pool(x, k, auto_pad, pads, strides):
    from oracle import ref_ops as R
    return np.stack([R.maxpool_image(x[i], k, auto_pad, pads, strides) for i in range(x.shape[0])])


CONV_CASES = [
    # N, C, H, W, M, kh, kw, strides, auto_pad, pads, bias
    (1, 1, 28, 28, 8, 5, 5, (1, 1), PAD_SAME_UPPER, (0, 0, 0, 0), False),   # MNIST Convolution28
    (2, 8, 14, 14, 16, 5, 5, (1, 1), PAD_SAME_UPPER, (0, 0, 0, 0), False),  # MNIST Convolution110
    (2, 3, 37, 41, 96, 7, 7, (2, 2), PAD_VALID, (0, 0, 0, 0), True),        # conv1-like, C=3 (zero-lane path)
    (3, 96, 9, 11, 16, 1, 1, (1, 1), PAD_VALID, (0, 0, 0, 0), True),        # squeeze 1x1
    (2, 16, 13, 13, 64, 3, 3, (1, 1), PAD_VALID, (1, 1, 1, 1), True),       # expand 3x3 pad 1 (pad promotion)
    (1, 64, 13, 13, 1000, 1, 1, (1, 1), PAD_VALID, (0, 0, 0, 0), True),     # conv10-like, ragged M
    (2, 5, 10, 9, 7, 3, 2, (2, 1), PAD_VALID, (1, 0, 2, 1), True),          # non-square kernel, asymmetric pads, odd C/M
    (1, 4, 10, 10, 6, 3, 3, (2, 2), PAD_SAME_UPPER, (0, 0, 0, 0), False),   # SAME with odd total pad (swapped split)
    (1, 2, 8, 7, 3, 4, 4, (1, 1), PAD_SAME_LOWER, (0, 0, 0, 0), True),      # SAME_LOWER, even kernel -> odd pad
    (1, 1, 5, 6, 1, 5, 2, (1, 1), PAD_VALID, (0, 0, 0, 0), False),          # convolution_op.rs:728 fixture shape
    (2, 3, 20, 17, 16, 3, 3, (2, 2), PAD_VALID, (1, 1, 1, 1), True),        # direct kernel, 16 filters, all four pads
    (3, 2, 9, 33, 12, 2, 5, (1, 2), PAD_VALID, (0, 2, 1, 2), True),         # direct kernel, M = 12 of 16 lanes used
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[f"conv{i}" for i in range(len(CONV_CASES))])
@pytest.mark.parametrize("relu", [False, True])
def test_conv2d(ctx, case, relu):
    from onnx_rusty_inference_engine_b200 import _lib as L
    N, C, H, W, M, kh, kw, strides, auto_pad, pads, has_bias = case
    rng = np.random.default_rng(hash(case) % (2**32))
    x = rng.standard_normal((N, C, H, W), dtype=np.float32) * 3
    w = rng.uniform(-1, 1, (M, C, kh, kw)).astype(np.float32) / np.sqrt(C * kh * kw)
    b = rng.uniform(-0.5, 0.5, (M,)).astype(np.float32) if has_bias else None
    eff_pad = PAD_NOTSET if any(p > 0 for p in pads) else auto_pad
    want = _oracle_conv(x, w, b, eff_pad, pads, strides)
    if relu:
        want = np.maximum(want, 0)
    tx, tw = ctx.tensor(x), ctx.tensor(w)
    tb = ctx.tensor(b) if has_bias else None
    y = L.conv2d(ctx, tx, tw, bias=tb, strides=strides, pads=pads, auto_pad=auto_pad, fuse_relu=relu)
    assert y.shape == want.shape
    assert_close(y.numpy(), want, f"conv {case}")


def test_conv2d_chan_add_and_concat_view(ctx):
    """Conv + folded Add([M,1,1]) + Relu written straight into a channel slice of a Concat result."""
    from onnx_rusty_inference_engine_b200 import _lib as L
    rng = np.random.default_rng(7)
    x = rng.standard_normal((2, 16, 9, 9), dtype=np.float32)
    w1 = rng.standard_normal((32, 16, 1, 1), dtype=np.float32) * 0.2
    w3 = rng.standard_normal((32, 16, 3, 3), dtype=np.float32) * 0.1
    add = rng.standard_normal((32, 1, 1), dtype=np.float32)
    want1 = np.maximum(_oracle_conv(x, w1, None, PAD_VALID, (0,) * 4, (1, 1)) + add[None], 0)
    want3 = np.maximum(_oracle_conv(x, w3, None, PAD_NOTSET, (1,) * 4, (1, 1)), 0)
    tx = ctx.tensor(x)
    out = L.DeviceTensor.alloc(ctx, (2, 64, 9, 9))
    v1, v3 = out.view_channels(0, 32), out.view_channels(32, 32)
    L.conv2d(ctx, tx, ctx.tensor(w1), chan_add=ctx.tensor(add), strides=(1, 1), fuse_relu=True, y=v1)
    L.conv2d(ctx, tx, ctx.tensor(w3), strides=(1, 1), pads=(1, 1, 1, 1), fuse_relu=True, y=v3)
    cat = L.concat(ctx, v1, v3, axis=1, y=out)  # both inputs already in place: zero-copy
    assert_close(cat.numpy(), np.concatenate([want1, want3], axis=1), "fire-style concat")


def test_conv2d_reference_fixtures(ctx):
    """Literal inputs of test_convolution_* (convolution_op.rs:728-841): (1,1,5,6) image * (1,1,5,2) kernel and
    (1,2,5,6) * (2,2,3,4); the reference only prints, the oracle supplies the expectation."""
    from onnx_rusty_inference_engine_b200 import _lib as L
    x1 = np.arange(1, 31, dtype=np.float32).reshape(1, 1, 5, 6)
    k1 = np.arange(1, 11, dtype=np.float32).reshape(1, 1, 5, 2)
    y = L.conv2d(ctx, ctx.tensor(x1), ctx.tensor(k1), strides=(1, 1))
    assert_close(y.numpy(), _oracle_conv(x1, k1, None, PAD_VALID, (0,) * 4, (1, 1)), "fixture 1")
    x2 = np.arange(1, 61, dtype=np.float32).reshape(1, 2, 5, 6)
    k2 = (np.arange(1, 49, dtype=np.float32) / 10).reshape(2, 2, 3, 4)
    y = L.conv2d(ctx, ctx.tensor(x2), ctx.tensor(k2), strides=(1, 1))
    assert_close(y.numpy(), _oracle_conv(x2, k2, None, PAD_VALID, (0,) * 4, (1, 1)), "fixture 2")


def test_conv2d_errors(ctx):
    from onnx_rusty_inference_engine_b200 import _lib as L
    x = ctx.tensor(np.zeros((1, 4, 8, 8), np.float32))
    w = ctx.tensor(np.zeros((4, 4, 3, 3), np.float32))
    with pytest.raises(L.B200Error, match="strides"):      # convolution_op.rs:285 unwrap
        L.conv2d(ctx, x, w)
    with pytest.raises(L.B200Error, match="group"):        # broken upstream, rejected
        L.conv2d(ctx, x, w, strides=(1, 1), group=2)
    with pytest.raises(L.B200Error, match="dilation"):
        L.conv2d(ctx, x, w, strides=(1, 1), dilations=(2, 2))
    w_bad = ctx.tensor(np.zeros((4, 3, 3, 3), np.float32))
    with pytest.raises(L.B200Error, match="C_in"):         # convolution_op.rs:252 assert
        L.conv2d(ctx, x, w_bad, strides=(1, 1))


POOL_CASES = [
    # N, C, H, W, k, strides, auto_pad, pads
    (2, 8, 28, 28, (2, 2), (2, 2), PAD_NOTSET, (0, 0, 0, 0)),   # MNIST Pooling66
    (2, 16, 14, 14, (3, 3), (3, 3), PAD_NOTSET, (0, 0, 0, 0)),  # MNIST Pooling160
    (2, 96, 21, 21, (3, 3), (2, 2), PAD_VALID, (0, 0, 0, 0)),   # SqueezeNet pool1-like
    (1, 64, 12, 12, (3, 3), (2, 2), PAD_NOTSET, (0, 0, 1, 1)),  # ceil-like pool after fire4
    (1, 64, 12, 12, (3, 3), (2, 2), PAD_VALID, (0, 0, 1, 1)),   # quirk: pads ignored without NOTSET (max_pool_op.rs:88)
    (1, 3, 9, 10, (2, 3), (1, 2), PAD_NOTSET, (1, 2, 0, 1)),    # odd C, asymmetric pads
    (1, 5, 7, 7, (3, 3), (2, 2), PAD_SAME_UPPER, (0, 0, 0, 0)),
    (2, 8, 33, 31, (3, 3), (2, 2), PAD_NOTSET, (1, 1, 1, 1)),   # strip kernel: two strips of output rows, all four pads
    (1, 12, 61, 9, (3, 3), (2, 2), PAD_VALID, (0, 0, 0, 0)),    # strip kernel: three strips, ragged last strip
]


@pytest.mark.parametrize("case", POOL_CASES, ids=[f"pool{i}" for i in range(len(POOL_CASES))])
@pytest.mark.parametrize("negative", [False, True])
def test_maxpool(ctx, case, negative):
    """negative=True feeds all-negative data: padded windows must then yield 0.0 (zero-fill, max_pool_op.rs:265-276)."""
    from onnx_rusty_inference_engine_b200 import _lib as L
    N, C, H, W, k, strides, auto_pad, pads = case
    rng = np.random.default_rng(hash(case) % (2**32))
    x = rng.standard_normal((N, C, H, W), dtype=np.float32)
    if negative:
        x = -np.abs(x) - 0.5
    want = _oracle_pool(x, k, auto_pad, pads, strides)
    y = L.maxpool2d(ctx, ctx.tensor(x), kernel=k, strides=strides, pads=pads, auto_pad=auto_pad)
    got = y.numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want), f"maxpool {case}: not bit-exact"


def test_relu_reference_fixture(ctx):
    """relu_op.rs:35-51: 35 values with one negative; expected tensor given upstream (:43-47)."""
    from onnx_rusty_inference_engine_b200 import _lib as L
    x = np.arange(1, 36, dtype=np.float32).reshape(1, 1, 5, 7).copy()
    x[0, 0, 0, 0] = -1.0
    got = L.relu(ctx, ctx.tensor(x)).numpy()
    assert np.array_equal(got, np.maximum(x, 0))
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 24, 5, 7), dtype=np.float32)
    assert np.array_equal(L.relu(ctx, ctx.tensor(x)).numpy(), np.maximum(x, 0))


def test_add(ctx):
    from onnx_rusty_inference_engine_b200 import _lib as L
    rng = np.random.default_rng(4)
    x = rng.standard_normal((2, 8, 28, 28), dtype=np.float32)
    b = rng.standard_normal((8, 1, 1), dtype=np.float32)
    assert np.array_equal(L.add(ctx, ctx.tensor(x), ctx.tensor(b)).numpy(), x + b[None])  # add_op.rs:75
    x2 = rng.standard_normal((1, 10), dtype=np.float32)
    b2 = rng.standard_normal((1, 10), dtype=np.float32)
    assert np.array_equal(L.add(ctx, ctx.tensor(x2), ctx.tensor(b2)).numpy(), x2 + b2)      # add_op.rs:84
    x3 = rng.standard_normal((5, 10), dtype=np.float32)
    assert np.array_equal(L.add(ctx, ctx.tensor(x3), ctx.tensor(b2)).numpy(), x3 + b2)      # batch-N extension
    with pytest.raises(L.B200Error):
        L.add(ctx, ctx.tensor(x), ctx.tensor(b2))


def test_matmul(ctx):
    from onnx_rusty_inference_engine_b200 import _lib as L
    from oracle import ref_ops as R
    import ctypes
    rng = np.random.default_rng(5)
    # (r, 256, 10): MNIST's Times212 -- tcgen05 path (rows as pixels, one channel tile); (300, 512, 200): 3 row tiles x 2
    # channel tiles x 16 k-blocks on tcgen05; K = 4 and K = 37 are not 16-byte rows: CUDA-core kernel
    for (r, k, n) in [(1, 256, 10), (7, 256, 10), (300, 512, 200), (3, 4, 3), (130, 37, 65)]:
        a = rng.standard_normal((r, k), dtype=np.float32)
        b = rng.standard_normal((k, n), dtype=np.float32)
        if k > 256:   # keep the partial sums O(1): with unit-variance operands and K = 512 the fp32 summation-order noise
            b /= np.float32(np.sqrt(k))   # on near-zero results alone exceeds 1e-4 relative (any two fp32 kernels disagree)
        bias = rng.standard_normal((1, n), dtype=np.float32)
        want = np.empty((r, n), np.float32)
        R.lib().ref_matmul(R._p(a), R._p(b), r, k, n, R._p(want))
        assert_close(L.matmul(ctx, ctx.tensor(a), ctx.tensor(b)).numpy(), want, f"matmul {r}x{k}x{n}")
        assert_close(L.matmul(ctx, ctx.tensor(a), ctx.tensor(b), bias=ctx.tensor(bias)).numpy(), want + bias,
                     f"matmul+bias {r}x{k}x{n}")


def test_matmul_weight_cache_follows_uploads(ctx):
    """The split / permuted tcgen05 weight tiles are cached on the right operand: a new upload must invalidate them."""
    from onnx_rusty_inference_engine_b200 import _lib as L
    rng = np.random.default_rng(8)
    a = rng.standard_normal((5, 64), dtype=np.float32)
    tb = ctx.tensor(rng.standard_normal((64, 16), dtype=np.float32))
    ta = ctx.tensor(a)
    L.matmul(ctx, ta, tb)
    b2 = rng.standard_normal((64, 16), dtype=np.float32)
    tb.upload(b2)
    assert_close(L.matmul(ctx, ta, tb).numpy(), (a.astype(np.float64) @ b2.astype(np.float64)), "matmul after re-upload")


def test_upload_into_channel_view_leaves_siblings_alone(ctx):
    """A channel view has the parent's pitch: uploading into it must write exactly its own lanes (the staged upload used
    to zero-fill [0, ld) from the view's base: sibling channels zeroed, and an out-of-bounds write on the last pixel)."""
    from onnx_rusty_inference_engine_b200 import _lib as L
    rng = np.random.default_rng(9)
    for (C, off, ln) in [(12, 4, 4), (10, 3, 5), (8, 0, 6)]:
        full = rng.standard_normal((2, C, 5, 7), dtype=np.float32)
        parent = ctx.tensor(full)
        part = rng.standard_normal((2, ln, 5, 7), dtype=np.float32)
        parent.view_channels(off, ln).upload(part)
        want = full.copy(); want[:, off:off + ln] = part
        assert np.array_equal(parent.numpy(), want), (C, off, ln)
        assert np.array_equal(parent.view_channels(off, ln).numpy(), part)


def test_context_close_before_tensor_gc():
    """Tensors keep their context alive: closing the context first and freeing the tensor later is not a use-after-free."""
    from onnx_rusty_inference_engine_b200 import _lib as L
    c = L.Context(0)
    t = c.tensor(np.ones((1, 4, 2, 2), np.float32))
    c.close()
    t.free()


def test_reshape_concat_dropout(ctx):
    from onnx_rusty_inference_engine_b200 import _lib as L
    rng = np.random.default_rng(6)
    x = rng.standard_normal((1, 16, 4, 4), dtype=np.float32)
    assert np.array_equal(L.reshape(ctx, ctx.tensor(x), [1, 256]).numpy(), x.reshape(1, 256))      # NCHW order
    p = rng.standard_normal((16, 4, 4, 10), dtype=np.float32)
    assert np.array_equal(L.reshape(ctx, ctx.tensor(p), [256, 10]).numpy(), p.reshape(256, 10))    # Parameter193
    assert np.array_equal(L.reshape(ctx, ctx.tensor(x), [0, 256]).numpy(), x.reshape(1, 256))      # 0 copies the dim
    with pytest.raises(L.B200Error):
        L.reshape(ctx, ctx.tensor(x), [1, -1])
    a = rng.standard_normal((2, 6, 5, 5), dtype=np.float32)
    b = rng.standard_normal((2, 3, 5, 5), dtype=np.float32)
    assert np.array_equal(L.concat(ctx, ctx.tensor(a), ctx.tensor(b), axis=1).numpy(), np.concatenate([a, b], 1))
    assert np.array_equal(L.dropout(ctx, ctx.tensor(a), 0.5).numpy(), a)                           # dropout_op.rs:66-71


@pytest.mark.parametrize("axis", [0, 1, 2, 3])
def test_concat_any_axis(ctx, axis):
    """concatenate_op.rs:31 takes any axis (ndarray::concatenate(Axis(axis)); both bundled models use 1)."""
    from onnx_rusty_inference_engine_b200 import _lib as L
    rng = np.random.default_rng(40 + axis)
    sa, sb = [2, 6, 5, 7], [2, 6, 5, 7]
    sb[axis] = 3
    a, b = rng.standard_normal(sa, dtype=np.float32), rng.standard_normal(sb, dtype=np.float32)
    got = L.concat(ctx, ctx.tensor(a), ctx.tensor(b), axis=axis).numpy()
    assert np.array_equal(got, np.concatenate([a, b], axis=axis))
    with pytest.raises(L.B200Error):
        L.concat(ctx, ctx.tensor(a), ctx.tensor(b), axis=4)
    if axis != 3:
        bad = list(sa); bad[3] = 2
        with pytest.raises(L.B200Error, match="non-axis dims differ"):
            L.concat(ctx, ctx.tensor(a), ctx.tensor(rng.standard_normal(bad, dtype=np.float32)), axis=axis)


def test_gap_softmax(ctx):
    from onnx_rusty_inference_engine_b200 import _lib as L
    from oracle import ref_ops as R
    rng = np.random.default_rng(8)
    x = rng.standard_normal((3, 1000, 13, 13), dtype=np.float32) * 4
    want = np.stack([x[i].reshape(1000, -1).astype(np.float32).sum(axis=1, dtype=np.float32) for i in range(3)])
    g = L.global_avgpool(ctx, ctx.tensor(x))
    assert g.shape == (3, 1000, 1, 1)
    gw = np.empty((3, 1000), np.float32)
    for i in range(3):
        R.lib().ref_global_avgpool(R._p(np.ascontiguousarray(x[i])), 1000, 169, R._p(gw[i]))
    assert_close(g.numpy().reshape(3, 1000), gw, "global_average_pool")
    # global_average_pool_op.rs:54-65 fixture: (1,2,4,4) of 1..16 twice
    f = np.tile(np.arange(1, 17, dtype=np.float32).reshape(1, 1, 4, 4), (1, 2, 1, 1))
    assert_close(L.global_avgpool(ctx, ctx.tensor(f)).numpy().reshape(-1), np.array([8.5, 8.5], np.float32), "gap fixture")
    # softmax over (C*H*W): reference fixture softmax_op.rs:59-67 is an overflow-stability case
    s = np.array([118.85734, 5640.1426, 2, 3, 1000, 1001, 1002, 1003], np.float32).reshape(1, 8, 1, 1)
    sw = np.empty((1, 8), np.float32)
    R.lib().ref_softmax_row(R._p(s.reshape(-1)), 8, R._p(sw[0]))
    got = L.softmax(ctx, ctx.tensor(s)).numpy()
    assert np.isfinite(got).all()
    assert_close(got, sw, "softmax fixture")
    y = rng.standard_normal((4, 10, 3, 2), dtype=np.float32) * 3   # H*W > 1: flatten in NCHW order
    yw = np.empty((4, 60), np.float32)
    for i in range(4):
        R.lib().ref_softmax_row(R._p(np.ascontiguousarray(y[i]).reshape(-1)), 60, R._p(yw[i]))
    assert_close(L.softmax(ctx, ctx.tensor(y)).numpy(), yw, "softmax HW>1")
