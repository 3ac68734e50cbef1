"""End-to-end parity on the GPU (BASELINE.json configs 1, 3, 4, 5 at test sizes):
MNIST-8 against the reference's bundled golden pair, batch-N against N independent batch-1 oracle runs, synthetic
SqueezeNet1.0 against the oracle, and size-independent properties at the full batch sizes (batch-position
invariance: image i of a batch-256 run is bitwise the batch-1 result; softmax rows sum to 1)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, MNIST_ONNX, assert_close, assert_close_mnist_noise

pytestmark = pytest.mark.gpu


def _golden():
    from oracle import onnx_wire as ow
    x = ow.load_tensor_pb(os.path.join(GOLDEN, "mnist_data_0.pb")).astype(np.float32)
    y = ow.load_tensor_pb(os.path.join(GOLDEN, "mnist_output_0.pb")).reshape(1, -1)
    return x, y


def test_mnist_golden_engine(ctx):
    """Config 1 through the graph-level path (b200_model_run)."""
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    x, want = _golden()
    eng = Engine(MNIST_ONNX, ctx=ctx)
    assert eng.in_chw == (1, 28, 28) and eng.out_per_image == 10
    got = eng(x.reshape(1, 1, 28, 28))
    assert_close(got, want, "mnist golden (engine)")
    assert int(got.argmax()) == int(want.argmax()) == 2
    eng.model.set_option("cuda_graph", 0)
    assert np.array_equal(eng(x.reshape(1, 1, 28, 28)), got), "graph replay and direct launches must agree bitwise"


def test_mnist_golden_node_walk(ctx):
    """Config 1 through the reference-shaped per-node interface: inference() -> node_inference() -> op functions."""
    from onnx_rusty_inference_engine_b200 import onnx_proto as P
    from onnx_rusty_inference_engine_b200.inference_engine import inference
    x, want = _golden()
    model = P.load_model(MNIST_ONNX)
    got = inference(model, x, ["Input3", "Parameter193"], ctx=ctx)   # names as in main.rs:14
    assert_close(got, want, "mnist golden (node walk)")
    assert int(got.argmax()) == 2


def test_group17_entry_point(capsys):
    from onnx_rusty_inference_engine_b200 import group17
    x, want = _golden()
    got = group17.onnx_make_inference(MNIST_ONNX, os.path.join(GOLDEN, "mnist_data_0.pb"),
                                      os.path.join(GOLDEN, "mnist_output_0.pb"), ["Input3", "Parameter193"])
    assert_close(got, want, "group17")
    out = capsys.readouterr().out
    assert "MNist-8 Inference results: Class 3-nth predicted." in out   # 1-based class, add_op.rs:100-104
    assert "Expected Data:" in out                                       # main.rs:41


def test_cli_binaries_mirror_main_rs():
    """main.rs:9-53 as a compiled binary over the C ABI and as `python -m`: the reference's result lines, exit 0 on a match."""
    import subprocess
    import sys
    from conftest import ROOT
    args = [MNIST_ONNX, os.path.join(GOLDEN, "mnist_data_0.pb"), os.path.join(GOLDEN, "mnist_output_0.pb"), "Input3", "Parameter193"]
    exe = os.path.join(ROOT, "onnx_rusty_inference_engine_b200", "lib", "onnx_rusty_inference_engine_bin")
    for cmd in ([exe] + args, [sys.executable, "-m", "onnx_rusty_inference_engine_b200"] + args):
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        assert "MNist-8 Inference results: Class 3-nth predicted." in r.stdout
        assert "Expected Data:" in r.stdout and "Match (1e-4 rel + 1e-5 abs, argmax): yes" in r.stdout


def test_mnist_batch_vs_oracle(ctx):
    """Config 5 at test size: batch 64 of N(0,10^2) images == 64 batch-1 oracle runs; batch-position invariant."""
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    from oracle import onnx_wire as ow, ref_model as rm
    xs = synth.synthetic_batch(64, chw=(1, 28, 28), seed=3)
    want = rm.run_batch(ow.load_model(MNIST_ONNX), xs, ["Input3", "Parameter193"], threads=4)
    eng = Engine(MNIST_ONNX, ctx=ctx)
    got = eng(xs)
    assert_close_mnist_noise(got, want, "mnist batch 64")
    one = eng(xs[17:18])
    assert np.array_equal(one[0], got[17]), "batch-position invariance (bitwise)"
    big = eng(np.tile(xs, (64, 1, 1, 1)))          # 4096 images: same 64 answers, 64 times
    assert np.array_equal(big.reshape(64, 64, 10), np.broadcast_to(got, (64, 64, 10)))


@pytest.mark.parametrize("batch", [1, 7, 8, 9, 300])
def test_mnist_fused_path_vs_oracle_and_node_plan(ctx, batch):
    """Config 5's fused path (mnist8_fused.cu: the 12-node graph as two launches, conv2 + pool + MatMul on tcgen05) against
    the oracle and against the node-by-node plan of the same model, on ragged batch sizes (groups of 8 images: a partial
    last group, a single image); bitwise batch-position invariance."""
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    from oracle import onnx_wire as ow, ref_model as rm
    xs = synth.synthetic_batch(batch, chw=(1, 28, 28), seed=100 + batch)
    want = rm.run_batch(ow.load_model(MNIST_ONNX), xs, ["Input3", "Parameter193"], threads=8)
    eng = Engine(MNIST_ONNX, ctx=ctx)
    assert eng.model.launches_per_run(batch) == 1, "the fused path should be one launch"
    got = eng(xs)
    assert_close_mnist_noise(got, want, f"mnist fused, batch {batch}")
    unit = synth.synthetic_batch(batch, chw=(1, 28, 28), seed=200 + batch, std=1.0)   # unit-scale input: plain tolerance
    assert_close(eng(unit), rm.run_batch(ow.load_model(MNIST_ONNX), unit, ["Input3", "Parameter193"], threads=8), f"mnist fused, unit-scale input, batch {batch}")
    assert np.array_equal(eng(xs), got), "run-to-run determinism"
    k = batch // 2
    assert np.array_equal(eng(xs[k:k + 1])[0], got[k]), "batch-position invariance (bitwise)"
    eng.model.set_option("fused_cnn", 1)                      # the two-launch form (pooled stem output through HBM)
    assert eng.model.launches_per_run(batch) == 2
    two = eng(xs)
    assert np.array_equal(two, got), "one-launch and two-launch forms run the same arithmetic"
    eng.model.set_option("fused_cnn", 0)
    assert eng.model.launches_per_run(batch) == 5
    assert_close_mnist_noise(eng(xs), want, f"mnist node-by-node plan, batch {batch}")


def test_mnist_config5_full_size(ctx):
    """Config 5 at BASELINE.json's full size on one device: 65,536 N(0,10^2) images (seed 3) through the fused path.  Images
    sampled from the run -- first / last group, both sides of the 148-CTA round-robin boundary, a ragged position -- against
    the oracle; bitwise equality of sampled groups with small-batch runs of the same images (batch-position invariance at
    full size); run-to-run determinism; and the pipelined host entry with two batch sizes interleaved."""
    import torch
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    from oracle import onnx_wire as ow, ref_model as rm
    n = 65536
    xs = synth.synthetic_batch(n, chw=(1, 28, 28), seed=3)
    eng = Engine(MNIST_ONNX, ctx=ctx)
    assert eng.model.launches_per_run(n) == 1
    got = eng(xs)
    assert got.shape == (n, 10) and np.isfinite(got).all()
    idx = [0, 7, 8, 1183, 1184, 9471, 32768, 40001, 65528, 65535, 12345, 54321, 148 * 8 - 1, 148 * 8, 2 * 148 * 8 + 3, 60000]
    want = rm.run_batch(ow.load_model(MNIST_ONNX), xs[idx], ["Input3", "Parameter193"], threads=8)
    assert_close_mnist_noise(got[idx], want, "mnist batch 65536, sampled images vs oracle")
    for g0 in (0, 1184, 40000, 65528):
        assert np.array_equal(eng(xs[g0:g0 + 8]), got[g0:g0 + 8]), f"group at {g0}: batch-position invariance"
    assert np.array_equal(eng(xs[65535:65536])[0], got[65535])
    assert np.array_equal(eng(xs), got), "run-to-run determinism at full size"
    a = torch.from_numpy(xs[:4096]).pin_memory(); b = torch.from_numpy(xs[4096:4096 + 1000]).pin_memory()
    oa = torch.empty((4096, 10)).pin_memory(); ob = torch.empty((1000, 10)).pin_memory()
    for _ in range(3):
        eng.run_pinned_async(a, oa)
        eng.run_pinned_async(b, ob)
    eng.sync()
    assert np.array_equal(oa.numpy(), got[:4096]) and np.array_equal(ob.numpy(), got[4096:5096])


def test_squeezenet_synth_vs_oracle(ctx, synth_onnx):
    """Config 3 at test size: 6 seeded images through all 66 nodes vs the oracle, both executors."""
    from onnx_rusty_inference_engine_b200 import onnx_proto as P, synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine, inference
    from oracle import onnx_wire as ow, ref_model as rm
    xs = synth.synthetic_batch(6, seed=1)
    want = rm.run_batch(ow.load_model(synth_onnx), xs, threads=6)
    eng = Engine(synth_onnx, ctx=ctx)
    got = eng(xs)
    assert got.shape == (6, 1000)
    assert_close(got, want, "squeezenet synth (engine)")
    assert (got.argmax(1) == want.argmax(1)).all()
    walk = inference(P.load_model(synth_onnx), xs[0], ["data_0"], ctx=ctx)
    assert_close(walk, want[0:1], "squeezenet synth (node walk)")
    eng.model.set_option("conv_path", 1)   # CUDA-core cross-check path must agree with the default path
    got_simt = eng(xs)
    assert_close(got_simt, want, "squeezenet synth (conv_path=1)")


def test_fire_fusion_matches_separate_launches(ctx, synth_onnx):
    """fire2 / fire3: expand1x1 + expand3x3 as one convolution (1x1 filters at the centre tap, exact zeros elsewhere)
    must give what the two separate launches give, and both must match the oracle."""
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    from oracle import onnx_wire as ow, ref_model as rm
    xs = synth.synthetic_batch(3, seed=7)
    eng = Engine(synth_onnx, ctx=ctx)
    fused = eng(xs)
    n_fused = eng.model.launches_per_run(3)
    eng.model.set_option("fire_fusion", 0)
    separate = eng(xs)
    n_sep = eng.model.launches_per_run(3)
    assert n_sep == n_fused + 2, (n_sep, n_fused)          # fire2 and fire3 (M1 + M3 = 128) fuse; fire4.. do not
    assert_close(fused, separate, "fire fusion vs separate launches")
    want = rm.run_batch(ow.load_model(synth_onnx), xs[:2], threads=2)
    assert_close(fused[:2], want, "fire fusion vs oracle")


@pytest.mark.parametrize("cin,squeeze,expand,hw,batch,fuses", [
    (8, 16, 64, 17, 4, True),       # fire2-like: both branches in one channel tile -> one launch
    (24, 32, 128, 11, 5, False),    # fire4-like: K = 288 (merged accumulator), ragged last pixel tile
    (16, 48, 192, 9, 7, False),     # fire6-like: channel tiles of 96
    (16, 64, 256, 13, 3, False),    # fire8-like: two channel tiles of 128, 18 k-blocks
])
def test_fire_module_alone(ctx, tmp_path, cin, squeeze, expand, hw, batch, fuses):
    """A Fire module as a model of its own, so the Concat result itself (channel-slice writes of both branches) is compared
    with the oracle element by element, with and without the expand fusion."""
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    from oracle import onnx_wire as ow, ref_model as rm
    path = str(tmp_path / "fire.onnx")
    with open(path, "wb") as f:
        f.write(synth.build_fire(cin, squeeze, expand, hw, seed=cin + expand))
    # unit-variance input: activations stay O(1), the scale at which the absolute 1e-5 term of the tolerance is meant
    xs = synth.synthetic_batch(batch, chw=(cin, hw, hw), seed=expand, std=1.0)
    want = rm.run_batch(ow.load_model(path), xs, threads=2)
    eng = Engine(path, ctx=ctx)
    assert eng.out_per_image == 2 * expand * hw * hw
    fused = eng(xs)
    n_fused = eng.model.launches_per_run(batch)
    assert_close(fused.reshape(want.shape), want, "fire module vs oracle")
    eng.model.set_option("fire_fusion", 0)
    separate = eng(xs)
    assert eng.model.launches_per_run(batch) == n_fused + (1 if fuses else 0)
    assert_close(separate.reshape(want.shape), want, "separate fire launches vs oracle")
    eng.model.set_option("fire_fusion", 1)
    eng.model.set_option("alt_order", 0)
    assert np.array_equal(eng(xs), fused), "tile walking direction must not change the bits"


@pytest.mark.parametrize("cin,squeeze,hw,pads,batch,tail,neg", [
    (32, 16, 55, (0, 0, 0, 0), 3, False, False),    # one k-block, 27 x 27 map: tiles of 4 pooled rows, the last of an image 3
    (96, 16, 109, (0, 0, 0, 0), 2, False, False),   # pool1 -> fire2 squeeze: three k-blocks, tiles of 2 pooled rows of 54, 109-pixel window boxes
    (64, 32, 54, (0, 0, 1, 1), 5, True, True),      # the pool after fire4: zero padding at the end, all-negative input (max with 0)
    (48, 64, 57, (1, 1, 0, 0), 2, True, True),      # padding in front, channels end inside the second k-block, 64 filters
    (256, 32, 54, (0, 0, 1, 1), 1, False, False),   # fire5 squeeze: eight k-blocks (raw ring wraps inside a tile)
    (512, 64, 27, (0, 0, 0, 0), 3, False, False),   # the pool after fire8 -> fire9 squeeze: 13 x 13 map, tiles of 7 + 6 pooled rows, 16 k-blocks
    (16, 8, 21, (0, 0, 0, 0), 5, True, True),       # 10 x 10 map in one tile, channels end inside the only k-block
    (32, 16, 201, (0, 0, 0, 0), 1, False, False),   # 100 pooled pixels per row: every lane quarter of the tile in use, 78 KB raw slots
    (32, 16, 261, (0, 0, 0, 0), 1, False, None),    # 130 pooled pixels per row: not eligible, stays MaxPool + Conv
])
def test_pool_fusion_matches_separate_launches(ctx, tmp_path, cin, squeeze, hw, pads, batch, tail, neg):
    """MaxPool 3x3 / 2 -> pointwise Conv as ONE tcgen05 launch (window rows by TMA, maximum taken by the converter warps):
    same bits as MaxPool kernel + convolution launch, one launch fewer, and the oracle's values
    (max_pool_op.rs:157-360 zero-fill padding, convolution_op.rs:407-504)."""
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    from oracle import onnx_wire as ow, ref_model as rm
    path = str(tmp_path / "pool_squeeze.onnx")
    with open(path, "wb") as f:
        f.write(synth.build_pool_squeeze(cin, squeeze, hw, pads=pads, seed=cin + hw, tail=tail))
    xs = synth.synthetic_batch(batch, chw=(cin, hw, hw), seed=hw, std=1.0)
    if neg:
        xs = -np.abs(xs) - np.float32(0.25)
    want = rm.run_batch(ow.load_model(path), xs, threads=2)
    eng = Engine(path, ctx=ctx)
    fused = eng(xs)
    n_fused = eng.model.launches_per_run(batch)
    kinds = [p["kind"] for p in eng.model.profile(batch, iters=1)]
    assert_close(fused.reshape(want.shape), want, "pool fusion vs oracle")
    if neg is None:
        assert "maxpool" in kinds and "maxpool+conv_tc" not in kinds, kinds
        return
    assert "maxpool+conv_tc" in kinds and "maxpool" not in kinds, kinds
    eng.model.set_option("pool_fusion", 0)
    separate = eng(xs)
    assert eng.model.launches_per_run(batch) == n_fused + 1
    assert "maxpool" in [p["kind"] for p in eng.model.profile(batch, iters=1)]
    assert np.array_equal(separate, fused), "the fused launch must give the bits of MaxPool + Conv"
    eng.model.set_option("pool_fusion", 1)
    eng.model.set_option("alt_order", 0)
    assert np.array_equal(eng(xs), fused), "tile walking direction must not change the bits"


def test_pool_fusion_random_shapes(ctx, tmp_path):
    """Seeded sweep over map sizes, channel counts, filter counts, batch sizes and zero padding at either end: the pool-fused
    launch must give the bits of MaxPool + Conv (which the other tests pin against the oracle) whenever the planner fuses."""
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    rng = np.random.default_rng(20260)
    fused_cases = 0
    for case in range(14):
        cin = int(rng.integers(2, 33)) * 4
        squeeze = int(rng.integers(1, 17)) * 4
        hw = int(rng.integers(9, 121))
        pads = tuple(int(v) for v in rng.integers(0, 2, size=4))
        batch = int(rng.integers(1, 4))
        path = str(tmp_path / f"ps{case}.onnx")
        with open(path, "wb") as f:
            f.write(synth.build_pool_squeeze(cin, squeeze, hw, pads=pads, seed=case, tail=bool(case & 1)))
        xs = synth.synthetic_batch(batch, chw=(cin, hw, hw), seed=100 + case, std=1.0)
        if case % 3 == 0:
            xs = -np.abs(xs)
        eng = Engine(path, ctx=ctx)
        fused = eng(xs)
        kinds = [p["kind"] for p in eng.model.profile(batch, iters=1)]
        eng.model.set_option("pool_fusion", 0)
        separate = eng(xs)
        assert np.array_equal(separate, fused), f"case {case}: cin={cin} squeeze={squeeze} hw={hw} pads={pads} batch={batch} kinds={kinds}"
        fused_cases += "maxpool+conv_tc" in kinds
    assert fused_cases >= 10, fused_cases


def _assert_same_with_nonfinite(got, want, what):
    got = np.asarray(got); want = np.asarray(want)
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want)), f"{what}: NaN pattern differs ({int(np.isnan(got).sum())} vs {int(np.isnan(want).sum())})"
    inf = np.isinf(want)
    assert np.array_equal(np.isinf(got), inf) and np.array_equal(got[inf], want[inf]), f"{what}: Inf pattern differs"
    fin = np.isfinite(want)
    assert_close(got[fin], want[fin], what)


def test_nonfinite_inputs_follow_the_reference(ctx, tmp_path, synth_onnx):
    """An Inf and a NaN reaching the squeeze output of a Fire module: the reference multiplies real taps only
    (convolution_op.rs:480), so they stay local to the outputs that read them (the 1x1 expand of a neighbouring pixel stays
    finite).  The tensor-core path cannot do that (hi / lo split: Inf - Inf = NaN; the Fire expand fusion adds 0 * x
    terms), so the run's input stage detects non-finite input and b200_model_run falls back to the CUDA-core fp32 plan by
    itself; the asynchronous entries report it at b200_model_sync."""
    import torch
    from onnx_rusty_inference_engine_b200 import _lib as L, synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    from oracle import onnx_wire as ow, ref_model as rm
    path = str(tmp_path / "fire.onnx")
    with open(path, "wb") as f:
        f.write(synth.build_fire(8, 16, 64, 17, seed=3))          # fire2-like: the expand fusion applies
    xs = synth.synthetic_batch(3, chw=(8, 17, 17), seed=4, std=1.0)
    xs[0, 2, 5, 6] = np.inf
    xs[1, 0, 9, 9] = np.nan
    with np.errstate(all="ignore"):
        want = rm.run_batch(ow.load_model(path), xs, threads=2)
    assert np.isinf(want).any() and not np.isnan(want).any()      # upstream's Relu (f32::max) turns a NaN into 0
    eng = Engine(path, ctx=ctx)
    got = eng(xs)                                                  # guard trips -> fallback plan
    _assert_same_with_nonfinite(got.reshape(want.shape), want, "fire module with Inf / NaN input")
    clean = synth.synthetic_batch(3, chw=(8, 17, 17), seed=4, std=1.0)
    assert np.isfinite(eng(clean)).all(), "the guard must reset after a non-finite batch"
    # asynchronous entry: reported at sync, then cleared
    xh = torch.from_numpy(xs).pin_memory(); oh = torch.empty((3, eng.out_per_image)).pin_memory()
    eng.run_pinned_async(xh, oh)
    with pytest.raises(L.B200Error, match="Inf or NaN"):
        eng.sync()
    eng.run_pinned_async(torch.from_numpy(clean).pin_memory(), oh)
    eng.sync()
    # the whole network (space-to-depth stem + Fire fusions): same contract
    big = synth.synthetic_batch(2, seed=21)
    big[1, 1, 100, 37] = -np.inf
    with np.errstate(all="ignore"):
        want_big = rm.run_batch(ow.load_model(synth_onnx), big, threads=2)
    _assert_same_with_nonfinite(Engine(synth_onnx, ctx=ctx)(big), want_big, "SqueezeNet with an Inf input pixel")


def test_space_to_depth_stem_matches_plain_layout(ctx, synth_onnx):
    """conv1 (7x7 / 2 on 3 channels) on the 2x2 space-to-depth copy of the input (a 4x4 / 1 convolution over 12 channels with
    zero-padded weights) against the plain channel-padded layout and the oracle."""
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    from oracle import onnx_wire as ow, ref_model as rm
    xs = synth.synthetic_batch(3, seed=13)
    eng = Engine(synth_onnx, ctx=ctx)
    s2d = eng(xs)
    eng.model.set_option("s2d", 0)
    plain = eng(xs)
    assert_close(s2d, plain, "space-to-depth stem vs plain layout")
    want = rm.run_batch(ow.load_model(synth_onnx), xs[:2], threads=2)
    assert_close(s2d[:2], want, "space-to-depth stem vs oracle")
    assert (s2d.argmax(1) == plain.argmax(1)).all()


def test_alt_order_is_bitwise_neutral(ctx, synth_onnx):
    """Launches that walk their tiles in alternating directions (L2 reuse) must give the same bits: tiles are independent."""
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    xs = synth.synthetic_batch(5, seed=9)
    eng = Engine(synth_onnx, ctx=ctx)
    a = eng(xs)
    eng.model.set_option("alt_order", 0)
    b = eng(xs)
    assert np.array_equal(a, b)


def test_squeezenet_batch_properties(ctx, synth_onnx):
    """Config 3 at full size (batch 256): batch-position invariance and softmax normalisation."""
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    eng = Engine(synth_onnx, ctx=ctx)
    xs = synth.synthetic_batch(8, seed=2)
    ref8 = eng(xs)
    big = eng(np.tile(xs, (32, 1, 1, 1)))          # 256 images
    assert big.shape == (256, 1000)
    assert np.array_equal(big.reshape(32, 8, 1000), np.broadcast_to(ref8, (32, 8, 1000))), "batch-position invariance"
    assert np.allclose(big.sum(1), 1.0, atol=1e-5)
    assert np.isfinite(big).all() and (big >= 0).all()
    for _ in range(6):                               # run-to-run determinism at full size
        assert np.array_equal(eng(np.tile(xs, (32, 1, 1, 1))), big)


def test_squeezenet_batch256_sampled_vs_oracle(ctx, synth_onnx):
    """Config 3 as SURVEY.md section 8(d) states it: the batch-256 run itself (seed 1, N(0,10^2)), >= 8 images sampled FROM
    ITS OUTPUT -- spread over the persistent CTAs' tile schedules: first, last, both sides of the 148-SM boundary --
    against the oracle, north_star tolerance and identical argmax."""
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    from oracle import onnx_wire as ow, ref_model as rm
    xs = synth.synthetic_batch(256, seed=1)
    eng = Engine(synth_onnx, ctx=ctx)
    got = eng(xs)
    assert got.shape == (256, 1000)
    idx = [0, 37, 100, 148, 149, 200, 254, 255]
    want = rm.run_batch(ow.load_model(synth_onnx), xs[idx], threads=8)
    assert_close(got[idx], want, "squeezenet batch 256, sampled images vs oracle")
    assert (got[idx].argmax(1) == want.argmax(1)).all()
    assert np.allclose(got.sum(1), 1.0, atol=1e-5) and np.isfinite(got).all()


def test_arena_liveness_reuse_and_plan_eviction(ctx, synth_onnx):
    """SURVEY.md section 8b "arena with liveness reuse": activations whose launch ranges are disjoint share bytes (the reference
    frees nothing until inference() returns); plans of batch sizes not used recently are evicted.  Results unchanged."""
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    eng = Engine(synth_onnx, ctx=ctx)
    arena, bump = eng.model.arena_bytes(256)
    print(f"arena at batch 256: {arena / 1e9:.2f} GB with reuse, {bump / 1e9:.2f} GB without")
    # the largest live set is conv1's input + output (0.12 + 1.17 GB) ~ 1.3 GB: demand <= 2x that, and a big cut
    assert arena <= 2.6e9 and arena < 0.4 * bump
    xs = synth.synthetic_batch(9, seed=33)
    first = eng(xs[:3])
    for b in (1, 2, 4, 5, 6, 7, 8, 9):          # more batch sizes than the plan cache keeps
        got = eng(xs[:b])
        assert np.array_equal(got[:min(b, 3)], first[:min(b, 3)]), f"batch {b}"
    assert np.array_equal(eng(xs[:3]), first), "re-planned after eviction"


def test_squeezenet_batch2048_sampled_vs_oracle(synth_onnx):
    """Config 4's per-GPU worst case on ONE device: batch 2048 (seed 2), 4 images sampled from its output against the oracle."""
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    from oracle import onnx_wire as ow, ref_model as rm
    xs = synth.synthetic_batch(2048, seed=2)
    eng = Engine(synth_onnx, device=0)
    arena, bump = eng.model.arena_bytes(2048)
    print(f"arena at batch 2048: {arena / 1e9:.2f} GB with reuse, {bump / 1e9:.2f} GB without")
    got = eng(xs)
    idx = [0, 777, 1500, 2047]
    want = rm.run_batch(ow.load_model(synth_onnx), xs[idx], threads=4)
    assert_close(got[idx], want, "squeezenet batch 2048, sampled images vs oracle")
    assert (got[idx].argmax(1) == want.argmax(1)).all()
    assert np.isfinite(got).all()
    eng.model.close()


def test_engine_on_torch_stream(synth_onnx):
    """Device-resident entry (b200_model_run_device) on torch's current stream, input/outputs as torch tensors."""
    import torch
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    torch.cuda.set_device(0)
    s = torch.cuda.Stream()
    xs = synth.synthetic_batch(4, seed=5)
    with torch.cuda.stream(s):
        eng = Engine(synth_onnx, device=0, stream=s.cuda_stream)
        host = eng(xs)
        dev = eng.run_torch(torch.from_numpy(xs).cuda())
        s.synchronize()   # only this stream: the backend's work must have been ordered on it
        assert np.array_equal(dev.cpu().numpy(), host)


def test_async_pipelined_run_matches_sync(synth_onnx):
    """b200_model_run_async: several batches in flight (H2D / compute / D2H on three streams) == synchronous runs."""
    import torch
    from onnx_rusty_inference_engine_b200 import synth
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    eng = Engine(synth_onnx, device=0)
    xs = [torch.from_numpy(synth.synthetic_batch(3, seed=20 + i)).pin_memory() for i in range(5)]
    outs = [torch.zeros((3, 1000), dtype=torch.float32).pin_memory() for _ in range(5)]
    for x, o in zip(xs, outs):
        eng.run_pinned_async(x, o)
    eng.sync()
    for x, o in zip(xs, outs):
        assert np.array_equal(o.numpy(), eng(x.numpy())), "pipelined result differs from the synchronous run"


def test_unknown_op_and_attr_are_errors(ctx):
    """model_inference.rs:158 (unknown op) and convolution_op.rs:160 (unknown attribute) panic upstream; here they
    are B200_EUNSUPPORTED at load time."""
    from onnx_rusty_inference_engine_b200 import _lib as L, onnx_proto as P
    w = np.zeros((2, 1, 1, 1), np.float32)
    def model(nodes):
        g = {"node": nodes, "name": "t", "initializer": [P.make_tensor("w", w)],
             "input": [P.make_value_info("x", [1, 1, 4, 4]), P.make_value_info("w", w.shape)],
             "output": [P.make_value_info("y", [1, 2, 4, 4])]}
        return P.encode("ModelProto", {"ir_version": 3, "graph": g, "opset_import": [{"domain": "", "version": 8}]})
    with pytest.raises(L.B200Error, match="NOT FOUND FOR NODE"):
        L.Model(ctx, model([P.make_node("Sigmoid", ["x"], ["y"], name="s")]))
    with pytest.raises(L.B200Error, match="ATTRIBUTE NAME FOR CONVOLUTION NOT FOUND"):
        L.Model(ctx, model([P.make_node("Conv", ["x", "w"], ["y"], name="c", strides=[1, 1], foo=1)]))
    with pytest.raises(L.B200Error, match="strides"):
        L.Model(ctx, model([P.make_node("Conv", ["x", "w"], ["y"], name="c")]))
    ok = L.Model(ctx, model([P.make_node("Conv", ["x", "w"], ["y"], name="c", strides=[1, 1])]))
    out = ok.run(np.ones((3, 1, 4, 4), np.float32))
    assert out.shape == (3, 32) and np.array_equal(out, np.zeros((3, 32), np.float32))
