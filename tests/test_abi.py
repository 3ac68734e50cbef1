"""The C-ABI library without a GPU: it loads, exports every symbol include/b200rt.h declares, refuses to run
without a device (no CPU fallback) and its host-side geometry matches the oracle's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from onnx_rusty_inference_engine_b200 import _lib as L


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "b200rt.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    syms = _declared_symbols()
    assert len(syms) >= 35
    lib = C.CDLL(L.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200rt.h but not exported"
        assert s in L.SIGNATURES, f"{s} has no ctypes signature in _lib.py"
    assert set(L.SIGNATURES) == set(syms)
    assert L.lib().b200_version().startswith(b"b200rt")


def test_rust_sys_crate_names_exactly_the_header_symbols():
    """rust/b200rt-sys/src/lib.rs is generated from the header (tools/gen_b200rt_sys.py; no Rust toolchain here to compile
    it): it must be up to date and declare every exported symbol, and the wrapper crate keeps the reference's ten op names."""
    import subprocess, sys
    gen = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_b200rt_sys.py")], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(ROOT, "rust", "b200rt-sys", "src", "lib.rs")) as f:
        committed = f.read()
    assert committed == gen, "rust/b200rt-sys/src/lib.rs is stale: python tools/gen_b200rt_sys.py > rust/b200rt-sys/src/lib.rs"
    assert sorted(set(re.findall(r"pub fn (b200_[a-z0-9_]+)\(", committed))) == _declared_symbols()
    ops = os.path.join(ROOT, "rust", "onnx-rusty-inference-engine-b200", "src", "inference_fp32_ops")
    for mod, fn in (("convolution_op", "convolution"), ("max_pool_op", "max_pool"), ("relu_op", "relu"), ("add_op", "add"), ("mul_op", "mul"),
                    ("reshape_op", "reshape"), ("concatenate_op", "concatenation"), ("dropout_op", "drop_out"),
                    ("global_average_pool_op", "global_average_pool"), ("softmax_op", "softmax")):
        with open(os.path.join(ops, mod + ".rs")) as f:
            assert f"pub fn {fn}(output_container: &Store" in f.read(), (mod, fn)


def test_no_cpu_fallback():
    if L.lib().b200_device_count() > 0:
        pytest.skip("a B200 is present")
    with pytest.raises(L.B200Error) as e:
        L.Context(0)
    assert e.value.code == -6   # B200_ENODEVICE
    from onnx_rusty_inference_engine_b200.inference_engine import Engine
    with pytest.raises(L.B200Error):
        Engine(os.path.join(ROOT, "tests", "golden", "mnist-8.onnx"))


def test_run_sharded_error_paths_without_a_gpu():
    """b200_model_run_sharded validates its arguments before touching a device (no compute without a GPU)."""
    lib = L.lib()
    x = np.zeros((4,), np.float32); y = np.zeros((4,), np.float32)
    assert lib.b200_model_run_sharded(None, 0, x.ctypes.data_as(C.c_void_p), 1, y.ctypes.data_as(C.c_void_p)) == -1
    arr = (C.c_void_p * 1)(None)
    assert lib.b200_model_run_sharded(arr, 1, x.ctypes.data_as(C.c_void_p), 1, y.ctypes.data_as(C.c_void_p)) == -1
    assert b"models[0] is NULL" in lib.b200_last_error()
    assert lib.b200_ctx_stream(None) is None


def _conv_dims(x, w, strides, pads, auto_pad):
    p = L.ConvParams(L._i64arr(strides, 2), L._i64arr(pads, 4), L._i64arr((0, 0), 2), 0, auto_pad, 0)
    y = (C.c_int64 * 4)()
    rc = L.lib().b200_conv2d_out_dims(L._i64arr(x), L._i64arr(w), C.byref(p), y)
    return rc, tuple(y)


def test_geometry_matches_oracle():
    """Host-side shape inference == the oracle's (convolution_op.rs:293-324, max_pool_op.rs:215-246)."""
    from oracle import ref_ops as R
    rng = np.random.default_rng(0)
    n_ok = 0
    for _ in range(300):
        H, W = int(rng.integers(3, 40)), int(rng.integers(3, 40))
        kh, kw = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        sh, sw = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        ap = int(rng.integers(0, 4))
        pads = [int(v) for v in rng.integers(0, 3, 4)] if ap == 3 else [0, 0, 0, 0]
        x = np.zeros((1, H, W), np.float32); w = np.zeros((2, 1, kh, kw), np.float32)
        try:
            want = R.conv2d_image(x, w, None, ap, pads, (sh, sw)).shape
        except R.RefPanic:
            want = None
        rc, got = _conv_dims((1, 1, H, W), (2, 1, kh, kw), (sh, sw), pads, ap)
        if want is None:
            assert rc != 0, (H, W, kh, kw, sh, sw, ap, pads)
        else:
            assert rc == 0 and got == (1, 2) + want[1:], (H, W, kh, kw, sh, sw, ap, pads, got, want)
            n_ok += 1
        pp = L.PoolParams(L._i64arr((kh, kw), 2), L._i64arr((sh, sw), 2), L._i64arr(pads, 4), ap, 0)
        y = (C.c_int64 * 4)()
        rc = L.lib().b200_maxpool2d_out_dims(L._i64arr((1, 1, H, W)), C.byref(pp), y)
        try:
            wantp = R.maxpool_image(x, (kh, kw), ap, pads, (sh, sw)).shape
        except R.RefPanic:
            wantp = None
        if wantp is None:
            assert rc != 0
        else:
            assert rc == 0 and tuple(y) == (1, 1) + wantp[1:]
    assert n_ok > 100


def test_reference_panics_become_error_codes():
    rc, _ = _conv_dims((1, 1, 8, 8), (2, 1, 3, 3), (0, 0), (0, 0, 0, 0), 0)     # strides missing (convolution_op.rs:285)
    assert rc == -1 and b"strides" in L.lib().b200_last_error()
    rc, _ = _conv_dims((1, 3, 8, 8), (2, 1, 3, 3), (1, 1), (0, 0, 0, 0), 0)     # C mismatch (convolution_op.rs:252)
    assert rc == -1
    rc, got = _conv_dims((1, 1, 8, 8), (2, 1, 3, 3), (1, 1), (1, 1, 1, 1), 0)   # pad promotion VALID -> NOTSET (:169-173)
    assert rc == 0 and got == (1, 2, 8, 8)
    pp = L.PoolParams(L._i64arr((3, 3), 2), L._i64arr((2, 2), 2), L._i64arr((0, 0, 1, 1), 4), 0, 0)
    y = (C.c_int64 * 4)()
    assert L.lib().b200_maxpool2d_out_dims(L._i64arr((1, 4, 54, 54)), C.byref(pp), y) == 0
    assert tuple(y) == (1, 4, 26, 26)                                             # pads ignored without NOTSET
    pp.auto_pad = 3
    assert L.lib().b200_maxpool2d_out_dims(L._i64arr((1, 4, 54, 54)), C.byref(pp), y) == 0
    assert tuple(y) == (1, 4, 27, 27)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm): one JSON line on stdout with the
    contract's keys and the GPU arm's metric / unit / workload; under torchrun only rank 0 works and prints."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, cwd=root, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "SqueezeNet1.0 images/sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "batch 256" in d["config"]["workload"] and "model" not in d["config"]
    env2 = dict(env, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r2 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, env=env2, cwd=root, timeout=120)
    assert r2.returncode == 0 and r2.stdout.strip() == "", (r2.returncode, r2.stdout, r2.stderr[-500:])
