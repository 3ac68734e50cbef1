"""Python entry point with the reference's module name and signature.

The reference declares a PyO3 module `group17` exporting
    onnx_make_inference(onnx_file: String, input_path: &str, output_path: &str, input_tensor_name: Vec<&str>)
(src/lib.rs:14-31, built by maturin per pyproject.toml:1-16; the body is commented out upstream).  It follows
main.rs:27-42: parse the model, read the input and expected-output TensorProto files, run inference, print the
predicted class and the expected data.  This implementation does the same through libb200rt.so and additionally
RETURNS the output array (the reference returns nothing and never compares).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from . import _lib as L
from .inference_engine import Engine


def onnx_make_inference(onnx_file: str, input_path: str, output_path: str, input_tensor_name: Sequence[str],
                        device: int = 0, quiet: bool = False) -> np.ndarray:
    x = L.read_tensor_pb(input_path)            # read_input_data, main.rs:36,44-53
    expected = L.read_tensor_pb(output_path)    # main.rs:37
    eng = Engine(onnx_file, device=device)
    # input_tensor_name: entries that are initializers are skipped upstream (utils.rs:35); the engine feeds the
    # single non-initializer graph input, which is what remains of the list for both bundled models.
    c, h, w = eng.in_chw
    if x.size != c * h * w:
        raise L.B200Error(-1, "input length != static model shape (utils.rs:40 from_shape_vec unwrap)")
    out = eng(x.reshape(1, c, h, w))
    if not quiet:
        best = int(np.argmax(out[0])) + 1       # 1-based, softmax_op.rs:36 / add_op.rs:100
        tag = "Squeezenet1.0-8" if out.shape[1] == 1000 else "MNist-8"
        print(f"\n{tag} Inference results: Class {best}-nth predicted.\nActual Data: {out.tolist()}")
        print(f"Expected Data: {expected.reshape(-1).tolist()}")  # main.rs:41
    return out
