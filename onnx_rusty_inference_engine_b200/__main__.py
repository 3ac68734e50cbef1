"""python -m onnx_rusty_inference_engine_b200 [model.onnx input.pb expected.pb input_name...]

The reference's binary entry (src/main.rs:9-53) in Python: with no arguments it runs what main.rs hard-codes (:17-20:
models/squeezenet1.0-8.onnx on squeezenet_data_0.pb / squeezenet_output_0.pb, input "data_0"); otherwise the four
arguments of read_and_make_inference (main.rs:27).  Prints the reference's result lines and exits 1 when the output is
outside the north_star tolerance of the expected data.  (The compiled equivalent is lib/onnx_rusty_inference_engine_bin.)
"""
import sys

import numpy as np

from . import _lib as L
from .group17 import onnx_make_inference


def main(argv) -> int:
    if len(argv) >= 3:
        onnx_file, input_path, output_path, names = argv[0], argv[1], argv[2], argv[3:]
    elif not argv:
        onnx_file, input_path, output_path, names = ("models/squeezenet1.0-8.onnx", "squeezenet_data_0.pb",
                                                     "squeezenet_output_0.pb", ["data_0"])
    else:
        print(__doc__, file=sys.stderr)
        return 2
    out = onnx_make_inference(onnx_file, input_path, output_path, names)
    want = L.read_tensor_pb(output_path).reshape(1, -1)
    if want.shape != out.shape:
        return 0
    ratio = float((np.abs(out - want) / (1e-5 + 1e-4 * np.abs(want))).max())
    ok = ratio <= 1.0 and int(out.argmax()) == int(want.argmax())
    print(f"Match (1e-4 rel + 1e-5 abs, argmax): {'yes' if ok else 'NO'} (max err/tol {ratio:.3f})")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
