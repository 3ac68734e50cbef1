"""B200-native fp32 operator backend for the hot path of jackperlo/onnx-rusty-inference-engine.

  _lib                 ctypes binding of libb200rt.so (include/b200rt.h) -- the C-ABI drop-in boundary
  inference_fp32_ops   the reference's ten op functions over device-resident tensors
  inference_engine     inference() / node_inference() and the graph-level Engine (CUDA graph)
  group17              onnx_make_inference(...) -- the reference's Python entry point
  sharding             batch sharding across GPUs (one process per GPU, torch.distributed)
  synth                seeded synthetic SqueezeNet1.0-8 generator
Importing the package does not load the CUDA library; the first call does, and fails loudly without it.
"""
__all__ = ["_lib", "inference_fp32_ops", "inference_engine", "group17", "sharding", "synth", "onnx_proto"]
