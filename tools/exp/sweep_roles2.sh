#!/bin/bash
mkdir -p gpurun_out
for m in 1 2 32 34 33 3 4 8 40 36; do
  echo "== B200_TC_DEBUG=$m" >> gpurun_out/sweep6.txt
  B200RT_LIB=$PWD/onnx_rusty_inference_engine_b200/lib/variants/libb200rt_dbg.so B200_TC_DEBUG=$m timeout 300 python tools/tc_bench.py conv1 f2_e3 f4_e3 f8_e3 conv10 >> gpurun_out/sweep6.txt 2>&1
done
cat gpurun_out/sweep6.txt
