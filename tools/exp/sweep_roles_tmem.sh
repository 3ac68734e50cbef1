#!/bin/bash
mkdir -p gpurun_out
for m in 35 1 2 32 34 3 4 39; do
  echo "== B200_TC_DEBUG=$m" >> gpurun_out/sweep8.txt
  B200RT_LIB=$PWD/onnx_rusty_inference_engine_b200/lib/variants/libb200rt_dbg.so B200_TC_DEBUG=$m timeout -s KILL 300 python tools/tc_bench.py conv1 f2_e3 f4_e3 f8_e3 conv10 f4_e1 f8_sq >> gpurun_out/sweep8.txt 2>&1
done
cat gpurun_out/sweep8.txt
