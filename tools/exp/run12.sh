#!/bin/bash
mkdir -p gpurun_out
for v in 000 100 010 001 111; do
echo "== variant $v (packed split, k table, x32 drain)" >> gpurun_out/sweep12.txt
B200RT_LIB=$PWD/onnx_rusty_inference_engine_b200/lib/variants/libb200rt_$v.so timeout -s KILL 300 python tools/tc_bench.py conv1 f2_e3 f4_e3 f8_e3 conv10 f4_e1 >> gpurun_out/sweep12.txt 2>&1
done
cat gpurun_out/sweep12.txt
