#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py -x -q 2>&1 | tail -3
H=$PWD/onnx_rusty_inference_engine_b200/lib/variants/libb200rt_head.so
echo "== new" > gpurun_out/sweep29.txt
timeout -s KILL 300 python tools/tc_bench.py conv1 f2_fused f4_e3 f6_e3 f8_e3 f2_sq f8_sq f4_e1 conv10 >> gpurun_out/sweep29.txt 2>&1
echo "== head" >> gpurun_out/sweep29.txt
B200RT_LIB=$H timeout -s KILL 300 python tools/tc_bench.py conv1 f2_fused f4_e3 f6_e3 f8_e3 f2_sq f8_sq f4_e1 conv10 >> gpurun_out/sweep29.txt 2>&1
cat gpurun_out/sweep29.txt
for i in 1 2; do
timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench.err; echo "new rc=$?"; cut -c1-150 gpurun_out/bench_a.json; tail -2 gpurun_out/bench.err
B200RT_LIB=$H timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench.err; echo "head rc=$?"; cut -c1-150 gpurun_out/bench_b.json
done
