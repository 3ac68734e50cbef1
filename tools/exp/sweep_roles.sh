#!/bin/bash
# GPU experiment batch 1: tests, bench, TMEM st layout discovery, role-isolation sweep.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 60 tools/exp/build/tmem_layout > gpurun_out/tmem_layout.txt 2>&1; echo "tmem_layout rc=$?"
for m in 0 1 2 32 34 35 39 16 4; do
  echo "== B200_TC_DEBUG=$m" >> gpurun_out/sweep.txt
  B200RT_LIB=$PWD/onnx_rusty_inference_engine_b200/lib/variants/libb200rt_dbg.so B200_TC_DEBUG=$m timeout 300 python tools/tc_bench.py conv1 f2_e3 f4_e3 f8_e3 conv10 f8_sq f4_e1 >> gpurun_out/sweep.txt 2>&1
done
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; cat gpurun_out/sweep.txt
