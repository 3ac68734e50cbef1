#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for m in 0 35; do
  echo "== B200_TC_DEBUG=$m" >> gpurun_out/sweep5.txt
  B200_TC_DEBUG=$m timeout 300 python tools/tc_bench.py >> gpurun_out/sweep5.txt 2>&1
done
cat gpurun_out/sweep5.txt
timeout 600 python bench.py --profile-out gpurun_out/per_launch.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
