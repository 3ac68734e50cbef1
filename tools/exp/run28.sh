#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench.err; echo "default rc=$?"; cut -c1-150 gpurun_out/bench_a.json; tail -2 gpurun_out/bench.err
B200_TC_EPI16=0 timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench.err; echo "epi16 off rc=$?"; cut -c1-150 gpurun_out/bench_b.json
done
