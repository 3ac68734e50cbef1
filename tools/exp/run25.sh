#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/bench_hint.json 2> gpurun_out/bench.err; echo "hint rc=$?"; cut -c1-150 gpurun_out/bench_hint.json; tail -2 gpurun_out/bench.err
B200_NO_L2HINT=1 timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/bench_nohint.json 2> gpurun_out/bench.err; echo "nohint rc=$?"; cut -c1-150 gpurun_out/bench_nohint.json
done
