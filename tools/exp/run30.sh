#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 240 python tools/tc_check.py > gpurun_out/tc_check.txt 2>&1; echo "tc_check rc=$?"; tail -2 gpurun_out/tc_check.txt
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python tools/tc_bench.py > gpurun_out/sweep30.txt 2>&1; cat gpurun_out/sweep30.txt
timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-150 gpurun_out/bench_a.json; tail -2 gpurun_out/bench.err
