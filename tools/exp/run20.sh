#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python tools/tc_bench.py > gpurun_out/sweep20.txt 2>&1; cat gpurun_out/sweep20.txt
timeout -s KILL 600 python bench.py --profile-out gpurun_out/per_launch.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
