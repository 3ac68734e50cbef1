#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout -s KILL 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/pl_alt.json > gpurun_out/bench_alt.json 2> gpurun_out/bench.err; echo "alt rc=$?"; cut -c1-150 gpurun_out/bench_alt.json; tail -2 gpurun_out/bench.err
timeout -s KILL 600 python bench.py --no-cpu-baseline --model-opt alt_order=0 --profile-out gpurun_out/pl_noalt.json > gpurun_out/bench_noalt.json 2> gpurun_out/bench.err; echo "noalt rc=$?"; cut -c1-150 gpurun_out/bench_noalt.json
done
