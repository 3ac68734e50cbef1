#!/bin/bash
# A/B of a model option inside ONE gpurun call (after the parity tests): bash tools/exp/ab_modelopt.sh fire_fusion_wide
mkdir -p gpurun_out
OPT=${1:-fire_fusion_wide}
timeout -s KILL 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout -s KILL 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/pl_on.json > gpurun_out/bench_on.json 2> gpurun_out/bench.err; echo "$OPT=1 rc=$?"; cut -c1-150 gpurun_out/bench_on.json; tail -2 gpurun_out/bench.err
timeout -s KILL 600 python bench.py --no-cpu-baseline --model-opt $OPT=0 --profile-out gpurun_out/pl_off.json > gpurun_out/bench_off.json 2> gpurun_out/bench.err; echo "$OPT=0 rc=$?"; cut -c1-150 gpurun_out/bench_off.json
done
