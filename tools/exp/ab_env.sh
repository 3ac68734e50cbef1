#!/bin/bash
# A/B of an environment knob inside ONE gpurun call: bash tools/exp/ab_env.sh B200_TC_MERGED 0
# (default build/knob first, then KNOB=VALUE), conv parity tests first.
mkdir -p gpurun_out
KNOB=${1:-B200_TC_MERGED}; VAL=${2:-0}
timeout -s KILL 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== default"; timeout -s KILL 300 python tools/tc_bench.py 2>&1 | tail -20
echo "== $KNOB=$VAL"; env $KNOB=$VAL timeout -s KILL 300 python tools/tc_bench.py 2>&1 | tail -20
for i in 1 2; do
timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/bench_on.json 2> gpurun_out/bench.err; echo "default rc=$?"; cut -c1-150 gpurun_out/bench_on.json; tail -2 gpurun_out/bench.err
env $KNOB=$VAL timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/bench_off.json 2> gpurun_out/bench.err; echo "$KNOB=$VAL rc=$?"; cut -c1-150 gpurun_out/bench_off.json
done
