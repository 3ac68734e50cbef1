#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 240 python tools/tc_check.py > gpurun_out/tc_check.txt 2>&1; echo "tc_check rc=$?"; tail -4 gpurun_out/tc_check.txt
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for m in 0; do
echo "== mask $m" >> gpurun_out/sweep13.txt
B200_TC_DEBUG=$m timeout -s KILL 300 python tools/tc_bench.py >> gpurun_out/sweep13.txt 2>&1
done
cat gpurun_out/sweep13.txt
timeout -s KILL 600 python bench.py --profile-out gpurun_out/per_launch.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.json
