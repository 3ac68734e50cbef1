#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 240 python tools/tc_check.py > gpurun_out/tc_check.txt 2>&1; echo "tc_check rc=$?"; tail -11 gpurun_out/tc_check.txt
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for m in 0 64 2 66; do
echo "== mask $m" >> gpurun_out/sweep11.txt
B200_TC_DEBUG=$m timeout -s KILL 300 python tools/tc_bench.py conv1 f2_e3 f4_e3 f8_e3 conv10 f4_e1 >> gpurun_out/sweep11.txt 2>&1
done
cat gpurun_out/sweep11.txt
