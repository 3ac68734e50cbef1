#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 240 python tools/tc_check.py > gpurun_out/tc_check.txt 2>&1; echo "tc_check rc=$?"; tail -3 gpurun_out/tc_check.txt
B200_TC_EPI16=1 timeout -s KILL 240 python tools/tc_check.py > gpurun_out/tc_check16.txt 2>&1; echo "tc_check(epi16 forced) rc=$?"; tail -3 gpurun_out/tc_check16.txt
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for e in 0 -1 1; do
echo "== EPI16=$e" >> gpurun_out/sweep27.txt
B200_TC_EPI16=$e timeout -s KILL 300 python tools/tc_bench.py f2_e1 f4_e1 f5_e1 f6_e1 f8_e1 f9_e1 f2_sq f8_sq f9_sq conv10 >> gpurun_out/sweep27.txt 2>&1
done
cat gpurun_out/sweep27.txt
for i in 1 2; do
timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench.err; echo "default rc=$?"; cut -c1-150 gpurun_out/bench_a.json; tail -2 gpurun_out/bench.err
B200_TC_EPI16=0 timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench.err; echo "epi16 off rc=$?"; cut -c1-150 gpurun_out/bench_b.json
done
