#!/bin/bash
mkdir -p gpurun_out
echo "== default" > gpurun_out/sweep23.txt
timeout -s KILL 300 python tools/tc_bench.py tiny1 tiny148 f5_e1 f6_e1 f8_e1 f9_e1 f9_e3 f9_sq f4_e1 >> gpurun_out/sweep23.txt 2>&1
echo "== BN=64" >> gpurun_out/sweep23.txt
B200_TC_BN=64 timeout -s KILL 300 python tools/tc_bench.py f5_e1 f6_e1 f8_e1 f9_e1 f9_e3 f4_e1 >> gpurun_out/sweep23.txt 2>&1
echo "== NACC=2" >> gpurun_out/sweep23.txt
B200_TC_NACC=2 timeout -s KILL 300 python tools/tc_bench.py f6_e1 >> gpurun_out/sweep23.txt 2>&1
cat gpurun_out/sweep23.txt
