#!/bin/bash
mkdir -p gpurun_out
for m in 0 64 128 4 68 132; do
echo "== mask $m" >> gpurun_out/sweep14.txt
B200_TC_DEBUG=$m timeout -s KILL 300 python tools/tc_bench.py f2_e1 f4_e1 f4_e3 f8_e3 conv10 >> gpurun_out/sweep14.txt 2>&1
done
cat gpurun_out/sweep14.txt
