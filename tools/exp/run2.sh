#!/bin/bash
mkdir -p gpurun_out
for m in 35 99 103 64; do
  echo "== B200_TC_DEBUG=$m" >> gpurun_out/sweep2.txt
  B200_TC_DEBUG=$m timeout 300 python tools/tc_bench.py conv1 f2_e3 f4_e3 f8_e3 conv10 >> gpurun_out/sweep2.txt 2>&1
done
cat gpurun_out/sweep2.txt
