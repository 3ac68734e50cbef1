#!/bin/bash
mkdir -p gpurun_out
for bn in 128 64; do for m in 35 163 291; do
  echo "== BN=$bn B200_TC_DEBUG=$m" >> gpurun_out/sweep4.txt
  B200_TC_BN=$bn B200_TC_DEBUG=$m timeout 300 python tools/tc_bench.py f4_e3 f8_e3 conv10 >> gpurun_out/sweep4.txt 2>&1
done; done
cat gpurun_out/sweep4.txt
