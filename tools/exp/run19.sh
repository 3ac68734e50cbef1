#!/bin/bash
mkdir -p gpurun_out
for L in conv1 f4_e1; do
  timeout -s KILL 300 python tools/tc_bench.py $L > gpurun_out/plain_$L.log 2>&1 &&
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/prof2_$L python tools/tc_bench.py $L > gpurun_out/ncu_$L.log 2>&1
  echo "$L rc=$?"
done
