#!/bin/bash
# narrower channel tiles (B200_TC_BN, experiments only): two accumulator stages fit below BN = 64, the drain overlaps the next tile
O=gpurun_out/ab_bn; mkdir -p $O; rm -f $O/*
L="conv1 f2_fused f4_e3 f6_e3 f8_e3 f9_e3 f4_e1 f8_e1 conv10"
for bn in 0 64 48 32; do
  echo "== BN=$bn" >> $O/bn.txt; B200_TC_BN=$bn timeout 300 python tools/tc_bench.py $L >> $O/bn.txt 2>&1
done
cat $O/bn.txt
