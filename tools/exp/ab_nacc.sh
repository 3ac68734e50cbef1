#!/bin/bash
# two accumulator stages + two A stages at BN = 96 (B200_TC_NACC=2, experiments only) against one + four: conv1 and fire6 expand3x3
O=gpurun_out/ab_nacc; mkdir -p $O; rm -f $O/*
for i in 1 2; do
  for n in 0 2; do echo "== NACC=$n" >> $O/nacc.txt; B200_TC_NACC=$n timeout 200 python tools/tc_bench.py conv1 f6_e3 f6_e1 >> $O/nacc.txt 2>&1; done
done
cat $O/nacc.txt
