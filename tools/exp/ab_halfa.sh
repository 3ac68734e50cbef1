#!/bin/bash
# half-stage A ring + two accumulator stages at 64 < BN <= 96 (default) against one accumulator stage + four A stages (B200_TC_NO_HALFA=1)
O=gpurun_out/ab_halfa; mkdir -p $O; rm -f $O/*
timeout 240 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
  for n in 0 1; do echo "== NO_HALFA=$n" >> $O/halfa.txt; B200_TC_NO_HALFA=$n timeout 120 python tools/tc_bench.py conv1 f6_e3 f4_e3 >> $O/halfa.txt 2>&1; done
done
cat $O/halfa.txt
