#!/bin/bash
# Which role bounds the fused fire2 launch?  tc_bench on f2_fused / f2_e3 under the role-mask build (make debug):
# 2 = no global loads, 4 = no global stores, 16 = no MMAs, 32 = no split / tcgen05.st, 256 = N = 16 MMAs
O=gpurun_out/masks_f2; mkdir -p $O
D=$PWD/onnx_rusty_inference_engine_b200/lib/variants/libb200rt_dbg.so
for m in 0 2 4 16 256 6 34 20 22 38 54; do
  echo "== mask=$m" >> $O/masks.txt
  B200_TC_DEBUG=$m B200RT_LIB=$D timeout 120 python tools/tc_bench.py f2_fused f2_e3 2>&1 >> $O/masks.txt
done
cat $O/masks.txt
