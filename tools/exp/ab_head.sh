#!/bin/bash
# A/B of the in-tree build against lib/variants/libb200rt_head.so (the previous commit's conv_tc.cu) in ONE gpurun call
O=gpurun_out/ab_head; mkdir -p $O
H=$PWD/onnx_rusty_inference_engine_b200/lib/variants/libb200rt_head.so
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q 2>&1 | tail -2
L="conv1 f2_fused f2_e3 f4_e3 f6_e3 f8_e3 f9_e3 f2_sq f8_sq f4_e1 f6_e1 conv10"
for i in 1 2; do
echo "== new" >> $O/ab.txt; timeout 300 python tools/tc_bench.py $L >> $O/ab.txt 2>&1
echo "== old" >> $O/ab.txt; B200RT_LIB=$H timeout 300 python tools/tc_bench.py $L >> $O/ab.txt 2>&1
done
for i in 1 2; do
  echo "== new bench" >> $O/ab.txt; timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extra 2>/dev/null | cut -c1-110 >> $O/ab.txt
  echo "== old bench" >> $O/ab.txt; B200RT_LIB=$H timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extra 2>/dev/null | cut -c1-110 >> $O/ab.txt
done
cat $O/ab.txt
