#!/bin/bash
# A/B of the raw-hi split (B200_SPLIT_RAW_HI=1, in-tree build) against the round-1 split (variants/libb200rt_head.so) in ONE call
O=gpurun_out/r2h; mkdir -p $O
H=$PWD/onnx_rusty_inference_engine_b200/lib/variants/libb200rt_head.so
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py -m gpu -x -q 2>&1 | tail -3
L="conv1 f2_fused f4_e3 f6_e3 f8_e3 f2_sq f8_sq f4_e1 conv10"
echo "== new" > $O/ab.txt; timeout 300 python tools/tc_bench.py $L >> $O/ab.txt 2>&1
echo "== old" >> $O/ab.txt; B200RT_LIB=$H timeout 300 python tools/tc_bench.py $L >> $O/ab.txt 2>&1
for i in 1 2; do
  echo "== new mnist/bench" >> $O/ab.txt; timeout 100 python tools/mnist_bench.py 65536 | cut -c1-100 >> $O/ab.txt; timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extra 2>/dev/null | cut -c1-110 >> $O/ab.txt
  echo "== old mnist/bench" >> $O/ab.txt; B200RT_LIB=$H timeout 100 python tools/mnist_bench.py 65536 | cut -c1-100 >> $O/ab.txt; B200RT_LIB=$H timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extra 2>/dev/null | cut -c1-110 >> $O/ab.txt
done
cat $O/ab.txt
