#!/bin/bash
# Round-2 evidence under gpurun (one GPU): bench lines, per-launch events, ncu launch list with DRAM bytes, --set full captures.
O=gpurun_out/artifacts_r2; mkdir -p $O
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-out $O/per_launch_events.json > $O/bench_20steps.json 2> $O/bench_20.err; echo "bench 20 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_ref.err; echo "bench ref rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/plain_launches.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
for L in conv1 f2_fused; do
  timeout 200 python tools/tc_bench.py $L > $O/plain_$L.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o $O/prof_$L python tools/tc_bench.py $L > $O/ncu_$L.log 2>&1
  echo "$L rc=$?"
done
# the pool-fused launch (pool1 | fire2 squeeze) is the second conv_tc_kernel launch of a run
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -f -o $O/prof_pool1_sq python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $O/ncu_pool1_sq.log 2>&1
echo "pool1_sq rc=$?"
timeout 100 python tools/mnist_bench.py 65536 3 > $O/plain_mnist.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:mnist8 -s 4 -c 2 -f -o $O/prof_mnist python tools/mnist_bench.py 65536 3 > $O/ncu_mnist.log 2>&1
echo "mnist ncu rc=$?"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/smi.txt
ls -la $O
