#!/bin/bash
# Regenerates the evidence under profiles/ (run under gpurun; outputs land in gpurun_out/artifacts).
O=gpurun_out/artifacts; mkdir -p $O
python bench.py --steps 20 --warmup 3 --profile-out $O/per_launch_events.json > $O/bench.json 2> $O/bench.err || exit 1
python bench.py --steps 2 --warmup 3 > $O/plain_launches.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 > $O/ncu_launches.log 2>&1
for L in conv1 f8_e3 f8_sq f4_e1; do
  python tools/tc_bench.py $L > $O/plain_$L.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o $O/prof_$L python tools/tc_bench.py $L > $O/ncu_$L.log 2>&1
  echo "$L rc=$?"
done
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/smi.txt
ls -la $O
