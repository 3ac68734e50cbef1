#!/bin/bash
# A/B of an environment knob of the library in ONE gpurun call: tools/exp/ab_envvar.sh VAR "layers..."
V=$1; L=${2:-conv1}
O=gpurun_out/ab_env; mkdir -p $O; : > $O/ab.txt
for i in 1 2; do
  echo "== $V=1" >> $O/ab.txt; env $V=1 timeout 300 python tools/tc_bench.py $L >> $O/ab.txt 2>&1
  echo "== $V=0" >> $O/ab.txt; env $V=0 timeout 300 python tools/tc_bench.py $L >> $O/ab.txt 2>&1
done
for i in 1 2 3; do
  echo "== $V=1 bench" >> $O/ab.txt; env $V=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extra 2>/dev/null | cut -c1-110 >> $O/ab.txt
  echo "== $V=0 bench" >> $O/ab.txt; env $V=0 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extra 2>/dev/null | cut -c1-110 >> $O/ab.txt
done
cat $O/ab.txt
