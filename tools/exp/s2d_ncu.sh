mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/l_on.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/l_off.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --model-opt s2d=0 > /dev/null 2>&1
for f in on off; do echo "== $f"; grep -E "s2d|nchw_to_rows|conv_tc_kernel" gpurun_out/l_$f.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*)//' | head -60 | awk '{k=$1; v=$NF; gsub(/"/,"",v); print k, v}' | sed -n 1,12p; done
