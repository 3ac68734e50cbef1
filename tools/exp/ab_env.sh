#!/bin/bash
# A/B of an experiment environment variable (default: B200_TC_NO_EMR=1 = the old single-barrier drain) on per-launch times
# (bench.py's roofline.launches table) and tc_bench layers, interleaved in ONE gpurun call.   usage: ab_env.sh [VAR=VALUE]
V=${1:-B200_TC_NO_EMR=1}
O=gpurun_out/ab_env; mkdir -p $O; rm -f $O/*
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py -m gpu -x -q 2>&1 | tail -4
L="conv1 f2_fused f4_e3 f6_e3 f8_e3 f9_e3 f4_e1 f8_e1 conv10"
for i in 1 2; do
  echo "== new" >> $O/layers.txt; timeout 300 python tools/tc_bench.py $L >> $O/layers.txt 2>&1
  echo "== $V" >> $O/layers.txt; env $V timeout 300 python tools/tc_bench.py $L >> $O/layers.txt 2>&1
done
for i in 1 2 3; do
  timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extra 2>/dev/null > $O/new_$i.json
  env $V timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extra 2>/dev/null > $O/old_$i.json
done
cat $O/layers.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_env/*.json")):
    try: d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, "no line", e); continue
    L = d["roofline"].get("launches") or []
    big = sorted(L, key=lambda x: -x["ms"])[:8]
    print(f, round(d["value"]), round(d["ms_per_step"], 4), d["clocks"].get("sm_mhz"), d["roofline"]["regime"], [(str(x["name"])[:14], round(x["ms"], 4)) for x in big])
PY
