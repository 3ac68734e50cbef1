#!/bin/bash
# A/B of a model option inside one process-independent pair of bench runs (bench.py --option k=v), interleaved, ONE gpurun call
# usage: ab_option.sh pool_fusion     (runs the default and k=0)
K=${1:-pool_fusion}
O=gpurun_out/ab_option; mkdir -p $O; rm -f $O/*
for i in 1 2 3; do
  timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extra 2>/dev/null > $O/on_$i.json
  timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extra --model-opt $K=0 2>/dev/null > $O/off_$i.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_option/*.json")):
    try: d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, "no line", e); continue
    L = d["roofline"].get("launches") or []
    sel = [x for x in L if "pool" in x["name"] or "squeeze" in x["name"]][:8]
    print(f, round(d["value"]), round(d["ms_per_step"], 4), d["clocks"].get("sm_mhz"), d["gpu_launches"], [(str(x["name"])[:26], x["kind"][:8], round(x["ms"], 4), x.get("gbs")) for x in sel])
PY
