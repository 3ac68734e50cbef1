#!/bin/bash
# A/B of per-launch times (bench.py's roofline.launches table) between the in-tree build and lib/variants/libb200rt_head.so
O=gpurun_out/ab_launches; mkdir -p $O
H=$PWD/onnx_rusty_inference_engine_b200/lib/variants/libb200rt_head.so
for i in 1 2 3; do
  timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extra 2>/dev/null > $O/new_$i.json
  B200RT_LIB=$H timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extra 2>/dev/null > $O/old_$i.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_launches/*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    L = d["roofline"].get("launches") or []
    print(f, round(d["value"]), d["ms_per_step"], d["clocks"].get("sm_mhz"), [(str(x["name"])[-28:], x["ms"]) for x in L[:9]])
PY
