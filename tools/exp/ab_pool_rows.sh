#!/bin/bash
# pooled rows per tile of the pool-fused launches (B200_TC_POOL_ROWS, experiments only): per-launch times from bench.py
O=gpurun_out/ab_pool_rows; mkdir -p $O; rm -f $O/*
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "pool_fusion" 2>&1 | tail -3
for r in 0 1 2 3 4 7; do
  B200_TC_POOL_ROWS=$r timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extra 2>/dev/null > $O/rows_$r.json
done
timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extra --model-opt pool_fusion=0 2>/dev/null > $O/off.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_pool_rows/*.json")):
    try: d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, "no line", e); continue
    L = d["roofline"].get("launches") or []
    sel = [x for x in L if "pool" in x["name"] or x["name"].startswith(("fire2_s", "fire5_s", "fire9_s"))]
    print(f, round(d["value"]), round(d["ms_per_step"], 4), d["clocks"].get("sm_mhz"), [(str(x["name"])[:12], round(x["ms"], 4), x.get("gbs")) for x in sel])
PY
