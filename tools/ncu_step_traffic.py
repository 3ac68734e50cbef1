"""Derive the per-step DRAM traffic summary from an ncu launch list of `bench.py --steps 2 --warmup 3`.
usage: ncu_step_traffic.py launches.csv > profiles/<round>_step_dram_traffic.json
The CSV comes from
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c N --csv
A step starts at the input transform (nchw_to_rows* or nchw_to_s2d*) and ends before the next one; the last COMPLETE step is kept."""
import csv, json, re, sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
launches = {}
order = []
for r in rows:
    i = int(r["ID"])
    if i not in launches:
        full = r["Kernel Name"]
        # conv_tc_kernel<HAS_ADD, EPI16, NOPAD, POOL, HALFA>: the pool-fused instances are bandwidth launches, counted apart
        pooled = re.search(r"conv_tc_kernel<(\(bool\))?\d, (\(bool\))?\d, (\(bool\))?\d, (\(bool\))?1, (\(bool\))?\d>", full) is not None
        launches[i] = {"kernel": full.split("(")[0][-48:] + ("POOL" if pooled else ""), "time": None, "dram_read": None, "dram_write": None}
        order.append(i)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    name = r["Metric Name"]
    if name == "gpu__time_duration.sum":
        launches[i]["time"] = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    elif name == "dram__bytes_read.sum":
        launches[i]["dram_read"] = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    elif name == "dram__bytes_write.sum":
        launches[i]["dram_write"] = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
seq = [launches[i] for i in order]
starts = [k for k, l in enumerate(seq) if "nchw_to_" in l["kernel"]]
# the last start that is followed by a gap_softmax launch is the last complete step
step = None
for s in reversed(starts):
    end = next((k for k in range(s + 1, len(seq)) if "gap_softmax" in seq[k]["kernel"]), None)
    nxt = next((k for k in starts if k > s), len(seq))
    if end is not None and end < nxt:
        step = seq[s:end + 1]
        break
if step is None:
    sys.exit("no complete step found")
conv = [l for l in step if "conv_tc_kernel" in l["kernel"] and not l["kernel"].endswith("POOL")]
pooled = [l for l in step if l["kernel"].endswith("POOL")]
out = {
    "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on "
           "bench.py --steps 2 --warmup 3; last complete step (tools/ncu_step_traffic.py)",
    "units": {"dram_read": "byte", "dram_write": "byte", "time": "ns"},
    "launches": step,
    "conv_tc_launches": len(conv),
    "conv_tc_dram_bytes_per_step": sum(l["dram_read"] + l["dram_write"] for l in conv),
    "conv_tc_time_per_step": sum(l["time"] for l in conv),
    "pool_conv_tc_launches": len(pooled),
    "pool_conv_tc_dram_bytes_per_step": sum(l["dram_read"] + l["dram_write"] for l in pooled),
    "pool_conv_tc_time_per_step": sum(l["time"] for l in pooled),
    "step_time": sum(l["time"] for l in step),
}
json.dump(out, sys.stdout, indent=1)
