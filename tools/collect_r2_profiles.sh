#!/bin/bash
# Copies the evidence brought back by tools/exp/artifacts_r2.sh (gpurun_out/artifacts_r2) into profiles/r2_* and derives
# the DRAM-traffic file (with the hash of the kernel sources) and the ncu summaries.  Run here (no GPU needed).
set -e
A=gpurun_out/artifacts_r2
python tools/ncu_step_traffic.py $A/launches.csv > gpurun_out/r2_traffic.json
python - <<'PY'
import json, sys
sys.path.insert(0, '.')
import bench
d = json.load(open('gpurun_out/r2_traffic.json'))
d['kernel_sources_sha'] = bench.kernel_sources_sha()
d['kernel_sources'] = list(bench.KERNEL_SOURCES)
json.dump(d, open('profiles/r2_step_dram_traffic.json', 'w'), indent=1)
print('traffic: conv launches', d['conv_tc_launches'], 'DRAM GB per step', d['conv_tc_dram_bytes_per_step'] / 1e9, 'sha', d['kernel_sources_sha'])
PY
cp $A/launches.csv profiles/r2_launches.csv
cp $A/bench_default.json profiles/r2_bench_default_100steps.json
cp $A/bench_20steps.json profiles/r2_bench_20steps.json
cp $A/bench_reference.json profiles/r2_bench_reference_arm.json
cp $A/per_launch_events.json profiles/r2_per_launch_events.json
python tools/ncu_summary.py $A/prof_conv1.ncu-rep > profiles/r2_conv1_ncu_summary.txt 2>&1
python tools/ncu_summary.py $A/prof_f2_fused.ncu-rep > profiles/r2_f2_fused_ncu_summary.txt 2>&1
python tools/ncu_summary.py $A/prof_pool1_sq.ncu-rep > profiles/r2_pool1_squeeze_ncu_summary.txt 2>&1
ncu -i $A/prof_mnist.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/mn2.csv
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/mn2.csv')))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = []
for r in rows[2:]:
    out.append('-----')
    for w in want:
        if w in hdr:
            i = hdr.index(w); out.append(f"{w:70s} {r[i][:70]} {units[i]}")
open('profiles/r2_mnist8_ncu_summary.txt', 'w').write("# ncu --set full --clock-control none --import-source on -k regex:mnist8 (tools/mnist_bench.py 65536): two invocations of the one-launch kernel (fused_cnn = 2)\n" + "\n".join(out) + "\n")
print("\n".join(out))
PY
