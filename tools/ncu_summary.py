"""Write a text summary of an `ncu --set full` report of conv_tc_kernel: key raw metrics + top stall sites (SASS).
usage: ncu_summary.py report.ncu-rep [invocation=1] > profiles/xxx_ncu_summary.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]; inv = sys.argv[2] if len(sys.argv) > 2 else "1"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[1 + int(inv)]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# ncu --set full --clock-control none  ({rep}, invocation {inv})")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:72s} {r[i][:70]} {units[i]}")
print("\n# top stall sites (SASS, stall samples accumulated up to each marker instruction)")
out = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_sass.py"), rep, inv, "150"], capture_output=True, text=True).stdout
print(out)
