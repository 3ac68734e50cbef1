"""Summarise an ncu report of conv_tc_kernel at SASS level: stall samples accumulated between marker instructions.
usage: ncu_sass.py report.ncu-rep <invocation-nr> [min_samples]"""
import csv, subprocess, sys, io
rep, inv = sys.argv[1], sys.argv[2]
mins = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:conv_tc:{inv}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot_s = sum(int(r[iS]) for r in data); tot_i = sum(int(r[iI]) for r in data)
print("total samples", tot_s, "total warp-inst", tot_i)
marks = ('UTCHMMA', 'UTMALDG', 'LDTM', 'STTM', 'LDG.E', 'STS.128', 'STG.E', 'UTCBAR', 'BAR.SYNC', 'SYNCS.ARRIVE', 'LDS.128', 'EXIT',
         'SYNCS.PHASECHK', 'FENCE', 'MEMBAR')
seg_s = seg_i = 0
agg = [0] * len(stall_cols)
for n, r in enumerate(data):
    seg_s += int(r[iS]); seg_i += int(r[iI])
    for j, c in enumerate(stall_cols): agg[j] += int(r[c] or 0)
    src = r[iSrc].strip()
    if any(m in src for m in marks):
        if seg_s >= mins:
            st = sorted(((v, hdr[stall_cols[j]][6:]) for j, v in enumerate(agg)), reverse=True)[:2]
            print(f"{n:5d} samples={seg_s:6d} ({100*seg_s/tot_s:4.1f}%) inst={seg_i:10d} exec={int(r[iI]):9d}  {src[:70]:70s} {st}")
        seg_s = seg_i = 0; agg = [0] * len(stall_cols)
