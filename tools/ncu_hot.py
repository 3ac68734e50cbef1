"""List the hottest SASS instructions of an ncu report (stall samples per instruction, top stall reasons).
usage: ncu_hot.py report.ncu-rep [kernel-regex=conv_tc] [invocation=1] [min_samples=40]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "conv_tc"
inv = sys.argv[3] if len(sys.argv) > 3 else "1"
mins = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{kre}:{inv}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(int(r[iS]) for r in data)
print("total samples", tot, "instructions", len(data))
for n, r in enumerate(data):
    sm = int(r[iS])
    if sm >= mins:
        st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{n:5d} {sm:6d} {100*sm/tot:5.1f}% exec={int(r[iI]):9d} {r[iSrc].strip()[:80]:80s} {st}")
