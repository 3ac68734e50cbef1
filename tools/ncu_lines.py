"""Summarise an ncu report per CUDA source line: stall samples and executed instructions for one kernel invocation.
usage: ncu_lines.py report.ncu-rep <invocation-nr> [top_n]"""
import csv, subprocess, sys, io
rep, inv = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda", "--kernel-id", f"::regex:conv_tc:{inv}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# find header row
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
iLine = hdr.index("Line") if "Line" in hdr else None
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot_s = sum(int(r[iS] or 0) for r in data); tot_i = sum(int(r[iI] or 0) for r in data)
print(rows[0][:2] if hi > 0 else "", "total samples", tot_s, "total warp-inst", tot_i)
for r in sorted(data, key=lambda r: -int(r[iS] or 0))[:topn]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    ln = r[iLine] if iLine is not None else "?"
    print(f"{int(r[iS] or 0):6d} {100*int(r[iS] or 0)/max(tot_s,1):5.1f}% inst={int(r[iI] or 0):9d}  L{ln:>4s} {r[iSrc].strip()[:90]:90s} {st}")
