"""Split the SASS of an ncu report into runs of instructions with (nearly) the same execution count and print, per run,
the executed warp instructions and the stall samples: shows which role / loop of a warp-specialised kernel spends the
issue slots.  usage: ncu_regions.py report.ncu-rep [kernel-regex=conv_tc] [invocation=1] [min_share_pct=1.0]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "conv_tc"
inv = sys.argv[3] if len(sys.argv) > 3 else "1"
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{kre}:{inv}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot_i = sum(int(r[iI]) for r in data); tot_s = sum(int(r[iS]) for r in data)
print(f"total warp instructions {tot_i}, samples {tot_s}")
runs = []; start = 0
for n in range(1, len(data) + 1):
    if n == len(data) or abs(int(data[n][iI]) - int(data[start][iI])) > 0.25 * max(int(data[start][iI]), 1):
        runs.append((start, n)); start = n
for a, b in runs:
    ins = sum(int(r[iI]) for r in data[a:b]); sm = sum(int(r[iS]) for r in data[a:b])
    if 100 * ins / tot_i >= minp or 100 * sm / tot_s >= minp:
        ops = {}
        for r in data[a:b]:
            op = r[iSrc].strip().split()[0 if not r[iSrc].strip().startswith('@') else 1].split('.')[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
        print(f"[{a:5d},{b:5d}) n={b-a:4d} exec/instr={int(data[a][iI]):9d} instr {100*ins/tot_i:5.1f}% samples {100*sm/tot_s:5.1f}%  {top}")
