"""Print a b200_model_profile JSON (bench.py --profile-out) as a per-launch table; optional second file to compare."""
import json, sys
p = json.load(open(sys.argv[1]))
q = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else None
for i, x in enumerate(p):
    tf = x['flops'] / x['ms'] / 1e9 if x['ms'] > 0 else 0
    gb = x['bytes'] / x['ms'] / 1e6 if x['ms'] > 0 else 0
    was = f" (was {q[i]['ms']:6.3f})" if q else ""
    print(f"{x['name'][:30]:30s} {x['kind']:11s} {x['ms']:7.3f} ms{was} {tf:7.2f} TF/s {gb:8.1f} GB/s")
print("total ms", sum(x['ms'] for x in p), (" was %.3f" % sum(x['ms'] for x in q)) if q else "")
