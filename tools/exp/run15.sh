#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' >> gpurun_out/sweep15.txt 2>&1
import re
s=open('tools/tc_bench.py').read()
s=s.replace('"conv10": (256, 512, 13, 13, 1000, 1, 1, 0),','"conv10": (256, 512, 13, 13, 1000, 1, 1, 0),\n    "f6_e3": (256, 48, 27, 27, 192, 3, 1, 1),\n    "f6_e1": (256, 48, 27, 27, 192, 1, 1, 0),')
open('/tmp/tc_bench2.py','w').write(s.replace("os.path.dirname(os.path.dirname(os.path.abspath(__file__)))", "'/root/repo'"))
PY
for nacc in 0 1; do for m in 0 35; do
echo "== NACC=$nacc mask $m" >> gpurun_out/sweep15.txt
B200_TC_NACC=$nacc B200_TC_DEBUG=$m timeout -s KILL 300 python /tmp/tc_bench2.py conv1 f6_e3 f6_e1 f2_e3 >> gpurun_out/sweep15.txt 2>&1
done; done
cat gpurun_out/sweep15.txt
