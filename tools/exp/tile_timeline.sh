#!/bin/bash
# Tile timeline of CTA 0 (role-mask build, B200_TC_DEBUG=1024): clock64 stamps of the MMA warp and of epilogue warp 0 at the
# hand-over points of each tile -> where a tile's time goes in the one-accumulator-stage layers.
O=gpurun_out/timeline; mkdir -p $O; rm -f $O/*
D=$PWD/onnx_rusty_inference_engine_b200/lib/variants/libb200rt_dbg.so
for L in conv1 f2_fused f4_e3 f8_e3; do
  B200_TC_DEBUG=1024 B200RT_LIB=$D timeout 120 python tools/tc_bench.py $L > $O/$L.out 2> $O/$L.err
  # keep the last timeline block
  python - "$O/$L.err" "$O/$L.txt" <<'PY'
import sys
lines = open(sys.argv[1]).read().splitlines()
starts = [i for i, l in enumerate(lines) if l.startswith("tile timeline")]
open(sys.argv[2], "w").write("\n".join(lines[starts[-1]:]) + "\n" if starts else "no timeline\n")
PY
  cat $O/$L.out; head -60 $O/$L.txt
done
