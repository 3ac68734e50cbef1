#!/bin/bash
# Round-2 multi-GPU evidence in ONE gpurun --gpus 8 call: concurrent-rank H2D ceiling, CUDA shard-equivalence tests,
# bench.py at 2 / 4 / 8 ranks.  Outputs in gpurun_out/r2_multi/.
O=gpurun_out/r2_multi; mkdir -p $O
if [ "$1" != "nocopy" ]; then timeout 240 tools/exp/build/h2d_concurrent > $O/h2d_concurrent.txt 2>&1; echo "h2d rc=$?"; fi
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -5 > $O/pytest_sharded.log; cat $O/pytest_sharded.log
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 > $O/bench_${n}gpu.json 2> $O/bench_${n}gpu.err
  echo "bench $n rc=$?"; cut -c1-200 $O/bench_${n}gpu.json
done
if [ "$1" != "nocopy" ]; then tail -40 $O/h2d_concurrent.txt; fi
