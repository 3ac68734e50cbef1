#!/bin/bash
# 2-GPU refresh: CUDA shard-equivalence tests and the 2-rank bench line (gpurun --gpus 2)
O=gpurun_out/r2_multi2; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -3 > $O/pytest_sharded.log; cat $O/pytest_sharded.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
echo "bench 2 rc=$?"; cut -c1-200 $O/bench_2gpu.json
