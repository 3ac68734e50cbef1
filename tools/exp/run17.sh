#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/mnist_bench.py 8192 > gpurun_out/mnist_8192.json 2>&1; cut -c1-700 gpurun_out/mnist_8192.json
timeout -s KILL 300 python tools/mnist_bench.py 65536 > gpurun_out/mnist_65536.json 2>&1; cut -c1-700 gpurun_out/mnist_65536.json
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu rc=$?"; cut -c1-600 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
