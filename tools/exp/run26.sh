#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python tools/mnist_bench.py 65536 > gpurun_out/mnist_65536.json 2>&1; cut -c1-500 gpurun_out/mnist_65536.json
B200_NO_DIRECT4=1 timeout -s KILL 300 python tools/mnist_bench.py 65536 > gpurun_out/mnist_65536_no4.json 2>&1; cut -c1-500 gpurun_out/mnist_65536_no4.json
