#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python tools/mnist_bench.py 65536 > gpurun_out/mnist_65536.json 2>&1; cut -c1-600 gpurun_out/mnist_65536.json
B200_NO_DIRECT_CONV=1 timeout -s KILL 300 python tools/mnist_bench.py 65536 > gpurun_out/mnist_65536_nodirect.json 2>&1; cut -c1-600 gpurun_out/mnist_65536_nodirect.json
timeout -s KILL 600 python bench.py --profile-out gpurun_out/per_launch.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
