#!/bin/bash
# one GPU: full GPU test suite + the default bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
timeout 400 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
cat gpurun_out/pytest_gpu.txt; cat gpurun_out/bench_c3.json | head -c 1500; tail -3 gpurun_out/bench_c3.err
