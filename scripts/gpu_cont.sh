#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 300 python scripts/bench_continuous.py > gpurun_out/bench_continuous.json 2> gpurun_out/bench_continuous.err
cat gpurun_out/bench_continuous.json; tail -3 gpurun_out/bench_continuous.err
timeout 400 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
head -c 600 gpurun_out/bench_c3.json
