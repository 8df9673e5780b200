#!/bin/bash
# quick single-GPU check: the whole GPU test suite + the default bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
head -c 700 gpurun_out/bench_c3.json; tail -n 3 gpurun_out/bench_c3.err
