#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r2q}
timeout 1200 python -m pytest tests -m gpu -q -rs 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python scripts/bench_continuous.py > gpurun_out/${TAG}_bench_continuous.json 2> gpurun_out/${TAG}_bench_continuous.err
cat gpurun_out/${TAG}_bench_continuous.json; tail -n 3 gpurun_out/${TAG}_bench_continuous.err
