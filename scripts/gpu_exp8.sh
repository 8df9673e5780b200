#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 600 python scripts/bench_mask.py > gpurun_out/bench_mask_c2.json 2> gpurun_out/bench_mask_c2.err
cat gpurun_out/bench_mask_c2.json; tail -n 5 gpurun_out/bench_mask_c2.err
timeout 300 python scripts/bench_mask.py --shape 640x880x880 --reps 3 > gpurun_out/bench_mask_c3.json 2> gpurun_out/bench_mask_c3.err
cat gpurun_out/bench_mask_c3.json; tail -n 5 gpurun_out/bench_mask_c3.err
