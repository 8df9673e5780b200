#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python scripts/debug_mask.py > gpurun_out/debug_mask.log 2>&1
cat gpurun_out/debug_mask.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_f64_dense.csv \
    python scripts/profile_step.py --workload c3 --intensity f64_dense --iters 10 > gpurun_out/prof_f64_dense.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_index.csv \
    python scripts/profile_step.py --workload c3 --intensity index --iters 10 > gpurun_out/prof_index.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
cat gpurun_out/bench_c3.json; tail -n 3 gpurun_out/bench_c3.err
