#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python scripts/debug_mask.py > gpurun_out/debug_mask.log 2>&1
tail -n 8 gpurun_out/debug_mask.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
cat gpurun_out/bench_c3.json; tail -n 3 gpurun_out/bench_c3.err
timeout 600 python scripts/bench_mask.py --cpu > gpurun_out/bench_mask_c2.json 2> gpurun_out/bench_mask_c2.err
cat gpurun_out/bench_mask_c2.json; tail -n 5 gpurun_out/bench_mask_c2.err
timeout 300 python scripts/bench_mask.py --shape 640x880x880 --reps 3 > gpurun_out/bench_mask_c3.json 2> gpurun_out/bench_mask_c3.err
cat gpurun_out/bench_mask_c3.json; tail -n 5 gpurun_out/bench_mask_c3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_mask_c3.csv \
    python scripts/bench_mask.py --shape 640x880x880 --reps 1 > gpurun_out/prof_mask.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_f64_dense.csv \
    python scripts/profile_step.py --workload c3 --intensity f64_dense --iters 10 > gpurun_out/prof_f64_dense.log 2>&1
