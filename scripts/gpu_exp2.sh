#!/bin/bash
# GPU tests (incl. the mask-side operations), mask bench with SciPy beside it, launch list of a short dense run, C5/8 on one GPU
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 600 python scripts/bench_mask.py --cpu > gpurun_out/bench_mask_c2.json 2> gpurun_out/bench_mask_c2.err
cat gpurun_out/bench_mask_c2.json; tail -n 5 gpurun_out/bench_mask_c2.err
timeout 300 python scripts/bench_mask.py --shape 640x880x880 --reps 3 > gpurun_out/bench_mask_c3.json 2> gpurun_out/bench_mask_c3.err
cat gpurun_out/bench_mask_c3.json; tail -n 5 gpurun_out/bench_mask_c3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_f64_dense.csv \
    python scripts/profile_step.py --workload c3 --intensity f64_dense --iters 10 > gpurun_out/prof_f64_dense.log 2>&1
timeout 300 python bench.py --workload c5 --scaling weak --steps 3 --warmup 3 > gpurun_out/bench_c5_weak1.json 2> gpurun_out/bench_c5_weak1.err
cat gpurun_out/bench_c5_weak1.json; tail -n 3 gpurun_out/bench_c5_weak1.err
