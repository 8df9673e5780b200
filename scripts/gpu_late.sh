#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 100 -c 1 -f -o gpurun_out/sweep_index_late \
    python scripts/profile_step.py --workload c3 --intensity index --iters 110 > gpurun_out/ncu_index_late.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 100 -c 1 -f -o gpurun_out/sweep_band_late \
    python scripts/profile_step.py --workload c3 --intensity f64_band --iters 110 > gpurun_out/ncu_band_late.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_index_full.csv \
      python scripts/profile_step.py --workload c3 --intensity index --iters 200 > gpurun_out/prof_index_full.log 2>&1
python bench.py --steps 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
cat gpurun_out/pytest_gpu.txt; cat gpurun_out/bench_c3.json | head -c 600
