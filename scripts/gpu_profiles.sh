#!/bin/bash
# final evidence pass: ncu launch list of the bench command + full captures of the main kernels
set -x
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_dense -s 6 -c 2 -f -o gpurun_out/final_sweep_dense \
    python scripts/profile_step.py --workload c3 --intensity f64_dense --iters 12 > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_band -s 40 -c 1 -f -o gpurun_out/final_sweep_index \
    python scripts/profile_step.py --workload c3 --intensity index --iters 50 > gpurun_out/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cancel -s 40 -c 1 -f -o gpurun_out/final_cancel \
    python scripts/profile_step.py --workload c3 --intensity index --iters 50 > gpurun_out/ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_table -s 40 -c 1 -f -o gpurun_out/final_table \
    python scripts/profile_step.py --workload c3 --intensity index --iters 50 > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out | tail -12
