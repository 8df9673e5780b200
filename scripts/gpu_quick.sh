#!/bin/bash
# Shorter gpurun call: GPU tests, the C3 bench line, full ncu captures of the two sweep kernels and k_cancel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 4 -c 2 -f -o gpurun_out/sweep_dense \
    python scripts/profile_step.py --workload c3 --intensity f64_dense --iters 8 > gpurun_out/ncu_dense.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 6 -c 1 -f -o gpurun_out/sweep_index \
    python scripts/profile_step.py --workload c3 --intensity index --iters 8 > gpurun_out/ncu_index.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cancel -s 4 -c 1 -f -o gpurun_out/cancel \
    python scripts/profile_step.py --workload c3 --intensity index --iters 8 > gpurun_out/ncu_cancel.log 2>&1
tail -3 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
