#!/bin/bash
# One gpurun call: GPU tests, bench lines, ncu launch list + full capture of the stencil kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
python bench.py --workload c1 --steps 2 --warmup 1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
for mode in f64_dense f64_band index; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${mode}.csv \
      python scripts/profile_step.py --workload c3 --intensity $mode --iters 10 > gpurun_out/prof_${mode}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 4 -c 2 -f -o gpurun_out/decide_dense \
    python scripts/profile_step.py --workload c3 --intensity f64_dense --iters 8 > gpurun_out/ncu_dense.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 4 -c 2 -f -o gpurun_out/decide_band \
    python scripts/profile_step.py --workload c3 --intensity f64_band --iters 8 > gpurun_out/ncu_band.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cancel -s 4 -c 1 -f -o gpurun_out/cancel \
    python scripts/profile_step.py --workload c3 --intensity f64_band --iters 8 > gpurun_out/ncu_cancel.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
