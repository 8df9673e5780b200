#!/bin/bash
# tests + bench + dense-sweep A/B timings on slab-sized volumes + launch list of a slab-sized run
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
tail -3 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
cat gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python scripts/sweep_time.py 640x880x880:7,10,6 82x880x880:7,10 130x2048x2048:8,10,6 128x1024x1024:8,10 > gpurun_out/sweep_time.jsonl 2> gpurun_out/sweep_time.err
cat gpurun_out/sweep_time.jsonl; tail -3 gpurun_out/sweep_time.err
timeout 300 python bench.py --workload c5 --scaling weak --steps 3 --warmup 3 > gpurun_out/bench_c5_weak1.json 2> gpurun_out/bench_c5_weak1.err
cat gpurun_out/bench_c5_weak1.json; tail -3 gpurun_out/bench_c5_weak1.err
