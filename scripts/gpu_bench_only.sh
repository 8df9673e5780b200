#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 400 python bench.py --steps 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
head -c 700 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
