#!/bin/bash
# one B200: GPU tests, slab / full-size timings with the fused tail on and off, short bench
set -x
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 900 python -m pytest tests -m gpu -q -x -rs 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
VRG_VERBOSE=1 timeout 300 python scripts/sweep_time.py 82x880x880:10 640x880x880:10 130x2048x2048:10 > gpurun_out/${TAG}_sweep_time.jsonl 2> gpurun_out/${TAG}_sweep_time.err
cat gpurun_out/${TAG}_sweep_time.jsonl; tail -n 5 gpurun_out/${TAG}_sweep_time.err
VRG_NO_FUSED_TAIL=1 timeout 300 python scripts/sweep_time.py 82x880x880:10 640x880x880:10 > gpurun_out/${TAG}_sweep_time_legacy.jsonl 2> gpurun_out/${TAG}_sweep_time_legacy.err
cat gpurun_out/${TAG}_sweep_time_legacy.jsonl; tail -n 5 gpurun_out/${TAG}_sweep_time_legacy.err
VRG_DENSE_ROWS=16 timeout 300 python scripts/sweep_time.py 82x880x880:10 > gpurun_out/${TAG}_sweep_time_r16.jsonl 2> gpurun_out/${TAG}_sweep_time_r16.err
VRG_DENSE_ROWS=8 timeout 300 python scripts/sweep_time.py 82x880x880:10 640x880x880:10 >> gpurun_out/${TAG}_sweep_time_r16.jsonl 2>> gpurun_out/${TAG}_sweep_time_r16.err
VRG_DENSE_ROWS=2 timeout 300 python scripts/sweep_time.py 82x880x880:10 >> gpurun_out/${TAG}_sweep_time_r16.jsonl 2>> gpurun_out/${TAG}_sweep_time_r16.err
cat gpurun_out/${TAG}_sweep_time_r16.jsonl
timeout 600 python bench.py --quick > gpurun_out/${TAG}_bench_quick.json 2> gpurun_out/${TAG}_bench_quick.err
head -c 1500 gpurun_out/${TAG}_bench_quick.json; tail -n 3 gpurun_out/${TAG}_bench_quick.err
