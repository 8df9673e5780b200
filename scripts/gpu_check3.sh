#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout 900 python -m pytest tests -m gpu -q -x -rs 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
VRG_VERBOSE=1 timeout 300 python scripts/sweep_time.py 82x880x880:10 640x880x880:10 > gpurun_out/${TAG}_sweep_time.jsonl 2> gpurun_out/${TAG}_sweep_time.err
cat gpurun_out/${TAG}_sweep_time.jsonl; tail -n 5 gpurun_out/${TAG}_sweep_time.err
VRG_DENSE_ROWS=4 timeout 300 python scripts/sweep_time.py 640x880x880:10 130x2048x2048:10 > gpurun_out/${TAG}_sweep_time_r4.jsonl 2> gpurun_out/${TAG}_sweep_time_r4.err
VRG_DENSE_ROWS=8 timeout 300 python scripts/sweep_time.py 82x880x880:10 640x880x880:10 130x2048x2048:10 >> gpurun_out/${TAG}_sweep_time_r4.jsonl 2>> gpurun_out/${TAG}_sweep_time_r4.err
cat gpurun_out/${TAG}_sweep_time_r4.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_tail" -s 20 -c 1 -f -o gpurun_out/${TAG}_tail \
    python scripts/profile_step.py --shape 82x880x880 --intensity f64_dense --iters 30 > gpurun_out/${TAG}_n2.log 2>&1
