#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r2i}
timeout 900 python -m pytest tests -m gpu -q -x -rs 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
VRG_VERBOSE=1 timeout 300 python scripts/sweep_time.py 82x880x880:10 640x880x880:10 > gpurun_out/${TAG}_sweep_time.jsonl 2> gpurun_out/${TAG}_sweep_time.err
cat gpurun_out/${TAG}_sweep_time.jsonl | cut -c1-400; grep -v "tail phase" gpurun_out/${TAG}_sweep_time.err | tail -n 8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_c3.csv \
    python scripts/profile_step.py --workload c3 --intensity f64_dense --iters 6 > gpurun_out/${TAG}_l.log 2>&1
VRG_HIST_PRIVATE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_c3_private.csv \
    python scripts/profile_step.py --workload c3 --intensity f64_dense --iters 6 > gpurun_out/${TAG}_l2.log 2>&1
grep -h "k_init_hist\|k_scan" gpurun_out/${TAG}_launches_c3.csv gpurun_out/${TAG}_launches_c3_private.csv | cut -c1-300
