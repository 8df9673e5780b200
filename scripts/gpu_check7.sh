#!/bin/bash
# round 2, pass t: whole GPU suite, wide-row dense sweep A/B on C4-shaped volumes, strict-mode bench
set -x
mkdir -p gpurun_out
TAG=r2t
timeout 900 python -m pytest tests -m gpu -q -x -rs 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python scripts/sweep_time.py 128x1024x1024:10 1024x1024x1024:10 > gpurun_out/${TAG}_sweep_time_wide.jsonl 2> gpurun_out/${TAG}_sweep_time_wide.err
VRG_NO_WIDE=1 timeout 300 python scripts/sweep_time.py 128x1024x1024:10 1024x1024x1024:10 > gpurun_out/${TAG}_sweep_time_nowide.jsonl 2>> gpurun_out/${TAG}_sweep_time_wide.err
cut -c1-330 gpurun_out/${TAG}_sweep_time_wide.jsonl gpurun_out/${TAG}_sweep_time_nowide.jsonl
timeout 600 python scripts/bench_strict.py > gpurun_out/${TAG}_bench_strict.jsonl 2> gpurun_out/${TAG}_bench_strict.err
cat gpurun_out/${TAG}_bench_strict.jsonl; tail -n 3 gpurun_out/${TAG}_bench_strict.err
