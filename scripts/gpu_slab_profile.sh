#!/bin/bash
# One B200: timings and ncu captures of a slab-sized run (82 x 880 x 880 = one rank's share of C3 on 8 GPUs) -- the proxy on which the
# per-iteration overhead of the 8-GPU strong-scaling run is tuned.
set -x
mkdir -p gpurun_out
TAG=${1:-slab}
timeout 300 python scripts/sweep_time.py 82x880x880:10 640x880x880:10 128x1024x1024:10 > gpurun_out/${TAG}_sweep_time.jsonl 2> gpurun_out/${TAG}_sweep_time.err
cat gpurun_out/${TAG}_sweep_time.jsonl; tail -n 3 gpurun_out/${TAG}_sweep_time.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/profile_step.py --shape 82x880x880 --intensity f64_dense --iters 40 > gpurun_out/${TAG}_l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_dense -s 20 -c 1 -f -o gpurun_out/${TAG}_sweep_dense \
    python scripts/profile_step.py --shape 82x880x880 --intensity f64_dense --iters 30 > gpurun_out/${TAG}_n1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_cancel|k_tail" -s 20 -c 1 -f -o gpurun_out/${TAG}_cancel \
    python scripts/profile_step.py --shape 82x880x880 --intensity f64_dense --iters 30 > gpurun_out/${TAG}_n2.log 2>&1
ls -la gpurun_out | tail
