#!/bin/bash
# gpurun --gpus 8 -- bash scripts/gpu_n8.sh : the 8-GPU bench lines -- C3 strong (the headline shape), C5 weak (2048x2048x128 per GPU),
# C4 strong
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
run() {  # name, extra args
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 \
      bench.py --gpus 8 --steps 3 --warmup 3 $3 > gpurun_out/$2.json 2> gpurun_out/$2.err
  tail -c 1400 gpurun_out/$2.json; tail -n 3 gpurun_out/$2.err
}
run 29511 scale_8 ""
run 29512 scale_c5_weak_8 "--workload c5 --scaling weak"
run 29513 scale_c4_8 "--workload c4"
