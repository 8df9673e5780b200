#!/bin/bash
# gpurun --gpus 8 -- bash scripts/gpu_n8_r2.sh TAG : the 8-GPU bench exactly as the driver launches it (C3 strong, parity against the
# whole-volume oracle, the C5 weak slab under `weak`)
TAG=${1:-r2}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/${TAG}_scale_8.json 2> gpurun_out/${TAG}_scale_8.err
echo "bench wall seconds: $SECONDS"
tail -c 1800 gpurun_out/${TAG}_scale_8.json; tail -n 3 gpurun_out/${TAG}_scale_8.err
