#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_n.sh N : only the N-GPU bench lines (dense + index)
N=${1:-8}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
tail -c 1300 gpurun_out/scale_$N.json; tail -3 gpurun_out/scale_$N.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 3 --warmup 3 --workload c5 > gpurun_out/scale_c5_$N.json 2> gpurun_out/scale_c5_$N.err
tail -c 1300 gpurun_out/scale_c5_$N.json; tail -3 gpurun_out/scale_c5_$N.err
