#!/bin/bash
# gpurun --gpus 4 -- bash scripts/gpu_multi4.sh : 4-rank parity tests (ranks with two neighbours) + the 4-GPU bench lines
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_multi_4.txt
cat gpurun_out/pytest_multi_4.txt
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/scale_4.json 2> gpurun_out/scale_4.err
tail -c 1200 gpurun_out/scale_4.json; tail -3 gpurun_out/scale_4.err
