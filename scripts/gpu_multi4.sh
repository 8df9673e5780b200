#!/bin/bash
# gpurun --gpus 4 -- bash scripts/gpu_multi4.sh : multi-GPU parity tests on 4 ranks (ranks with two neighbours) + C3 / C4 bench lines
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_multi_4.txt
cat gpurun_out/pytest_multi_4.txt
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/scale_4.json 2> gpurun_out/scale_4.err
tail -c 1500 gpurun_out/scale_4.json; tail -n 3 gpurun_out/scale_4.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 4 --steps 3 --warmup 3 --workload c4 > gpurun_out/scale_c4_4.json 2> gpurun_out/scale_c4_4.err
tail -c 1500 gpurun_out/scale_c4_4.json; tail -n 3 gpurun_out/scale_c4_4.err
