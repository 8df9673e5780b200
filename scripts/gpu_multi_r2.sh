#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_multi_r2.sh N TAG : multi-GPU parity tests, then the N-GPU bench exactly as the driver launches
# it (parity against the whole-volume oracle, the C5 weak slab, drop-in timing where applicable)
N=${1:-2}
TAG=${2:-r2}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_multi_$N.txt
cat gpurun_out/${TAG}_pytest_multi_$N.txt
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_scale_$N.json 2> gpurun_out/${TAG}_scale_$N.err
echo "bench wall seconds: $SECONDS"
tail -c 2500 gpurun_out/${TAG}_scale_$N.json; tail -n 3 gpurun_out/${TAG}_scale_$N.err
