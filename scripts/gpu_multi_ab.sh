#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_multi_ab.sh N [TAG] : multi-GPU parity tests, then the N-GPU bench on the production path
# (fused tail kernel) and, for comparison, on the separate kernels (VRG_NO_FUSED_TAIL=1: halo exchange on a second stream)
N=${1:-2}
TAG=${2:-r2}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_multi_$N.txt
cat gpurun_out/${TAG}_pytest_multi_$N.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_scale_$N.json 2> gpurun_out/${TAG}_scale_$N.err
tail -c 2500 gpurun_out/${TAG}_scale_$N.json; tail -n 5 gpurun_out/${TAG}_scale_$N.err
VRG_NO_FUSED_TAIL=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 3 --warmup 3 --quick --no-parity > gpurun_out/${TAG}_scale_${N}_separate.json 2> gpurun_out/${TAG}_scale_${N}_separate.err
tail -c 1200 gpurun_out/${TAG}_scale_${N}_separate.json; tail -n 3 gpurun_out/${TAG}_scale_${N}_separate.err
