#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_multi_ab.sh N : multi-GPU parity tests, then the N-GPU bench with the halo exchange
# overlapped (default) and serial (VRG_P2P_SERIAL=1)
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_multi_$N.txt
cat gpurun_out/pytest_multi_$N.txt
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
tail -c 1500 gpurun_out/scale_$N.json; tail -n 3 gpurun_out/scale_$N.err
VRG_P2P_SERIAL=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/scale_${N}_serial.json 2> gpurun_out/scale_${N}_serial.err
tail -c 1500 gpurun_out/scale_${N}_serial.json; tail -n 3 gpurun_out/scale_${N}_serial.err
