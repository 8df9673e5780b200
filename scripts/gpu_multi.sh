#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_multi.sh N : NCCL parity tests + the scaling bench at N GPUs
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_multi_$N.txt
cat gpurun_out/pytest_multi_$N.txt
for n in $(seq 1 $N); do
  if [ $n -eq 1 ] || [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; then
    if [ $n -eq 1 ]; then
      timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
    else
      timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
      timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
        bench.py --gpus $n --steps 3 --warmup 3 --intensity index > gpurun_out/scale_index_$n.json 2> gpurun_out/scale_index_$n.err
    fi
    tail -c 1500 gpurun_out/scale_$n.json; tail -3 gpurun_out/scale_$n.err
  fi
done
