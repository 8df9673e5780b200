#!/bin/bash
# gpurun --gpus 8 -- bash scripts/gpu_n8_ab.sh TAG : 8-GPU C3 line on the production path and with one switch flipped at a time
TAG=${1:-r2}
set -x
mkdir -p gpurun_out
run() {  # port, name, env
  env $3 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 \
      bench.py --gpus 8 --steps 3 --warmup 3 --quick --no-parity > gpurun_out/${TAG}_scale_8_$2.json 2> gpurun_out/${TAG}_scale_8_$2.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${TAG}_scale_8_$2.json") if l.startswith("{")][-1])
r = d["roofline"]
print("$2", "value %.0f ms_per_step %.2f sweep_ms %.4f tail_ms %.4f split %s launches/step %.0f" % (d["value"], d["ms_per_step"], r["ms_per_launch"], r.get("tail_ms_per_launch") or 0, r.get("step_split_ms"), d["gpu_launches"] / d["steps"]))
PY
}
L=/root/repo/arterynetwork_b200/csrc
if [ "$2" = "libs" ]; then
  run 29541 async256 VRG_B200_LIB=$L/libvrg_alt_B.so
  run 29542 reg96 VRG_B200_LIB=$L/libvrg_alt_E.so
  run 29543 reg96_async256 VRG_B200_LIB=$L/libvrg_alt_F.so
else
  run 29541 default VRG_X=1
  run 29542 noahead VRG_NO_AHEAD=1
  run 29543 inorder VRG_PIPELINE=0
fi
