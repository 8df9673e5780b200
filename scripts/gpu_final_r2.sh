#!/bin/bash
# One B200: the round's final evidence pass (TAG = file prefix).  Tests, the default bench line, the reference arm, the other
# single-GPU configs, mask-side / continuous / strict benches, the ncu launch list of the bench command and a full capture of the sweep.
TAG=${1:-r2z}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
SECONDS=0
timeout 600 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
echo "default bench wall seconds: $SECONDS"
head -c 900 gpurun_out/${TAG}_bench_c3.json; echo; tail -n 2 gpurun_out/${TAG}_bench_c3.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
cat gpurun_out/${TAG}_bench_ref.json
for w in c4 c2 c1; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --quick > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  head -c 300 gpurun_out/${TAG}_bench_$w.json; echo; tail -n 1 gpurun_out/${TAG}_bench_$w.err
done
timeout 300 python scripts/bench_mask.py --shape 640x880x880 --reps 3 > gpurun_out/${TAG}_bench_mask_c3.json 2> gpurun_out/${TAG}_bench_mask_c3.err
cat gpurun_out/${TAG}_bench_mask_c3.json
timeout 300 python scripts/bench_continuous.py > gpurun_out/${TAG}_bench_continuous.json 2> gpurun_out/${TAG}_bench_continuous.err
cat gpurun_out/${TAG}_bench_continuous.json
timeout 300 python scripts/bench_strict.py > gpurun_out/${TAG}_bench_strict.jsonl 2> gpurun_out/${TAG}_bench_strict.err
cut -c1-300 gpurun_out/${TAG}_bench_strict.jsonl
timeout 300 python scripts/dropin_time.py 2>&1 | grep -v Warning | tail -3 > gpurun_out/${TAG}_dropin_time.txt
cat gpurun_out/${TAG}_dropin_time.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --quick --no-parity > gpurun_out/${TAG}_bench_under_ncu.json 2> gpurun_out/${TAG}_bench_under_ncu.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_dense -s 6 -c 2 -f -o gpurun_out/${TAG}_sweep_dense \
    python scripts/profile_step.py --workload c3 --intensity f64_dense --iters 12 > gpurun_out/${TAG}_ncu1.log 2>&1
ls -la gpurun_out | tail -5
