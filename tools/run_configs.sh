#!/bin/bash
# BASELINE configs 3 / 4 / 5 on the GPUs of this box (harness/train_step.py, harness/sweep.py); one JSON line per run is
# appended to $OUT (default gpurun_out/configs.jsonl).  Usage: tools/run_configs.sh "1 2 4 8"   (GPU counts to run)
set -u
cd "$(dirname "$0")/.."
OUT=${OUT:-gpurun_out/configs.jsonl}
mkdir -p "$(dirname "$OUT")"
NS=${1:-1}
port=29511
run() {  # run <n_gpus> <script> <args...>
  n=$1; shift
  port=$((port + 1))
  if [ "$n" = "1" ]; then
    timeout 600 python "$@" --out "$OUT" || echo "{\"failed\": \"$*\", \"n_gpus\": 1}" >> "$OUT"
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $port \
      "$@" --out "$OUT" || echo "{\"failed\": \"$*\", \"n_gpus\": $n}" >> "$OUT"
  fi
}
for n in $NS; do
  if [ "$n" = "1" ]; then
    # config 3: full train step, batch 128, one GPU
    run 1 harness/train_step.py --batch 128 --epochs 2 --steps-per-epoch 12 --val-batches 2
    run 1 harness/train_step.py --batch 128 --epochs 2 --steps-per-epoch 12 --val-batches 2 --loss b200-heads
  fi
  # config 4: DDP, global batch 1024, 64^3 grid
  run "$n" harness/train_step.py --batch 1024 --epochs 2 --steps-per-epoch 6 --val-batches 1
  # config 5: 128^3 sweep over 4096 pairs
  run "$n" harness/sweep.py --pairs 4096
done
cat "$OUT"
