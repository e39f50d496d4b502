#!/bin/bash
# 2 GPUs: the same small Coach iteration with SyncBatchNorm (default) and with per-rank BatchNorm statistics (--local-bn).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for v in "" "--local-bn"; do
  timeout 300 $TR --master-port 29531 examples/coach_iteration.py --games 48 --sims 25 --arena 8 --epochs 1 --compact $v > gpurun_out/r2_coach_n2_bn$v.log 2>&1
  echo "rc=$? $v"; grep "^train" gpurun_out/r2_coach_n2_bn$v.log | cut -c1-260
done
