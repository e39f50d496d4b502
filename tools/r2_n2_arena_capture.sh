#!/bin/bash
# 2 GPUs: a small Coach iteration under torchrun -- the Arena of the iteration now replays a captured (net, wave) step
# (BatchedArena(capture=True)) with an NCCL process group alive.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    examples/coach_iteration.py --games 64 --sims 50 --arena 40 --epochs 1 --compact > gpurun_out/r2_coach_n2_capture.log 2>&1
echo "rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_coach_n2_capture.log | tail -12 | cut -c1-300
