#!/bin/bash
# 2 GPUs: both bench arms under torchrun, as the driver's scaling run launches them.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --impl reference --gpus 2 --steps 8 --warmup 3 > gpurun_out/final_n2_reference.json 2> gpurun_out/final_n2_reference.err; echo "reference rc=$?"
timeout 400 $TR --master-port 29522 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/final_n2_default.json 2> gpurun_out/final_n2_default.err; echo "default rc=$?"
grep -h '^{' gpurun_out/final_n2_reference.json | cut -c1-300
grep -h '^{' gpurun_out/final_n2_default.json | python tools/fmt_bench.py
