#!/bin/bash
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 600 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 examples/coach_iteration.py --games 32 --sims 25 --arena 16 --epochs 1 --max-turns 40 --gather > gpurun_out/coach_n2.log 2>&1
tail -5 gpurun_out/coach_n2.log
cut -c1-400 gpurun_out/bench_n2.json
