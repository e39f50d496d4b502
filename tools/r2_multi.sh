#!/bin/bash
# Round 2 multi-GPU evidence: usage tools/r2_multi.sh N [games_coach sims_coach arena]
N=${1:-2}; CG=${2:-64}; CS=${3:-100}; AR=${4:-100}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
set -x
$TR --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -c 400 gpurun_out/r2_bench_n$N.err
$TR --master-port 29512 bench.py --gpus $N --sims 200 --steps 4 --warmup 3 > gpurun_out/r2_bench_sims200_n$N.json 2> gpurun_out/r2_bench_sims200_n$N.err; tail -c 400 gpurun_out/r2_bench_sims200_n$N.err
$TR --master-port 29513 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n$N.json 2> gpurun_out/r2_bench_ref_n$N.err; tail -c 400 gpurun_out/r2_bench_ref_n$N.err
timeout 1200 $TR --master-port 29514 examples/coach_iteration.py --games $CG --sims $CS --arena $AR --epochs 1 --compact > gpurun_out/r2_coach_n$N.log 2>&1
tail -12 gpurun_out/r2_coach_n$N.log
for f in r2_bench_n$N r2_bench_sims200_n$N r2_bench_ref_n$N; do cut -c1-330 gpurun_out/$f.json; done
