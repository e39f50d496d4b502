#!/bin/bash
# Round 2, run A: validate the new bench contract on hardware and take fresh ncu captures of the tree kernel as shipped at
# the end of round 1 (the "before" of this round's kernel work).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 8 --warmup 3 > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_reference.err
python bench.py --net stub --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_stub.json 2> gpurun_out/r2a_bench_stub.err
ncu --set full --clock-control none --import-source on -k regex:tree_wave -s 350 -c 2 -o gpurun_out/r2a_tree_default -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines > gpurun_out/ncu_a2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tree_wave -s 350 -c 2 -o gpurun_out/r2a_tree_stub -f python bench.py --net stub --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_a3.log 2>&1
for f in default reference stub; do cut -c1-1500 gpurun_out/r2a_bench_$f.json; tail -3 gpurun_out/r2a_bench_$f.err; done
