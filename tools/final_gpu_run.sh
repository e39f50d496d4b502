#!/bin/bash
# Round-end evidence run on the GPU box: parity tests, smoke, the three bench modes, the reference arm, the ncu launch
# list and the ncu --set full captures the summaries under profiles/ are made from.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --net stub --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stub.json 2> gpurun_out/bench_stub.err
python bench.py --workload rules --steps 20 --warmup 3 > gpurun_out/bench_rules.json 2> gpurun_out/bench_rules.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/err.log
# launch list of the default command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 1400 --csv --log-file gpurun_out/launches_default_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
# full captures of the top kernels
ncu --set full --clock-control none --import-source on -k regex:tree_wave -s 350 -c 2 -o gpurun_out/tree_default_r1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tree_wave -s 350 -c 2 -o gpurun_out/tree_stub_r1 -f python bench.py --net stub --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rules_mask_step -s 3 -c 2 -o gpurun_out/rules_r1 -f python bench.py --workload rules --steps 3 --warmup 3 > gpurun_out/ncu4.log 2>&1
# the net's dominant kernels (library code): tensor-pipe utilisation of the 512-channel convolutions
ncu --set full --clock-control none -k regex:"conv3d_fprop|implicit_gemm_fprop" -s 60 -c 3 -o gpurun_out/net_convs_r1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu5.log 2>&1
# BASELINE configs[3] per-GPU shape (4096 games x 200 sims/move): runs, faults would raise
python bench.py --sims 200 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sims200.json 2> gpurun_out/bench_sims200.err
python tools/trace_wave.py 4096 default 2>&1 | tail -12 > gpurun_out/trace_default.log
python tools/trace_wave.py 4096 stub 2>&1 | tail -16 > gpurun_out/trace_stub.log
for f in default stub rules reference sims200; do python tools/fmt_bench.py < gpurun_out/bench_$f.json 2>/dev/null || cut -c1-600 gpurun_out/bench_$f.json; done
