#!/bin/bash
# Round-2 evidence run on one B200: parity tests, smoke, the bench lines the driver will take (default / reference) plus
# stub, rules and the 200-sim shape, the ncu launch list and the ncu --set full captures the summaries under profiles/
# are made from (tools/summarize_profiles.py r2), per-warp traces.
set -x
T=r2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
python bench.py --net stub --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_stub.json 2> gpurun_out/${T}_bench_stub.err
python bench.py --workload rules --steps 20 --warmup 3 > gpurun_out/${T}_bench_rules.json 2> gpurun_out/${T}_bench_rules.err
python bench.py --sims 200 --steps 4 --warmup 3 --no-cpu-baseline --no-extra-rooflines > gpurun_out/${T}_bench_sims200.json 2> gpurun_out/${T}_bench_sims200.err
# launch list of the default command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 1400 --csv --log-file gpurun_out/launches_default_${T}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines > gpurun_out/ncu1.log 2>&1
# full captures of our kernels
ncu --set full --clock-control none --import-source on -k regex:tree_wave -s 350 -c 2 -o gpurun_out/tree_default_${T} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tree_wave -s 350 -c 2 -o gpurun_out/tree_stub_${T} -f python bench.py --net stub --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rules_mask_step -s 3 -c 2 -o gpurun_out/rules_${T} -f python bench.py --workload rules --steps 3 --warmup 3 > gpurun_out/ncu4.log 2>&1
ncu --set full --clock-control none -k regex:"conv3d_fprop|implicit_gemm_fprop" -s 60 -c 3 -o gpurun_out/net_convs_${T} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines > gpurun_out/ncu5.log 2>&1
python tools/trace_wave.py 4096 default 2>&1 | tail -12 > gpurun_out/${T}_trace_default.log
python tools/trace_wave.py 4096 stub 2>&1 | tail -16 > gpurun_out/${T}_trace_stub.log
for f in default stub reference sims200; do python tools/fmt_bench.py < gpurun_out/${T}_bench_$f.json 2>/dev/null || cut -c1-600 gpurun_out/${T}_bench_$f.json; tail -2 gpurun_out/${T}_bench_$f.err; done
cut -c1-700 gpurun_out/${T}_bench_rules.json
