#!/bin/bash
# GPU box: the default bench line with roofline_tree_stub_16k, the ncu --set full capture of the 16 384-game stub launch it
# cites, and the per-warp trace of such a launch.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err
echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench_default.err
python tools/fmt_bench.py < gpurun_out/r2b_bench_default.json 2>/dev/null | head -60
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tree_wave -s 350 -c 2 -o gpurun_out/tree_stub16k_r2 -f \
    python bench.py --net stub --games 16384 --nodes 640 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu16k.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu16k.log | cut -c1-300
timeout 300 python tools/trace_wave.py 16384 stub 2>&1 | tail -16 > gpurun_out/r2_trace_stub16k.log
cat gpurun_out/r2_trace_stub16k.log
ls -la gpurun_out/*.ncu-rep
