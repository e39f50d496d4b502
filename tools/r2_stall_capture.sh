#!/bin/bash
# stall samples per source line of the shipped tree kernel, 24 launches per mode (tools/stall_profile.py reads these);
# kept apart from tools/r2_evidence.sh because gpurun brings back at most 64 MiB per call
T=r2
mkdir -p gpurun_out
ncu --section SourceCounters --section WarpStateStats --clock-control none --import-source on -k regex:tree_wave -s 350 -c 24 -o gpurun_out/stall_default_${T} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines > gpurun_out/ncu6.log 2>&1
ncu --section SourceCounters --section WarpStateStats --clock-control none --import-source on -k regex:tree_wave -s 350 -c 24 -o gpurun_out/stall_stub_${T} -f python bench.py --net stub --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu7.log 2>&1
ls -la gpurun_out/
