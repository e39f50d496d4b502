#!/bin/bash
# quick A/B + one ncu --set full capture (source-level) of the tree kernel in stub and default mode
TAG=${1:-f}
bash tools/r2_quick.sh $TAG
ncu --set full --clock-control none --import-source on -k regex:tree_wave -s 350 -c 2 -o gpurun_out/${TAG}_tree_stub -f python bench.py --net stub --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tree_wave -s 350 -c 2 -o gpurun_out/${TAG}_tree_default -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines > /dev/null 2>&1
ls -la gpurun_out/${TAG}_tree_*.ncu-rep
