#!/bin/bash
# dev tool (GPU box): ncu per-launch duration (median over 200 launches) + event-timed bench line, per library variant
P=alpha-zero-general-inflexion_b200
cp $P/libifx_b200.so /tmp/orig.so
for so in "$@"; do
  cp $so $P/libifx_b200.so
  n=$(basename $so .so)
  for net in default stub; do
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tree_wave -s 300 -c 200 --csv --log-file gpurun_out/q_${n}_${net}.csv python bench.py --net $net --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
    python - <<PY
import csv, statistics
rows=[r for r in csv.reader(open("gpurun_out/q_${n}_${net}.csv")) if len(r)>5 and r[-1].replace('.','').isdigit()]
v=[float(r[-1])/1e3 for r in rows]
print("$n $net ncu us: median %.2f mean %.2f p90 %.2f" % (statistics.median(v), sum(v)/len(v), sorted(v)[int(0.9*len(v))]))
PY
    python bench.py --net $net --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python tools/fmt_bench.py
  done
done
cp /tmp/orig.so $P/libifx_b200.so
