#!/bin/bash
# dev tool (GPU box): per-launch gpu__time_duration of the wave kernel under ncu + per-warp trace, for library variants
P=alpha-zero-general-inflexion_b200
cp $P/libifx_b200.so /tmp/orig.so
for so in "$@"; do
  cp $so $P/libifx_b200.so
  n=$(basename $so .so)
  echo "== $so"
  for net in default stub; do
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tree_wave -s 300 -c 300 --csv --log-file gpurun_out/wave_${n}_${net}.csv python bench.py --net $net --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
    python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/wave_${n}_${net}.csv")) if len(r)>5 and r[-1].replace('.','').isdigit()]
v=[float(r[-1]) for r in rows]
u=rows[0][-2] if rows else ''
import statistics
print("$net ncu per-launch duration: n=%d mean %.2f median %.2f min %.2f max %.2f %s" % (len(v), sum(v)/max(len(v),1), statistics.median(v) if v else 0, min(v) if v else 0, max(v) if v else 0, u))
PY
    python tools/trace_wave.py 4096 $net 2>&1 | tail -12
  done
done
cp /tmp/orig.so $P/libifx_b200.so
