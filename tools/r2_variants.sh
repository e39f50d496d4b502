#!/bin/bash
# dev tool (GPU box): for each library variant under variants/*.so -- ncu duration of 8 waves + the event-timed bench line
P=alpha-zero-general-inflexion_b200
cp $P/libifx_b200.so /tmp/orig.so
for so in "$@"; do
  cp $so $P/libifx_b200.so
  n=$(basename $so .so)
  for net in default stub; do
    ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:tree_wave -s 352 -c 8 --csv --log-file gpurun_out/v_${n}_${net}.csv python bench.py --net $net --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines > /dev/null 2>&1
    python - <<PY
import csv, statistics
rows=[r for r in csv.reader(l for l in open("gpurun_out/v_${n}_${net}.csv") if l.startswith('"'))]
h=rows[0]; im=h.index("Metric Name"); iv=h.index("Metric Value")
d=[float(r[iv].replace(",",""))/1e3 for r in rows[1:] if r[im]=="gpu__time_duration.sum"]
i=[float(r[iv].replace(",","")) for r in rows[1:] if r[im]=="smsp__inst_executed.sum"]
print("$n $net ncu us: median %.2f min %.2f; instr/launch %.0f" % (statistics.median(d), min(d), statistics.median(i)))
PY
    python bench.py --net $net --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines 2>/dev/null | python tools/fmt_bench.py
  done
done
cp /tmp/orig.so $P/libifx_b200.so
