#!/bin/bash
# Quick A/B of the tree kernel: parity tests, the two bench modes, and ncu duration / instruction counts of two waves each.
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra-rooflines > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}_default.err
python bench.py --net stub --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_stub.json 2> gpurun_out/${TAG}_stub.err
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_per_inst_issued.ratio,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -k regex:tree_wave -s 350 -c 4 --csv --log-file gpurun_out/${TAG}_ncu_default.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:tree_wave -s 350 -c 4 --csv --log-file gpurun_out/${TAG}_ncu_stub.csv python bench.py --net stub --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python - <<PY
import json,csv
for f in ("default","stub"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json"%f)); r=d["roofline"]
        print(f, "value %.4g ms/step %.2f frac %.3f launch_us %.2f V/sim %.2f e2e %.4g" % (d["value"], d["ms_per_step"], r["frac"], r["avg_launch_us"], r["visits_per_sim"], (d.get("e2e") or {}).get("value",0)))
    except Exception as e:
        print(f, "FAILED", e, open("gpurun_out/${TAG}_%s.err"%f).read()[-800:])
    try:
        rows=[r for r in csv.reader(l for l in open("gpurun_out/${TAG}_ncu_%s.csv"%f) if l.startswith('"'))]
        h=rows[0]; im=h.index("Metric Name"); iv=h.index("Metric Value"); ii=h.index("ID")
        agg={}
        for r in rows[1:]:
            agg.setdefault(r[im],[]).append(r[iv])
        for k,v in agg.items(): print("   ",k,v)
    except Exception as e:
        print("ncu parse failed",e)
PY
