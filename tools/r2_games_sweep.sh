#!/bin/bash
# GPU box: the tree kernel's roofline fraction against the number of games per launch (stub evaluator: tree + rules only).
# One warp per game and all 4 096 resident at once means a launch lasts as long as its deepest descent; with more games than
# resident warps, CTAs that finish early make room and the depth tail is amortised.
mkdir -p gpurun_out
for g in 4096 8192 16384 32768; do
  timeout 300 python bench.py --net stub --games $g --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines \
      > gpurun_out/sweep_stub_$g.json 2> gpurun_out/sweep_stub_$g.err
  echo "games $g rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep_stub_$g.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("games", $g, "value %.3e" % d["value"], "frac %.3f" % r["frac"], "avg_launch_us %.2f" % r["avg_launch_us"],
          "frac_dram", r.get("frac_dram"), "share", "%.3f" % r["share_of_step"])
except Exception as e:
    print("parse failed", e)
PY
done
for g in 8192 16384; do
  timeout 600 python bench.py --games $g --steps 6 --warmup 3 --no-cpu-baseline --no-extra-rooflines \
      > gpurun_out/sweep_default_$g.json 2> gpurun_out/sweep_default_$g.err
  echo "default games $g rc=$?"; tail -c 1500 gpurun_out/sweep_default_$g.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value %.3e e2e %.3e frac %.3f us %.2f ms/step %.1f' % (d['value'], d['e2e']['value'], r['frac'], r['avg_launch_us'], d['ms_per_step']))"
done
