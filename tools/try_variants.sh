#!/bin/bash
# usage (on the GPU box): tools/try_variants.sh <lib1.so> <lib2.so> ...   -- dev tool: quick bench of library variants
P=alpha-zero-general-inflexion_b200
cp $P/libifx_b200.so /tmp/orig.so
fmt='
import sys,json
for line in sys.stdin:
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print("sims/s %.3e e2e %s ms/step %.2f tree_us %.1f GB/s %.0f frac %.3f share %.3f V %.2f" % (d["value"], d["e2e"] and "%.3e" % d["e2e"]["value"], d["ms_per_step"], r["avg_launch_us"], r["achieved"], r["frac"], r["share_of_step"], r["visits_per_sim"]))
    else: print(line.rstrip()[-300:])
'
for so in "$@"; do
  cp $so $P/libifx_b200.so
  echo "== $so"
  python bench.py --net stub --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "$fmt"
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "$fmt"
done
cp /tmp/orig.so $P/libifx_b200.so
