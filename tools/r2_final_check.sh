#!/bin/bash
# GPU box: what the driver runs at round end -- the GPU suite, smoke(), both bench arms.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
S=$(date +%s); timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference rc=$? $(( $(date +%s) - S )) s"
S=$(date +%s); timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "default rc=$? $(( $(date +%s) - S )) s"
python tools/fmt_bench.py < gpurun_out/final_bench_default.json | head -5
cut -c1-400 gpurun_out/final_bench_reference.json
