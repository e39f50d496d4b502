#!/bin/bash
# GPU box: the captured (net, wave) step -- its tests, then tools/bench_leaf_parallel.py with the eager and captured columns.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_late_additions.py tests/test_virtual_loss.py tests/test_coach_arena.py tests/test_dropin_api.py \
    tests/test_leaf_gather.py tests/test_selfplay.py -x -q -m gpu > gpurun_out/captured_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/captured_tests.log
tail -5 gpurun_out/captured_tests.log
timeout 600 python tools/bench_leaf_parallel.py 200 > gpurun_out/leaf_parallel_captured.txt 2>&1
echo "bench rc=$?"
cat gpurun_out/leaf_parallel_captured.txt
