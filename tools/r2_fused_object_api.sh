#!/bin/bash
# GPU box: the object API on the fused head -- its tests, configs[0] and the single-game lines of bench_leaf_parallel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_late_additions.py tests/test_dropin_api.py tests/test_coach_arena.py tests/test_real_net_parity.py -x -q -m gpu > gpurun_out/fused_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/fused_tests.log
timeout 300 python tools/bench_config0.py ours > gpurun_out/r2_config0_fused.txt 2>&1; cat gpurun_out/r2_config0_fused.txt | tail -3
timeout 300 python tools/bench_leaf_parallel.py 200 2>&1 | grep "MCTS.getActionProb" > gpurun_out/object_api_fused.txt; cat gpurun_out/object_api_fused.txt
