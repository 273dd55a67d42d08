#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_warp.py -m gpu -q --timeout 900 -p no:cacheprovider -s > gpurun_out/warp_tests.log 2>&1; tail -5 gpurun_out/warp_tests.log; grep -E "err|rel-L2|FAILED|Error" gpurun_out/warp_tests.log | head -20
timeout 600 python tools/warp_bench.py > gpurun_out/warp_bench.log 2>&1; echo "== warp bench exit $?"; tail -n 3 gpurun_out/warp_bench.log
