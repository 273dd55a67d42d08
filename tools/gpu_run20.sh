#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_frontend.py -m gpu -q --timeout 600 -p no:cacheprovider -x -s 2>&1 | tail -25
timeout 600 python tools/frontend_bench.py > gpurun_out/frontend_bench.log 2>&1; echo "== frontend exit $?"; tail -n 4 gpurun_out/frontend_bench.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r20.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/bench_r20.log | cut -c1-400
