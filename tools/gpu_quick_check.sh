#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "gemm or conv" -p no:cacheprovider --timeout 300 > gpurun_out/pytest_conv_quick.log 2>&1; echo "== conv/gemm tests exit $?"; tail -2 gpurun_out/pytest_conv_quick.log
timeout 600 python tools/timeline.py > gpurun_out/timeline_quick.txt 2>&1; echo "== timeline exit $?"; sed -n 4,14p gpurun_out/timeline_quick.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/bench_quick.log | cut -c1-200
