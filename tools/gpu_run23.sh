#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "groupnorm or layernorm" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "small" 2>&1 | tail -3
timeout 600 python tools/timeline.py > gpurun_out/timeline_r23.txt 2>&1; echo "== timeline exit $?"; head -22 gpurun_out/timeline_r23.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r23.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/bench_r23.log | cut -c1-200
