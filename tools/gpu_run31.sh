#!/bin/bash
# full GPU suite, step timeline, bench (default), after the attention issuer split + CTA-pair convs
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu_r31.log 2>&1; echo "== pytest -m gpu exit $?"; tail -3 gpurun_out/pytest_gpu_r31.log
timeout 600 python tools/timeline.py > gpurun_out/timeline_r31.txt 2>&1; echo "== timeline exit $?"; sed -n 3,22p gpurun_out/timeline_r31.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r31.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/bench_r31.log | cut -c1-300
