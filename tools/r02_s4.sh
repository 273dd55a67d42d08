#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "== pytest -m gpu exit $?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r02_pytest_gpu.log | tail -20
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_engine.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/r02_bench_engine.log | cut -c1-400
LADI_ENGINE=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_pyseq.log 2>&1; echo "== bench (python sequencing) exit $?"; tail -n 1 gpurun_out/r02_bench_pyseq.log | cut -c1-400
