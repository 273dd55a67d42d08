#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/diag_d512.py > gpurun_out/r02_diag_d512.txt 2>&1; echo "== diag exit $?"; cat gpurun_out/r02_diag_d512.txt | cut -c1-200
timeout 1700 python -m pytest tests/ -q -m gpu -p no:cacheprovider --timeout 900 -s > gpurun_out/r02_pytest_gpu.log 2>&1; echo "== pytest -m gpu exit $?"; grep -E "passed|failed|FAILED|Error|rel-L2|mean\|" gpurun_out/r02_pytest_gpu.log | tail -40
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_quick.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/r02_bench_quick.log | cut -c1-3000
timeout 600 python tools/profile_ops.py 8 512 384 r02_op_profile.txt > /dev/null 2>&1; echo "== profile exit $?"; head -75 gpurun_out/r02_op_profile.txt
