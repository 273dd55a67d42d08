#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "== pytest -m gpu exit $?"; grep -E "passed|failed|FAILED" gpurun_out/r02_pytest_gpu_final.log | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python tools/xattn_bench.py > gpurun_out/r02_xattn_bench.jsonl 2>&1
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1_final.log 2>&1; echo "== bench exit $?"; grep '^{"metric"' gpurun_out/r02_bench_n1_final.log | tail -1 > gpurun_out/r02_bench_n1_final.json; cut -c1-330 gpurun_out/r02_bench_n1_final.json
timeout 600 python tools/timeline.py > gpurun_out/r02_timeline_final.txt 2>&1; echo "== timeline exit $?"; sed -n 3,22p gpurun_out/r02_timeline_final.txt
./tools/ubench/softmax_loop > gpurun_out/r02_softmax_loop_v2.txt 2>&1; cat gpurun_out/r02_softmax_loop_v2.txt | cut -c1-200
