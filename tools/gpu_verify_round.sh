#!/bin/bash
# round-1 final verification: full GPU suite, smoke, step timeline, bench (with CPU baseline), ncu launch list of the bench command
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu_verify.log 2>&1; echo "== pytest -m gpu exit $?"; tail -3 gpurun_out/pytest_gpu_verify.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python tools/timeline.py > gpurun_out/timeline_verify.txt 2>&1; echo "== timeline exit $?"; sed -n 3,22p gpurun_out/timeline_verify.txt
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_verify.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/bench_verify.log | cut -c1-250
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_verify.csv python bench.py --steps 1 --warmup 1 --ddim-steps 4 --no-cpu-baseline > gpurun_out/ncu_bench_verify.log 2>&1; echo "== ncu list exit $?"
python tools/summarize_launches.py gpurun_out/launches_verify.csv "ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 1 --warmup 1 --ddim-steps 4 --no-cpu-baseline" > gpurun_out/launches_summary_verify.txt; head -14 gpurun_out/launches_summary_verify.txt; rm -f gpurun_out/launches_verify.csv
