#!/bin/bash
# round-2 evidence run (one B200): full -m gpu suite, smoke, bench lines (default / no-CFG / 1024x768 / reference arm), step timeline, per-op
# profiles, ncu --set full of every kernel, ncu launch list of the bench command.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "== pytest -m gpu exit $?"; grep -E "passed|failed|FAILED" gpurun_out/r02_pytest_gpu.log | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/r02_bench_n1.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --guidance 1.0 --no-cpu-baseline > gpurun_out/r02_bench_nocfg.log 2>&1; echo "== bench nocfg exit $?"; tail -n 1 gpurun_out/r02_bench_nocfg.log | cut -c1-200
timeout 900 python bench.py --steps 3 --warmup 3 --height 1024 --width 768 --batch 4 --no-cpu-baseline > gpurun_out/r02_bench_1024.log 2>&1; echo "== bench 1024 exit $?"; tail -n 1 gpurun_out/r02_bench_1024.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.log 2>&1; echo "== reference arm exit $?"; tail -n 1 gpurun_out/r02_bench_reference_arm.log | cut -c1-300
timeout 600 python tools/timeline.py > gpurun_out/r02_timeline.txt 2>&1; echo "== timeline exit $?"; sed -n 3,4p gpurun_out/r02_timeline.txt
timeout 600 python tools/profile_ops.py 8 512 384 r02_op_profile.txt > /dev/null 2>&1; echo "== op profile exit $?"
timeout 600 python tools/profile_ops.py 4 1024 768 r02_op_profile_1024.txt > /dev/null 2>&1; echo "== op profile 1024 exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_kernels python tools/ncu_kernels.py > gpurun_out/r02_ncu_kernels.log 2>&1; echo "== ncu kernels exit $?"; tail -2 gpurun_out/r02_ncu_kernels.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --ddim-steps 4 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1; echo "== ncu list exit $?"
python tools/summarize_launches.py gpurun_out/r02_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 1 --warmup 1 --ddim-steps 4 --no-cpu-baseline" > gpurun_out/r02_launches_summary.txt; head -20 gpurun_out/r02_launches_summary.txt; rm -f gpurun_out/r02_launches.csv
ls -la gpurun_out/r02_kernels.ncu-rep
