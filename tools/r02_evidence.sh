#!/bin/bash
# round-2 evidence run (one B200): bench lines (default / no-CFG / 1024x768 / reference arm), step timeline, per-op profiles, ncu --set full of
# every kernel (reduced on the box to a raw CSV + summary: the .ncu-rep itself is over gpurun's 64 MiB return limit), ncu launch list of the bench
# command.  Everything lands in gpurun_out/ (copied to profiles/ by hand).  The -m gpu suite is run separately (tools/r02_check.sh).
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/r02_bench_n1.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --guidance 1.0 --no-cpu-baseline > gpurun_out/r02_bench_nocfg.log 2>&1; echo "== bench nocfg exit $?"; tail -n 1 gpurun_out/r02_bench_nocfg.log | cut -c1-200
timeout 900 python bench.py --steps 3 --warmup 3 --height 1024 --width 768 --batch 4 --no-cpu-baseline > gpurun_out/r02_bench_1024.log 2>&1; echo "== bench 1024 exit $?"; tail -n 1 gpurun_out/r02_bench_1024.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.log 2>&1; echo "== reference arm exit $?"; tail -n 1 gpurun_out/r02_bench_reference_arm.log | cut -c1-300
timeout 600 python tools/timeline.py > gpurun_out/r02_timeline.txt 2>&1; echo "== timeline exit $?"; sed -n 3,4p gpurun_out/r02_timeline.txt
timeout 600 python tools/profile_ops.py 8 512 384 r02_op_profile.txt > /dev/null 2>&1; echo "== op profile exit $?"
timeout 600 python tools/profile_ops.py 4 1024 768 r02_op_profile_1024.txt > /dev/null 2>&1; echo "== op profile 1024 exit $?"
./tools/ubench/softmax_loop > gpurun_out/r02_softmax_loop.txt 2>&1; cat gpurun_out/r02_softmax_loop.txt
timeout 1200 ncu --set full --clock-control none --profile-from-start off -o /tmp/r02_kernels python tools/ncu_kernels.py > gpurun_out/r02_ncu_kernels.log 2>&1; echo "== ncu kernels exit $?"; tail -2 gpurun_out/r02_ncu_kernels.log
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv 2>/dev/null; ls -la /tmp/r02_kernels.ncu-rep gpurun_out/r02_kernels_raw.csv
python tools/ncu_summarize.py gpurun_out/r02_kernels_raw.csv "ncu --set full --clock-control none --profile-from-start off python tools/ncu_kernels.py" --traffic-json gpurun_out/r02_ncu_traffic.json "[8, 512, 384, true]" > gpurun_out/r02_ncu_kernels.txt; head -50 gpurun_out/r02_ncu_kernels.txt | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/r02_launches.csv python bench.py --steps 1 --warmup 1 --ddim-steps 4 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1; echo "== ncu list exit $?"
python tools/summarize_launches.py /tmp/r02_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 1 --warmup 1 --ddim-steps 4 --no-cpu-baseline" > gpurun_out/r02_launches_summary.txt
du -sh gpurun_out
