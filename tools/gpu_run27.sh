#!/bin/bash
# full GPU suite with the CTA-pair conv kernels on by default, step timeline, bench, ncu --set full of the conv/attention kernels
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu_r27.log 2>&1; echo "== pytest -m gpu exit $?"; tail -3 gpurun_out/pytest_gpu_r27.log
timeout 600 python tools/timeline.py > gpurun_out/timeline_r27.txt 2>&1; echo "== timeline exit $?"; sed -n 3,22p gpurun_out/timeline_r27.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r27.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/bench_r27.log | cut -c1-300
LADI_CONV_2CTA=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r27_single.log 2>&1; echo "== bench (single-CTA) exit $?"; tail -n 1 gpurun_out/bench_r27_single.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"convgemm|attention" -c 8 -o gpurun_out/r01_pair_kernels python tools/ncu_conv.py > gpurun_out/ncu_pair.log 2>&1; echo "== ncu full exit $?"; tail -2 gpurun_out/ncu_pair.log
ncu -i gpurun_out/r01_pair_kernels.ncu-rep --page raw --csv > gpurun_out/r01_pair_kernels_raw.csv 2>/dev/null; ls -la gpurun_out/r01_pair_kernels*
