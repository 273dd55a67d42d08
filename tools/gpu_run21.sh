#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "pipeline" 2>&1 | tail -4
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r21.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/bench_r21.log | cut -c1-250
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --height 1024 --width 768 --batch 4 > gpurun_out/bench_1024_cfg.log 2>&1; echo "== 1024cfg exit $?"; tail -n 1 gpurun_out/bench_1024_cfg.log | cut -c1-250
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ddim-steps 100 > gpurun_out/bench_n100.log 2>&1; echo "== n100 exit $?"; tail -n 1 gpurun_out/bench_n100.log | cut -c1-250
