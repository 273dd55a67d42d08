#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -k "adapter" --timeout 300 -p no:cacheprovider -s 2>&1 | tail -4
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --guidance 1.0 > gpurun_out/bench_nocfg.log 2>&1; echo "== nocfg exit $?"; tail -n 1 gpurun_out/bench_nocfg.log | cut -c1-330
timeout 1200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --height 1024 --width 768 --batch 4 --guidance 1.0 > gpurun_out/bench_1024.log 2>&1; echo "== 1024 exit $?"; tail -n 2 gpurun_out/bench_1024.log | cut -c1-330
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference3.log 2>&1; tail -1 gpurun_out/bench_reference3.log | grep -o '"cpu_baseline.*' | cut -c1-400
