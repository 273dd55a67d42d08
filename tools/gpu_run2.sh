#!/bin/bash
# model-level parity + smoke + first bench + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/models.log 2>&1
echo "== models exit $?" | tee gpurun_out/run2_summary.txt
tail -n 30 gpurun_out/models.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?" | tee -a gpurun_out/run2_summary.txt; tail -n 5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench1.log 2>&1; echo "== bench exit $?" | tee -a gpurun_out/run2_summary.txt; tail -n 12 gpurun_out/bench1.log
