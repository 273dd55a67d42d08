#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/models.log 2>&1
echo "== models exit $?" | tee gpurun_out/run3_summary.txt
grep -E "rel-L2|mean\|engine|passed|failed|FAILED" gpurun_out/models.log
timeout 600 python tools/profile_ops.py 8 > gpurun_out/profile_ops.log 2>&1; echo "== profile exit $?" | tee -a gpurun_out/run3_summary.txt
head -c 5000 gpurun_out/op_profile.txt
# ncu launch list of the bench command (short: 1 step, 1 warm-up) -- shares, not absolutes
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --ddim-steps 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "== ncu list exit $?" | tee -a gpurun_out/run3_summary.txt
