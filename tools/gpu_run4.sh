#!/bin/bash
mkdir -p gpurun_out
for fam in gemm conv attention; do
  timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "$fam" --timeout 120 -p no:cacheprovider > gpurun_out/ops_$fam.log 2>&1
  echo "== ops $fam exit $?" | tee -a gpurun_out/run4_summary.txt
  grep -E "passed|failed" gpurun_out/ops_$fam.log; grep -E "^FAILED|watchdog|Error" gpurun_out/ops_$fam.log | head -20
done
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/models.log 2>&1
echo "== models exit $?" | tee -a gpurun_out/run4_summary.txt
grep -E "rel-L2|mean\|engine|passed|failed|FAILED" gpurun_out/models.log
timeout 600 python tools/profile_ops.py 8 > gpurun_out/profile_ops.log 2>&1; echo "== profile exit $?" | tee -a gpurun_out/run4_summary.txt
head -c 4500 gpurun_out/op_profile.txt
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench2.log 2>&1; echo "== bench exit $?" | tee -a gpurun_out/run4_summary.txt; tail -n 3 gpurun_out/bench2.log | cut -c1-1800
