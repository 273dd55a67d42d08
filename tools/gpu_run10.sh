#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/run10_summary.txt
for fam in gemm conv attention "groupnorm or layernorm or softmax or pointwise or ddim"; do
  name=$(echo "$fam" | cut -d' ' -f1)
  timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "$fam" --timeout 120 -p no:cacheprovider > gpurun_out/ops_$name.log 2>&1
  echo "== ops $name exit $?" | tee -a gpurun_out/run10_summary.txt
  grep -E "passed|failed" gpurun_out/ops_$name.log; grep -E "^FAILED|watchdog|Error" gpurun_out/ops_$name.log | head -20
done
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/models.log 2>&1
echo "== models exit $?" | tee -a gpurun_out/run10_summary.txt
grep -E "rel-L2|mean\|engine|passed|failed|FAILED" gpurun_out/models.log
timeout 600 python tools/timeline.py > gpurun_out/timeline.log 2>&1; head -18 gpurun_out/timeline.txt
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench4.log 2>&1; echo "== bench exit $?" | tee -a gpurun_out/run10_summary.txt; tail -n 1 gpurun_out/bench4.log | cut -c1-200
