#!/bin/bash
mkdir -p gpurun_out
for fam in gemm conv "groupnorm or layernorm"; do
  name=$(echo "$fam" | cut -d' ' -f1)
  timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "$fam" --timeout 120 -p no:cacheprovider > gpurun_out/ops_$name.log 2>&1
  echo "== ops $name exit $?" | tee -a gpurun_out/run6_summary.txt
  grep -E "passed|failed" gpurun_out/ops_$name.log; grep -E "^FAILED|watchdog|Error" gpurun_out/ops_$name.log | head -20
done
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/models.log 2>&1
echo "== models exit $?" | tee -a gpurun_out/run6_summary.txt
grep -E "rel-L2|mean\|engine|passed|failed|FAILED" gpurun_out/models.log
timeout 600 python tools/microbench.py > gpurun_out/microbench.txt 2>&1; head -30 gpurun_out/microbench.txt
timeout 600 python tools/profile_ops.py 8 > gpurun_out/profile_ops.log 2>&1; echo "== profile exit $?" | tee -a gpurun_out/run6_summary.txt
head -c 2500 gpurun_out/op_profile.txt; grep family gpurun_out/op_profile.txt | head -8
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench3.log 2>&1; echo "== bench exit $?" | tee -a gpurun_out/run6_summary.txt; tail -n 1 gpurun_out/bench3.log | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"convgemm|attention" -c 8 -o gpurun_out/r01_top_kernels python tools/ncu_conv.py > gpurun_out/ncu_full.log 2>&1; echo "== ncu full exit $?" | tee -a gpurun_out/run6_summary.txt; tail -3 gpurun_out/ncu_full.log
