#!/bin/bash
# Runs the GPU op tests one kernel family per process (a trapped kernel poisons only its own process); logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/probe_gpu.txt 2>&1
for fam in gemm conv attention "groupnorm or layernorm or softmax or pointwise or ddim"; do
  name=$(echo "$fam" | cut -d' ' -f1)
  timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "$fam" --timeout 120 -p no:cacheprovider > gpurun_out/probe_$name.log 2>&1
  echo "== $name exit $?" | tee -a gpurun_out/probe_summary.txt
  tail -n 25 gpurun_out/probe_$name.log
done
