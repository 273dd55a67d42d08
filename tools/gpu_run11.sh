#!/bin/bash
# deliverables pass: the driver's own commands + multi-GPU arm + reference arm
mkdir -p gpurun_out
rm -f gpurun_out/run11_summary.txt
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest -m gpu exit $?" | tee -a gpurun_out/run11_summary.txt; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?" | tee -a gpurun_out/run11_summary.txt; tail -1 gpurun_out/smoke.log
nproc; lscpu | grep -E "Model name|^CPU\(s\)" 
timeout 1200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.log 2>&1; echo "== bench reference exit $?" | tee -a gpurun_out/run11_summary.txt; tail -1 gpurun_out/bench_reference.log | cut -c1-900
timeout 1500 python bench.py > gpurun_out/bench_default.log 2>&1; echo "== bench default exit $?" | tee -a gpurun_out/run11_summary.txt; tail -1 gpurun_out/bench_default.log | cut -c1-2500
