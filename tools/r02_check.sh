#!/bin/bash
# round-2 GPU check: the new kernels' unit tests first (fail fast), then the whole -m gpu suite, smoke, a short bench and the step timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider --timeout 300 -k "d512 or up2x or ln_fold or rowstat or ddim or binarise" > gpurun_out/r02_newops.log 2>&1; echo "== new ops exit $?"; tail -15 gpurun_out/r02_newops.log
timeout 1700 python -m pytest tests/ -q -m gpu -p no:cacheprovider --timeout 900 -s > gpurun_out/r02_pytest_gpu.log 2>&1; echo "== pytest -m gpu exit $?"; grep -E "passed|failed|FAILED|Error|rel-L2|mean\|" gpurun_out/r02_pytest_gpu.log | tail -40
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_quick.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/r02_bench_quick.log | cut -c1-1500
timeout 600 python tools/timeline.py > gpurun_out/r02_timeline.txt 2>&1; echo "== timeline exit $?"; sed -n 3,24p gpurun_out/r02_timeline.txt
