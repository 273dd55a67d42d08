#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dataprep.py -x -q -k "attention or pose or image_out or pil" -p no:cacheprovider --timeout 300 > gpurun_out/pytest_attn_r32.log 2>&1; echo "== attention+dataprep tests exit $?"; tail -4 gpurun_out/pytest_attn_r32.log
rm -f gpurun_out/attn_bench_r32.jsonl
timeout 120 python tools/attn_bench.py >> gpurun_out/attn_bench_r32.jsonl 2>gpurun_out/attn_bench_r32.err; echo "== bench exit $?"; cat gpurun_out/attn_bench_r32.jsonl
