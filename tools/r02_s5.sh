#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -q -m gpu -p no:cacheprovider --timeout 600 -k "attention or engine_abi" 2>&1 | tail -6
ATTN_VARIANTS=5,6 timeout 600 python tools/attn_bench.py > gpurun_out/r02_attn_bench.jsonl 2> gpurun_out/r02_attn_bench.err; echo "== attn bench exit $?"; cat gpurun_out/r02_attn_bench.jsonl
