#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "attention" -p no:cacheprovider --timeout 300 > gpurun_out/pytest_attn_r29.log 2>&1; echo "== attention tests exit $?"; tail -4 gpurun_out/pytest_attn_r29.log
LADI_B200_LIB=ladi_vton_b200/libladi_b200_trace.so timeout 300 python tools/attn_trace.py > gpurun_out/attn_trace_r29.txt 2>&1; echo "== trace exit $?"; grep -A30 "softmax A.half0" gpurun_out/attn_trace_r29.txt | head -32; grep -A27 "softmax B.half0" gpurun_out/attn_trace_r29.txt | tail -3;  grep -A26 "MMA issuer of tile A" gpurun_out/attn_trace_r29.txt | head -16
for d in 0 350 700 1050 1400; do LADI_ATTN_B_DELAY=$d timeout 120 python tools/attn_bench.py >> gpurun_out/attn_bench_r29.jsonl 2>gpurun_out/attn_bench_r29.err; done; echo "== bench exit $?"; cat gpurun_out/attn_bench_r29.jsonl
