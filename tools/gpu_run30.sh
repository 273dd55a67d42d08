#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "attention" -p no:cacheprovider --timeout 300 > gpurun_out/pytest_attn_r30.log 2>&1; echo "== attention tests exit $?"; tail -4 gpurun_out/pytest_attn_r30.log
LADI_B200_LIB=ladi_vton_b200/libladi_b200_trace.so timeout 300 python tools/attn_trace.py > gpurun_out/attn_trace_r30.txt 2>&1; echo "== trace exit $?"; grep "mean over" gpurun_out/attn_trace_r30.txt; grep -A14 "softmax A.half0" gpurun_out/attn_trace_r30.txt | tail -6; grep -A14 "softmax B.half0" gpurun_out/attn_trace_r30.txt | tail -6;  grep -A14 "MMA issuer of tile A" gpurun_out/attn_trace_r30.txt | tail -6
rm -f gpurun_out/attn_bench_r30.jsonl
for d in 0 700 1400; do LADI_ATTN_B_DELAY=$d timeout 120 python tools/attn_bench.py >> gpurun_out/attn_bench_r30.jsonl 2>gpurun_out/attn_bench_r30.err; done; echo "== bench exit $?"; cat gpurun_out/attn_bench_r30.jsonl
