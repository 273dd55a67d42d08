#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/attn_bench_r33.jsonl
ATTN_VARIANTS=5,6 timeout 200 python tools/attn_bench.py >> gpurun_out/attn_bench_r33.jsonl 2>gpurun_out/attn_bench_r33.err; echo "== attn A/B exit $?"; cat gpurun_out/attn_bench_r33.jsonl
timeout 300 python tools/bn_sweep.py > gpurun_out/bn_sweep_r33.jsonl 2>gpurun_out/bn_sweep_r33.err; echo "== bn sweep exit $?"; cat gpurun_out/bn_sweep_r33.jsonl; tail -3 gpurun_out/bn_sweep_r33.err
