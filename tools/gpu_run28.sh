#!/bin/bash
mkdir -p gpurun_out
LADI_B200_LIB=ladi_vton_b200/libladi_b200_trace.so timeout 300 python tools/attn_trace.py > gpurun_out/attn_trace_r28.txt 2>&1; echo "== trace exit $?"; head -70 gpurun_out/attn_trace_r28.txt; tail -28 gpurun_out/attn_trace_r28.txt
timeout 120 tools/ubench/mufu > gpurun_out/mufu_r28.txt 2>&1; grep "1024" gpurun_out/mufu_r28.txt
