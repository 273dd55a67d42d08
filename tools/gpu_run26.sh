#!/bin/bash
# CLI mirror tests, then the CTA-pair (cta_group::2) conv kernel: parity tests and A/B timing against the single-CTA kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -s -p no:cacheprovider --timeout 600 > gpurun_out/pytest_cli_r26.log 2>&1; echo "== cli tests exit $?"; grep -E "rel-L2|mean\||passed|failed|Error|error" gpurun_out/pytest_cli_r26.log | tail -12
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "pair" -p no:cacheprovider --timeout 300 > gpurun_out/pytest_pair_r26.log 2>&1; echo "== pair tests exit $?"; tail -15 gpurun_out/pytest_pair_r26.log
timeout 300 python tools/pair_bench.py > gpurun_out/pair_bench_r26.jsonl 2> gpurun_out/pair_bench_r26.err; echo "== pair bench exit $?"; cat gpurun_out/pair_bench_r26.jsonl; tail -5 gpurun_out/pair_bench_r26.err
