#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_warp.py -m gpu -q --timeout 900 -p no:cacheprovider -s 2>&1 | tail -40
