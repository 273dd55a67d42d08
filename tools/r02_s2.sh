#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_d512.py > gpurun_out/r02_diag_d512.txt 2>&1; echo "== diag exit $?"; cat gpurun_out/r02_diag_d512.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider --timeout 300 -k "up2x" 2>&1 | tail -5
for fold in 0 1; do for up in 1 0; do
  LADI_LN_FOLD=$fold LADI_UP2X=$up timeout 600 python tools/timeline.py > gpurun_out/r02_timeline_fold${fold}_up${up}.txt 2>&1; echo "== timeline fold=$fold up=$up exit $?"; sed -n 3,4p gpurun_out/r02_timeline_fold${fold}_up${up}.txt
done; done
