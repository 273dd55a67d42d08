#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ncu_gn.py 2>&1 | tail -8
NCU_GN_TIMING=0 timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"gn_|layernorm" -c 6 -o gpurun_out/r01_gn_kernels python tools/ncu_gn.py > gpurun_out/ncu_gn.log 2>&1; echo "== ncu exit $?"; tail -2 gpurun_out/ncu_gn.log
ncu -i gpurun_out/r01_gn_kernels.ncu-rep --page raw --csv > gpurun_out/r01_gn_kernels_raw.csv 2>/dev/null; ls -la gpurun_out/r01_gn_kernels*
