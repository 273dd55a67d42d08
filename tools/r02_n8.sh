#!/bin/bash
# BASELINE configs[3] and [4] on the 8 GPUs of one box (gpurun --gpus 8): 1024x768 batch 4/GPU (CFG 7.5 = CLI default, and guidance 1.0), and
# 512x384, 100 DDIM steps, CFG 7.5.  One torchrun per config; rank 0 prints the JSON line.
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02_bench_n8_$name.log 2>&1
  echo "== $name exit $?"; grep '^{"metric"' gpurun_out/r02_bench_n8_$name.log | tail -1 > gpurun_out/r02_bench_n8_$name.json; cut -c1-420 gpurun_out/r02_bench_n8_$name.json
}
run c4_1024x768_b4_cfg --height 1024 --width 768 --batch 4
run c5_n100_cfg --ddim-steps 100
run c4_1024x768_b4_nocfg --height 1024 --width 768 --batch 4 --guidance 1.0
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv | head -9
