"""Times the fused attention kernel (C ABI, CUDA events, median of 20 after 5 warm-ups, L2 flushed) on the UNet's self-attention
shapes.  One JSON line per shape.  LADI_ATTN_B_DELAY=<cycles> tunes the phase offset of the pair kernel's two query tiles."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladi_vton_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
for label, B, N, H in (("512x384 level 0 (64x48)", 16, 3072, 5), ("512x384 level 1 (32x24)", 16, 768, 10), ("512x384 level 2 (16x12)", 16, 192, 20),
                       ("1024x768 level 0 (128x96)", 8, 12288, 5)):
    q = torch.randn((B, N, 3 * H * 64), device=dev).bfloat16()
    C = H * 64
    variants = [int(v) for v in os.environ.get("ATTN_VARIANTS", "0").split(",")]
    fl = 4.0 * B * H * N * N * 64
    line = dict(shape=label, batch=B, tokens=N, heads=H, b_delay=os.environ.get("LADI_ATTN_B_DELAY", "default"))
    for rep in range(2):  # variants interleaved, two rounds: same process, same clocks
        for var in variants:
            if var in (2, 4, 5, 6) and (N <= 128 or N < 512):
                continue
            run = lambda: ops.attention(q[..., :C], q[..., C:2 * C], q[..., 2 * C:], H, 0.125, variant=var)
            for _ in range(5):
                run()
            ts = []
            for _ in range(20):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            us = ts[len(ts) // 2] * 1e3
            line[f"v{var}_us_r{rep}"] = round(us, 1)
            line[f"v{var}_tflops_r{rep}"] = round(fl / us / 1e6, 1)
    print(json.dumps(line), flush=True)
