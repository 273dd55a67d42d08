"""Cross-attention (77 text tokens) timing: one-tile-per-CTA kernel (variant 1) vs the persistent short-K/V kernel (variant 8), UNet batch 16, the three
UNet levels + mid block.  Median of 20 after 5 warm-ups, L2 flushed.  One JSON line per shape."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladi_vton_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
for label, B, N, H in (("level 0 (64x48)", 16, 3072, 5), ("level 1 (32x24)", 16, 768, 10), ("level 2 (16x12)", 16, 192, 20), ("mid (8x6)", 16, 48, 20)):
    C = H * 64
    q = torch.randn((B, N, C), device=dev).bfloat16()
    kv = torch.randn((B, 77, 2 * C), device=dev).bfloat16()
    line = dict(shape=label, batch=B, nq=N, heads=H, nkv=77)
    outs = {}
    for rep in range(2):
        for var in (1, 8):
            run = lambda: ops.attention(q, kv[..., :C], kv[..., C:], H, 0.125, variant=var)
            for _ in range(5):
                outs[var] = run()
            ts = []
            for _ in range(20):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            line[f"v{var}_us_r{rep}"] = round(ts[len(ts) // 2] * 1e3, 1)
    line["max_abs_diff_v8_vs_v1"] = float((outs[8].float() - outs[1].float()).abs().max())
    print(json.dumps(line), flush=True)
