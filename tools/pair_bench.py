"""A/B timing of the CTA-pair (tcgen05 cta_group::2) conv/GEMM kernel against the single-CTA kernel on the UNet's dominant shapes
(UNet batch 16 = 8 images with CFG, 512x384).  CUDA events, 20 iterations after 5 warm-ups, L2 flushed between iterations by a
256 MB write.  Prints one JSON line per shape.

    python tools/pair_bench.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [  # (label, n, h, w, cin, cout, ksize, residual)
    ("conv3x3 320->320 @64x48", 16, 64, 48, 320, 320, 3, False),
    ("conv3x3 640->640 @32x24", 16, 32, 24, 640, 640, 3, False),
    ("conv3x3 1280->1280 @16x12", 16, 16, 12, 1280, 1280, 3, False),
    ("conv3x3 960->320 @64x48 (up-block concat width)", 16, 64, 48, 960, 320, 3, False),
    ("conv3x3 1920->640 @32x24", 16, 32, 24, 1920, 640, 3, False),
    ("gemm 49152x320 -> 320 +res (attn out-proj)", 16, 64, 48, 320, 320, 1, True),
    ("gemm 49152x320 -> 960 (qkv)", 16, 64, 48, 320, 960, 1, False),
    ("gemm 49152x320 -> 2560 GEGLU", 16, 64, 48, 320, 2560, 1, False),
    ("gemm 49152x1280 -> 320 +res (ff out)", 16, 64, 48, 1280, 320, 1, True),
    ("gemm 12288x640 -> 5120 GEGLU", 16, 32, 24, 640, 5120, 1, False),
    ("gemm 12288x2560 -> 640 +res", 16, 32, 24, 2560, 640, 1, True),
]


def main():
    from ladi_vton_b200 import ops, weights
    dev = torch.device("cuda:0")
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    g = torch.Generator().manual_seed(0)
    for label, n, h, w, cin, cout, ks, res in SHAPES:
        x = torch.randn((n, h, w, cin), generator=g).to(dev, torch.bfloat16)
        wt = torch.randn((cout, cin, ks, ks), generator=g) * (ks * ks * cin) ** -0.5
        wp = (weights.pack_conv(wt, [cin]) if ks == 3 else weights.pack_linear(wt.view(cout, cin))).to(dev)
        b = torch.randn(cout, generator=g).to(dev)
        geglu = "GEGLU" in label
        r = torch.randn((n, h, w, cout), generator=g).to(dev, torch.bfloat16) if res else None
        out = torch.empty((n, h, w, cout // 2 if geglu else cout), dtype=torch.bfloat16, device=dev)
        line = dict(shape=label, gflop=round(2.0 * n * h * w * cout * cin * ks * ks / 1e9, 1))
        outs = {}
        for name, pair in (("single", False), ("pair", True)):
            def run():
                ops.conv2d([x], wp, cout, ksize=ks, bias=b, residual=r, act=ops.ACT_GEGLU if geglu else ops.ACT_NONE, out=out, pair=pair,
                           split_k=False)
            for _ in range(5):
                run()
            ts = []
            for _ in range(20):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            us = ts[len(ts) // 2] * 1e3
            line[name + "_us"] = round(us, 1)
            line[name + "_tflops"] = round(line["gflop"] * 1e9 / (us * 1e-6) / 1e12, 1)
            outs[name] = out.clone()
        line["bit_identical"] = bool(torch.equal(outs["single"], outs["pair"]))
        line["speedup"] = round(line["single_us"] / line["pair_us"], 3)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
