"""N-tile (BN) sweep of the conv/GEMM kernel on the UNet's GEMM / conv shapes, CTA-pair and single-CTA mode, same process (CUDA events,
median of 15 after 3 warm-ups, L2 flushed).  The library's own choice is the column "auto".  One JSON line per shape."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladi_vton_b200 import ops, weights

SHAPES = [  # (label, n, h, w, cin, cout, ksize, residual, geglu)
    ("gemm 49152x320->320 +res", 16, 64, 48, 320, 320, 1, True, False),
    ("gemm 49152x320->960", 16, 64, 48, 320, 960, 1, False, False),
    ("gemm 49152x320->2560 GEGLU", 16, 64, 48, 320, 2560, 1, False, True),
    ("gemm 49152x1280->320 +res", 16, 64, 48, 1280, 320, 1, True, False),
    ("gemm 12288x640->640 +res", 16, 32, 24, 640, 640, 1, True, False),
    ("gemm 12288x640->1920", 16, 32, 24, 640, 1920, 1, False, False),
    ("gemm 12288x640->5120 GEGLU", 16, 32, 24, 640, 5120, 1, False, True),
    ("gemm 12288x2560->640 +res", 16, 32, 24, 2560, 640, 1, True, False),
    ("gemm 3072x1280->1280 +res", 16, 16, 12, 1280, 1280, 1, True, False),
    ("gemm 3072x1280->10240 GEGLU", 16, 16, 12, 1280, 10240, 1, False, True),
    ("conv3x3 320->320 @64x48", 16, 64, 48, 320, 320, 3, False, False),
    ("conv3x3 640->640 @32x24", 16, 32, 24, 640, 640, 3, False, False),
    ("conv3x3 1280->1280 @16x12", 16, 16, 12, 1280, 1280, 3, False, False),
    ("conv3x3 128->128 @512x384 (VAE, batch 8)", 8, 512, 384, 128, 128, 3, False, False),
]
dev = torch.device("cuda:0")
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
g = torch.Generator().manual_seed(0)
for label, n, h, w, cin, cout, ks, res, geglu in SHAPES:
    x = torch.randn((n, h, w, cin), generator=g).to(dev, torch.bfloat16)
    wt = torch.randn((cout, cin, ks, ks), generator=g) * (ks * ks * cin) ** -0.5
    wp = (weights.pack_conv(wt, [cin]) if ks == 3 else weights.pack_linear(wt.view(cout, cin))).to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    r = torch.randn((n, h, w, cout), generator=g).to(dev, torch.bfloat16) if res else None
    out = torch.empty((n, h, w, cout // 2 if geglu else cout), dtype=torch.bfloat16, device=dev)
    line = dict(shape=label, gflop=round(2.0 * n * h * w * cout * cin * ks * ks / 1e9, 1))
    for pair in (True, False):
        for bn in (0, 128, 160, 192, 256):
            if geglu and bn not in (0, 128, 256):
                continue
            if bn > cout and bn != 0:
                continue
            def run():
                ops.conv2d([x], wp, cout, ksize=ks, bias=b, residual=r, act=ops.ACT_GEGLU if geglu else ops.ACT_NONE, out=out, pair=pair, split_k=False,
                           force_bn=bn)
            try:
                for _ in range(3):
                    run()
            except RuntimeError:
                continue
            ts = []
            for _ in range(15):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            line[("pair" if pair else "single") + "_" + ("auto" if bn == 0 else str(bn))] = round(ts[len(ts) // 2] * 1e3, 1)
    print(json.dumps(line), flush=True)
