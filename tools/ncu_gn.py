"""Driver for `ncu --set full --cache-control none` on the HBM/L2-bound normalisation kernels at the UNet level-1 shape
(batch 16, 64x48, 320 channels): producer conv -> gn_stats -> gn_apply -> layernorm, plus an isolated CUDA-event timing loop and a
same-size torch copy as the bandwidth yardstick (printed when run without ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ladi_vton_b200 import ops, weights  # noqa: E402

dev = torch.device("cuda:0")
r = lambda *s: torch.randn(s, device=dev).bfloat16()
x = r(16, 64, 48, 320)
w = weights.pack_conv(torch.randn(320, 320, 3, 3, device=dev) * 0.01, [320])
b = torch.zeros(320, device=dev)
g, be = torch.ones(320, device=dev), torch.zeros(320, device=dev)
ws = ops.GroupNormWS(dev)
y = None
for _ in range(2):
    h = ops.conv2d([x], w, 320, bias=b)
    y = ops.groupnorm([h], g, be, 32, 1e-5, ws, silu=True)
    z = ops.layernorm(h.view(-1, 320), g, be)
torch.cuda.synchronize()


def timed(fn, n=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


if os.environ.get("NCU_GN_TIMING", "1") == "1":
    h = ops.conv2d([x], w, 320, bias=b)
    dst = torch.empty_like(h)
    print("groupnorm (stats+apply) us:", round(timed(lambda: ops.groupnorm([h], g, be, 32, 1e-5, ws, silu=True)), 2))
    print("layernorm us:", round(timed(lambda: ops.layernorm(h.view(-1, 320), g, be)), 2))
    print("torch copy 31MB->31MB us:", round(timed(lambda: dst.copy_(h)), 2))
    print("torch sum 31MB us:", round(timed(lambda: h.float().sum() if False else torch.sum(h, dtype=torch.float32)), 2))
    big = r(16, 256, 192, 128)  # 201 MB: larger than L2
    dstb = torch.empty_like(big)
    print("torch copy 201MB us:", round(timed(lambda: dstb.copy_(big)), 2), "->", "GB/s", round(2 * big.numel() * 2 / timed(lambda: dstb.copy_(big)) / 1e3, 1))
    gb, bb = torch.ones(128, device=dev), torch.zeros(128, device=dev)
    print("groupnorm 201MB us:", round(timed(lambda: ops.groupnorm([big], gb, bb, 32, 1e-6, ws, silu=True)), 2))
