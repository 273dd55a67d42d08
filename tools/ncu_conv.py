"""Tiny driver for `ncu --set full`: a few launches of the dominant kernel at UNet shapes (conv3x3 640->640 @32x24, batch 16;
conv3x3 320->320 @64x48; GEMM 49152x320x320 with residual) and of the attention kernel (self, 3072 tokens)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ladi_vton_b200 import ops, weights  # noqa: E402

dev = torch.device("cuda:0")
r = lambda *s: torch.randn(s, device=dev).bfloat16()
x = r(16, 32, 24, 640); w = weights.pack_conv(torch.randn(640, 640, 3, 3, device=dev) * 0.01, [640]); b = torch.zeros(640, device=dev)
x2 = r(16, 64, 48, 320); w2 = weights.pack_conv(torch.randn(320, 320, 3, 3, device=dev) * 0.01, [320]); b2 = torch.zeros(320, device=dev)
a = r(49152, 320); w3 = weights.pack_linear(torch.randn(320, 320, device=dev) * 0.05); res = r(49152, 320)
q = r(16, 3072, 960)
for _ in range(2):
    ops.conv2d([x], w, 640, bias=b)
    ops.conv2d([x2], w2, 320, bias=b2)
    ops.gemm(a, w3, 320, bias=b2, residual=res)
    ops.attention(q[..., :320], q[..., 320:640], q[..., 640:], 5, 0.125)
torch.cuda.synchronize()
