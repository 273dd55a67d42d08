"""Localise errors of the wide single-head attention kernel: per (query tile, value half) max error for several shapes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ladi_vton_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for D in (512, 256):
    for nq, nkv, kscale in ((128, 128, 2.0), (128, 256, 2.0), (256, 128, 2.0), (128, 256, 0.05), (128, 384, 2.0), (256, 256, 2.0), (128, 3072, 2.0)):
        g = torch.Generator().manual_seed(1)
        q = torch.randn((1, nq, D), generator=g).to(dev).bfloat16()
        k = (torch.randn((1, nkv, D), generator=g) * kscale).to(dev).bfloat16()
        v = torch.randn((1, nkv, D), generator=g).to(dev).bfloat16()
        y = ops.attention_d512(q, k, v, D ** -0.5).float()
        torch.cuda.synchronize()
        ref = torch.softmax((q[0].float() @ k[0].float().t()) * D ** -0.5, dim=-1) @ v[0].float()
        err = (y[0] - ref).abs()
        line = f"D={D} nq={nq} nkv={nkv} kscale={kscale}: max|ref|={ref.abs().max():.3f} "
        for qt in range(nq // 128):
            for vh in range(D // 256):
                e = err[qt * 128:(qt + 1) * 128, vh * 256:(vh + 1) * 256]
                line += f" [q{qt} v{vh}: {e.max():.4f}]"
        # which tile's contribution is missing?  compare with partial references
        if nkv == 256:
            s = (q[0].float() @ k[0].float().t()) * D ** -0.5
            p = torch.softmax(s, dim=-1)
            only0 = (p[:, :128] @ v[0, :128].float()) / p[:, :128].sum(-1, keepdim=True)
            only1 = (p[:, 128:] @ v[0, 128:].float()) / p[:, 128:].sum(-1, keepdim=True)
            line += f" | vs tile0-only {(y[0] - only0).abs().max():.4f} vs tile1-only {(y[0] - only1).abs().max():.4f} vs unnormalised-sum {(y[0] - (only0 + only1)).abs().max():.4f}"
        print(line, flush=True)
