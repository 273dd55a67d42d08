"""P read-out with a SHIFTED one-hot V: out[row, (key + shift) % 256] = P[row, key] -- separates key-side (P / k-step) from column-side (V chunk / O) faults."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ladi_vton_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
D = 512
g = torch.Generator().manual_seed(1)
for nkv in (256, 384):
    for shift in (0, 128, 64):
        for kscale in (0.0, 1.0):
            nq = 128
            q = torch.randn((1, nq, D), generator=g).to(dev).bfloat16()
            k = (torch.randn((1, nkv, D), generator=g) * kscale).to(dev).bfloat16()
            v = torch.zeros((1, nkv, D), device=dev, dtype=torch.bfloat16)
            keys = torch.arange(nkv)
            v[0, keys, (keys + shift) % 256] = 1.0
            y = ops.attention_d512(q, k, v, D ** -0.5).float()[0]
            torch.cuda.synchronize()
            p = torch.softmax((q[0].float() @ k[0].float().t()) * D ** -0.5, dim=-1)
            ref = torch.zeros((nq, 256), device=dev)
            ref.index_add_(1, ((keys + shift) % 256).to(dev), p)
            e = (y[:, :256] - ref).abs()
            cols = [round(float(e[:, 32 * i:32 * i + 32].max()), 4) for i in range(8)]
            rows = [round(float(e[32 * i:32 * i + 32].max()), 4) for i in range(4)]
            print(f"nkv={nkv} shift={shift} kscale={kscale}: err per 32-col group {cols} per 32-row group {rows} rowsum {float(y[0, :256].sum()):.3f}", flush=True)
