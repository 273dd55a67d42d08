"""Read the attention probabilities out of the wide-head kernel: V = one-hot(key) so that out[row, key] = P[row, key]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ladi_vton_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
D, nq, nkv = 512, 128, 256
g = torch.Generator().manual_seed(1)
for kscale in (0.0, 1.0):
    q = torch.randn((1, nq, D), generator=g).to(dev).bfloat16()
    k = (torch.randn((1, nkv, D), generator=g) * kscale).to(dev).bfloat16()
    v = torch.zeros((1, nkv, D), device=dev, dtype=torch.bfloat16)
    v[0, torch.arange(nkv), torch.arange(nkv)] = 1.0
    y = ops.attention_d512(q, k, v, D ** -0.5).float()[0]
    torch.cuda.synchronize()
    ref = torch.softmax((q[0].float() @ k[0].float().t()) * D ** -0.5, dim=-1)
    got = y[:, :nkv]
    print(f"kscale={kscale}: row sums got {got.sum(-1)[:4].tolist()} (expect 1)")
    for t in range(2):
        sl = slice(128 * t, 128 * (t + 1))
        e = (got[:, sl] - ref[:, sl]).abs()
        print(f"  keys of tile {t}: max err {e.max():.5f}  mean got {got[:, sl].mean():.5f} mean ref {ref[:, sl].mean():.5f}")
    r = 5
    print("  row 5 got  tile0[:6]", [round(x, 5) for x in got[r, :6].tolist()], "tile1[:6]", [round(x, 5) for x in got[r, 128:134].tolist()])
    print("  row 5 ref  tile0[:6]", [round(x, 5) for x in ref[r, :6].tolist()], "tile1[:6]", [round(x, 5) for x in ref[r, 128:134].tolist()])
    # per 16-key group error of tile 1, and per row-quarter
    e1 = (got[:, 128:256] - ref[:, 128:256]).abs()
    print("  tile1 err per 16-key group:", [round(float(e1[:, 16 * i:16 * i + 16].max()), 4) for i in range(8)])
    print("  tile1 err per 32-row group:", [round(float(e1[32 * i:32 * i + 32].max()), 4) for i in range(4)])
    e0 = (got[:, :128] - ref[:, :128]).abs()
    print("  tile0 err per 16-key group:", [round(float(e0[:, 16 * i:16 * i + 16].max()), 4) for i in range(8)])
    print("  other columns (256..511) max:", float(y[:, 256:].abs().max()))
