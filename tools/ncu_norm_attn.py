"""Driver for one `ncu --set full` capture of the HBM/L2-bound normalisation kernels (gn_stats, gn_apply, layernorm at the UNet level-0
shape: batch 16, 64x48, 320 channels) and the self-attention kernel (3072 tokens, 5 heads):
    ncu --set full --clock-control none --import-source on -k regex:"gn_|layernorm|attention" -c 8 -o out python tools/ncu_norm_attn.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladi_vton_b200 import ops
dev = torch.device("cuda:0")
h = torch.randn((16, 64, 48, 320), device=dev).bfloat16()
g, be = torch.ones(320, device=dev), torch.zeros(320, device=dev)
ws = ops.GroupNormWS(dev)
q = torch.randn((16, 3072, 960), device=dev).bfloat16()
for _ in range(2):
    ops.groupnorm([h], g, be, 32, 1e-5, ws, silu=True)
    ops.layernorm(h.view(-1, 320), g, be)
    ops.attention(q[..., :320], q[..., 320:640], q[..., 640:], 5, 0.125)
torch.cuda.synchronize()
