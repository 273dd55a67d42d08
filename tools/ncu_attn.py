import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladi_vton_b200 import ops
dev = torch.device("cuda:0")
q = torch.randn((16, 3072, 960), device=dev).bfloat16()
for _ in range(2):
    ops.attention(q[..., :320], q[..., 320:640], q[..., 640:], 5, 0.125)
torch.cuda.synchronize()
