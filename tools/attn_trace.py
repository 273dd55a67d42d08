"""Per-phase cycle trace (clock64) of one CTA of the pair attention kernel: softmax warpgroups A/B (two halves each) and the MMA
thread.  Prints cycle deltas per KV tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladi_vton_b200 import ops
dev = torch.device("cuda:0")
q = torch.randn((16, 3072, 960), device=dev).bfloat16()
tr = torch.zeros(8192, dtype=torch.int64, device=dev)
for _ in range(2):
    ops.attention(q[..., :320], q[..., 320:640], q[..., 640:], 5, 0.125, variant=int(sys.argv[1]) if len(sys.argv) > 1 else 2, trace=tr)
torch.cuda.synchronize()
t = tr.cpu().numpy().astype("int64")
t0 = t[t > 0].min()
names = {2: "A.half0", 3: "A.half1", 4: "B.half0", 5: "B.half1"}
for k, nm in names.items():
    print(f"== softmax {nm}: per tile [wait s_full | pass1 max | exchange | pass2 exp | wait o_ready | P store | fence+arrive] (cycles), start rel")
    for j in range(0, 24):
        s = t[1024 * k + 8 * j: 1024 * k + 8 * j + 8]
        if s[0] == 0:
            continue
        d = [int(s[i + 1] - s[i]) for i in range(7)]
        print(f"  j={j:2d} start {int(s[0] - t0):8d}  {d}  total {int(s[7] - s[0])}")
print("== MMA thread: per iter [wait pA | issue PV_A,S_A(+kv wait) | wait pB | issue PV_B,S_B] (cycles)")
for j in range(0, 25):
    s = t[1024 * 6 + 8 * j: 1024 * 6 + 8 * j + 5]
    if s[0] == 0:
        continue
    s8 = t[1024 * 6 + 8 * j: 1024 * 6 + 8 * j + 8]
    fine = [int(s8[5] - s8[1]), int(s8[6] - s8[5]), int(s8[7] - s8[6]), int(s8[2] - s8[7])] if s8[7] else []
    print(f"  j={j:2d} start {int(s[0] - t0):8d}  {[int(s[i + 1] - s[i]) if s[i + 1] and s[i] else -1 for i in range(4)]}   A-part: [PV issue, commit, kv_full wait, S issue+commit] = {fine}")
