"""Per-phase cycle trace (clock64) of CTA (0,0,0) of the pair attention kernel (lazy single-pass softmax, variant 5): the four softmax
warpgroup halves (tile A / tile B x key-column halves) and the MMA issuer.  Needs the instrumented build:

    make -C ladi_vton_b200/csrc EXTRA=-DLADI_ATTN_TRACE BUILD=../../build/trace OUT=../libladi_b200_trace.so
    LADI_B200_LIB=ladi_vton_b200/libladi_b200_trace.so python tools/attn_trace.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladi_vton_b200 import ops
dev = torch.device("cuda:0")
q = torch.randn((16, 3072, 960), device=dev).bfloat16()
tr = torch.zeros(8192, dtype=torch.int64, device=dev)
for _ in range(2):
    ops.attention(q[..., :320], q[..., 320:640], q[..., 640:], 5, 0.125, variant=int(sys.argv[1]) if len(sys.argv) > 1 else 5, trace=tr)
torch.cuda.synchronize()
t = tr.cpu().numpy().astype("int64")
t0 = t[t > 0].min()
names = {2: "A.half0", 3: "A.half1", 4: "B.half0", 5: "B.half1"}
for k, nm in names.items():
    print(f"== softmax {nm}: per tile (variant 5) [wait s_full | ld0 | math0+st0 | ld1 | math1+st1 | wait st+arrive | max exchange]; (variant 6) "
          f"[wait s_full | first ld + o_ready | wait token | exps 0-31 | exps 32-63 | st+arrive | max exchange] (cycles), start rel")
    tot = []
    for j in range(0, 24):
        s = t[1024 * k + 8 * j: 1024 * k + 8 * j + 8]
        if s[0] == 0:
            continue
        d = [int(s[i + 1] - s[i]) for i in range(7)]
        tot.append(d)
        print(f"  j={j:2d} start {int(s[0] - t0):8d}  {d}  total {int(s[7] - s[0])}")
    if len(tot) > 4:
        import numpy as np
        print("  mean over j>=2:", [int(x) for x in np.array(tot[2:]).mean(0)], "sum", int(np.array(tot[2:]).sum(1).mean()))
for wg, nm in ((0, "A (warp 1)"), (1, "B (warp 3)")):
    print(f"== MMA issuer of tile {nm}: per KV tile [wait p_full | issue P V + commits | wait kv_full(j+1) | issue S(j+1)] (cycles), start rel")
    for j in range(0, 24):
        s = t[1024 * (6 + wg) + 8 * j: 1024 * (6 + wg) + 8 * j + 5]
        if s[0] == 0:
            continue
        print(f"  j={j:2d} start {int(s[0] - t0):8d}  {[int(s[i + 1] - s[i]) if s[i + 1] and s[i] else -1 for i in range(4)]}")
