"""The same as tools/ncu_kernels.py for the kernels of the widened rows: CLIP front-end (csrc/frontend.cu), warping module
(csrc/warp.cu), pose heat-maps -- one launch each at the shapes src/inference.py runs them (batch 8, 512x384 images, 256x192 warping grid).
    ncu --set full --clock-control none --profile-from-start off -o /tmp/r02_frontend python tools/ncu_kernels_frontend.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ladi_vton_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
r = lambda *s: torch.randn(s, device=dev).bfloat16()
f = lambda *s: torch.randn(s, device=dev)
B = 8
src = torch.randint(0, 49408, (B * 77,), device=dev, dtype=torch.int32)
tok, pos_t, wemb = r(49408, 1024), r(77, 1024), r(B * 16, 1024)
pixels = f(B, 3, 224, 224)
patch, cls, pos_v = r(B * 256, 1280), r(1280), r(257, 1280)
cloth = f(B, 3, 512, 384).clamp(-1, 1)
mean, std = torch.tensor([0.48, 0.46, 0.41], device=dev), torch.tensor([0.27, 0.26, 0.28], device=dev)
x64 = r(B, 256, 192, 64)
fa, fb = r(B, 16, 12, 512), r(B, 16, 12, 512)
theta = f(B, 50)
n_ctrl = 25
inv_k, tcr = f(n_ctrl + 3, n_ctrl + 3), f(256 * 192, n_ctrl + 3)
low_grid = (torch.rand((B, 256, 192, 2), device=dev) * 2 - 1)
warp_out = torch.zeros((B, 512, 384, 8), dtype=torch.bfloat16, device=dev)
x24 = r(B, 512, 384, 24)
xf = f(B, 512, 384, 8)
kp = torch.rand((B, 18, 2), device=dev) * 300


def everything():
    ops.clip_embed(src, tok, wemb, pos_t, 77)
    ops.patchify(pixels, 14, 640)
    ops.vit_assemble(patch, cls, pos_v, B)
    ops.clip_preprocess(cloth, 224, 224, mean, std, quantise=True)
    ops.resize_aa(cloth, 256, 192)
    ops.space_to_depth2(x64)
    ops.channel_affine_(x64, f(64), f(64))
    ops.l2norm_channels_(fa)
    ops.feature_correlation(fa, fb)
    ops.tps_grid(theta, inv_k, tcr, n_ctrl)
    ops.warp_grid_sample(low_grid, cloth, warp_out)
    ops.maxpool2(x24)
    ops.upsample2x_bilinear_ac(x64)
    ops.nhwc_f32_to_nchw_clamp(xf, 3, -1.0, 1.0)
    ops.pose_heatmaps(kp, 512, 384)


everything()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
everything()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
