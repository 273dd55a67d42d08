"""One launch of EVERY kernel of libladi_b200.so at the shapes the bench runs (BASELINE configs[1]: UNet batch 16 @ 64x48 latents; VAE batch 8 @
512x384), for ONE `ncu --set full` capture ("each kernel ships a committed ncu capture"):

    ncu --set full --clock-control none --import-source on -o gpurun_out/r02_kernels python tools/ncu_kernels.py
    ncu -i gpurun_out/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv ; python tools/ncu_summarize.py ... > profiles/r02_ncu_kernels.txt

Each op runs once as warm-up OUTSIDE the profiled range (cudaProfilerStart/Stop brackets the second pass), so the capture holds one launch per
kernel instance."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ladi_vton_b200 import ops, weights  # noqa: E402

dev = torch.device("cuda:0")
r = lambda *s: torch.randn(s, device=dev).bfloat16()
f = lambda *s: torch.randn(s, device=dev)
B = 16
ws = ops.GroupNormWS(dev)
conv_w = lambda co, ci: weights.pack_conv(torch.randn(co, ci, 3, 3, device=dev) * (9 * ci) ** -0.5, [ci])
lin_w = lambda co, ci: weights.pack_linear(torch.randn(co, ci, device=dev) * ci ** -0.5)
x320, x640, x1280, x8 = r(B, 64, 48, 320), r(B, 32, 24, 640), r(B, 16, 12, 1280), r(B, 8, 6, 1280)
w320, w640, w1280 = conv_w(320, 320), conv_w(640, 640), conv_w(1280, 1280)
b320, b640, b1280 = f(320), f(640), f(1280)
a320, res320 = r(B * 3072, 320), r(B * 3072, 320)
l320, lqkv, lff1 = lin_w(320, 320), lin_w(960, 320), lin_w(2560, 320)
bff1 = f(2560)
up_w = weights.pack_conv_up2x(torch.randn(640, 640, 3, 3, device=dev) * (9 * 640) ** -0.5, [640])
qkv = r(B, 3072, 960)
q77, kv77 = r(B, 3072, 320), r(B, 77, 640)
qkv192 = r(B, 192, 3 * 1280)
qkv512 = r(8, 3072, 1536)
g320, be320 = torch.ones(320, device=dev), torch.zeros(320, device=dev)
vae128 = r(8, 512, 384, 128)
wv128 = conv_w(128, 128)
b128 = f(128)
w_out = weights.pack_conv(torch.randn(4, 320, 3, 3, device=dev) * 0.02, [320])
eps = torch.empty((B, 64, 48, 4), dtype=torch.float32, device=dev)
lat = f(8, 4, 64, 48)
uin = torch.zeros((B, 64, 48, 32), dtype=torch.bfloat16, device=dev)
coef = torch.rand((50, 8), device=dev)
step = torch.zeros(2, dtype=torch.int32, device=dev)
img, mask, pose = f(8, 3, 512, 384).clamp(-1, 1), (f(8, 1, 512, 384) > 0).float(), f(8, 18, 512, 384).abs()
mom = f(8, 64, 48, 8)
dec = torch.randn((8, 512, 384, 4), device=dev)
flags = torch.zeros(2, dtype=torch.int32, device=dev)
s_rows = f(3072, 3072)
cls_q, cls_kv = r(8, 1280), r(8, 257, 2560)
sq = r(8, 257, 3 * 1280)


def everything():
    ops.conv2d([x640], w640, 640, bias=b640)                                          # convgemm<256, pair>: conv3x3 640->640 @32x24
    ops.conv2d([x320], w320, 320, bias=b320)                                          # convgemm<160, pair>: conv3x3 320->320 @64x48
    ops.gemm(a320, l320, 320, bias=b320, residual=res320)                             # convgemm<128, pair>: the K=320 GEMM with residual
    ops.gemm(a320, lqkv, 960)                                                         # convgemm<192, pair>: QKV projection
    ops.gemm(a320, lff1, 2560, bias=bff1, act=ops.ACT_GEGLU)                          # GEGLU epilogue
    ops.conv2d([x8], w1280, 1280, bias=b1280)                                         # split-K (8x6 level) + splitk_reduce_kernel
    ops.conv2d([x640], w640, 640, bias=b640, stride=2, pad_lo=1)                      # stride-2 (parity planes)
    ops.conv2d([x640], up_w, 640, bias=b640, up2x=True)                               # fused nearest-2x + conv3x3 (sub-pixel)
    ops.conv2d([x320], w_out, 4, bias=f(4), out=eps, out_fp32=True)                   # conv_out: N = 4, fp32, direct epilogue
    ops.conv2d([vae128], wv128, 128, bias=b128)                                       # VAE full-resolution 128->128
    ops.attention(qkv[..., :320], qkv[..., 320:640], qkv[..., 640:], 5, 0.125)        # attention_pair_kernel (3072 tokens)
    ops.attention(q77, kv77[..., :320], kv77[..., 320:], 5, 0.125)                    # attention_single_kernel<SHORT> (77 text tokens)
    ops.attention(qkv192[..., :1280], qkv192[..., 1280:2560], qkv192[..., 2560:], 20, 0.125)  # attention_single_kernel (192 tokens)
    ops.attention_d512(qkv512[..., :512], qkv512[..., 512:1024], qkv512[..., 1024:], 512 ** -0.5)  # VAE mid-block attention
    ops.groupnorm([x320], g320, be320, 32, 1e-5, ws, silu=True)                       # gn_stats + gn_apply
    ops.groupnorm([x320], g320, be320, 32, 1e-5, ws, silu=True, add=x320)             # gn_apply<add>
    ops.layernorm(a320, g320, be320)
    ops.ddim_cfg_step(eps, lat, uin, True, 7.5, coef, step)
    ops.add(x320, x320)
    ops.upsample2x(x640)
    ops.check_binarise_(img, mask, flags)
    ops.nchw_to_nhwc(img, torch.zeros((8, 512, 384, 8), dtype=torch.bfloat16, device=dev), gate=mask)
    ops.nchw_to_nhwc(lat, uin[:8], c_off=0)
    ops.nhwc_to_nchw(eps, 4)
    ops.bilinear_down8(pose)
    ops.inv_mask_rows(mask, 8)
    ops.posterior_sample(mom, lat, 0.18215)
    ops.image_out(dec)
    ops.image_out_u8(dec)
    ops.softmax_rows(s_rows, 512 ** -0.5)
    ops.cls_attention(cls_q, cls_kv, 16, 80, 80 ** -0.5)
    ops.attention_small(sq[..., :1280], sq[..., 1280:2560], sq[..., 2560:], 16, 80 ** -0.5)


everything()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
everything()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
