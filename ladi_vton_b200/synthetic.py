"""Synthetic workload of BASELINE.json / SURVEY.md section 8(d): person/garment/pose/mask tensors shaped like the outputs of
/root/reference/src/dataset/vitonhd.py (image, cloth in [-1,1]; 18 Gaussian pose heat-maps; binary inpaint mask) and
random-init weights of the reference architectures (hubconf.py:16-53), all seeded (CLI default seed 1234,
src/inference.py:57).  No dataset, checkpoint or network access is needed."""
import math

import torch

from .unet import unet_param_shapes
from .vae import vae_param_shapes

EMASC_IN, EMASC_OUT = [128, 128, 128, 256, 512], [128, 256, 512, 512, 512]  # hubconf.py:41-42


def emasc_channels(vae_ch):
    c = list(vae_ch)
    return [c[0], c[0], c[0], c[1], c[2]], [c[0], c[1], c[2], c[3], c[3]]


def emasc_param_shapes(cin, cout):
    S = {}
    for i, (a, b) in enumerate(zip(cin, cout)):
        S[f"conv.{i}.0.weight"], S[f"conv.{i}.0.bias"] = (a, a, 3, 3), (a,)
        S[f"conv.{i}.2.weight"], S[f"conv.{i}.2.bias"] = (b, a, 3, 3), (b,)
    return S


def random_state_dict(shapes, seed, device="cpu", fast=False):
    """PyTorch-default-like init (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for conv/linear weights and biases, ones/zeros for
    norms), drawn in key order from one seeded generator on `device` (CPU draws are reproducible across machines and are
    what the parity tests share between oracle and engine; bench.py draws on the GPU to skip the host cost)."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    fan = {}
    block = torch.empty(1 << 22, device=device).uniform_(-1.0, 1.0, generator=g) if fast else None
    for k, shp in shapes.items():
        base = k.rsplit(".", 1)[0]
        leaf = base.rsplit(".", 1)[-1]
        is_norm = leaf.startswith("norm") or leaf in ("group_norm", "conv_norm_out", "layer_norm1", "layer_norm2", "post_layernorm", "final_layer_norm", "pre_layrnorm")
        if is_norm:
            sd[k] = torch.ones(shp, device=device) if k.endswith("weight") else torch.zeros(shp, device=device)
            continue
        if k.endswith("weight"):
            fan[base] = math.prod(shp[1:])
            bound = 1.0 / math.sqrt(fan[base])
        else:
            bound = 1.0 / math.sqrt(fan.get(base, shp[0]))
        if fast:  # timing-only weights (CPU baseline legs): tile one random block instead of drawing ~1e9 values
            n = math.prod(shp)
            t = torch.empty(n, device=device)
            for o in range(0, n, block.numel()):  # in-place tiling: one pass over fresh memory, no temporaries
                m = min(block.numel(), n - o)
                torch.mul(block[:m], bound, out=t[o:o + m])
            sd[k] = t.view(shp)
        else:
            sd[k] = torch.empty(shp, device=device).uniform_(-bound, bound, generator=g)
    return sd


def synthetic_inputs(B, H, W, seed=1234, ctx_dim=1024, n_pose=18):
    g = torch.Generator().manual_seed(seed)
    image = torch.rand((B, 3, H, W), generator=g) * 2 - 1
    cloth = torch.rand((B, 3, H, W), generator=g) * 2 - 1
    mask = torch.zeros((B, 1, H, W))
    mh, mw = int(H * 0.6), int(W * 0.58)  # centred rectangle, ~35 % of the pixels
    mask[:, :, (H - mh) // 2:(H - mh) // 2 + mh, (W - mw) // 2:(W - mw) // 2 + mw] = 1.0
    ys = torch.arange(H, dtype=torch.float32)[None, None, :, None]
    xs = torch.arange(W, dtype=torch.float32)[None, None, None, :]
    cy = torch.rand((B, n_pose, 1, 1), generator=g) * H
    cx = torch.rand((B, n_pose, 1, 1), generator=g) * W
    pose = torch.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / 81.0)  # sigma = 9 as src/utils/posemap.py:29-31
    prompt = torch.randn((B, 77, ctx_dim), generator=g)
    negative = torch.randn((B, 77, ctx_dim), generator=g)
    return dict(image=image, mask_image=mask, pose_map=pose, warped_cloth=cloth, prompt_embeds=prompt, negative_prompt_embeds=negative)


SMALL_UNET = dict(block_out_channels=(64, 128, 256, 256), attention_head_dim=(1, 2, 4, 4), cross_attention_dim=128, sample_size=16)
SMALL_VAE = dict(block_out_channels=(64, 128, 256, 256))


def build_state_dicts(unet_cfg=None, vae_cfg=None, seed=1234, device="cpu"):
    unet_cfg, vae_cfg = unet_cfg or {}, vae_cfg or {}
    from .vae import SD2_VAE
    vch = {**SD2_VAE, **vae_cfg}["block_out_channels"]
    ein, eout = emasc_channels(vch)
    return dict(unet=random_state_dict(unet_param_shapes(unet_cfg), seed, device),
                vae=random_state_dict(vae_param_shapes(vae_cfg), seed + 1, device),
                emasc=random_state_dict(emasc_param_shapes(ein, eout), seed + 2, device),
                emasc_channels=(ein, eout))


def build_pipeline(device, unet_cfg=None, vae_cfg=None, seed=1234, sds=None, weights_on_device=False):
    """Random-init engine pipeline (the hubconf.py constructors' architectures, no checkpoint)."""
    from . import EMASC, AutoencoderKL, DDIMScheduler, StableDiffusionTryOnePipeline, UNet2DConditionModel
    sds = sds or build_state_dicts(unet_cfg, vae_cfg, seed, device if weights_on_device else "cpu")
    unet = UNet2DConditionModel(**(unet_cfg or {})).load_state_dict(sds["unet"])
    vae = AutoencoderKL(**(vae_cfg or {})).load_state_dict(sds["vae"])
    emasc = EMASC(*sds["emasc_channels"]).load_state_dict(sds["emasc"])
    pipe = StableDiffusionTryOnePipeline(vae=vae, text_encoder=None, tokenizer=None, unet=unet, scheduler=DDIMScheduler(),
                                         emasc=emasc, emasc_int_layers=[1, 2, 3, 4, 5])
    return pipe.to(device), sds
