"""ORACLE (test infrastructure only) -- EMASC, mask_features, DDIM scheduler, inversion adapter,
mask/image preparation; fp32 CPU restatements.

Each function cites the reference line range it follows.  diffusers/transformers pieces are
restated from SURVEY.md Appendix A.6/A.8 (upstream source absent -> "parity unpinned" for those;
pinned by the DDIM timestep/alpha known answers in tests/test_oracle_pins.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class EMASC(nn.Module):
    """/root/reference/src/models/emasc.py:11-40 ('nonlinear' type; state-dict keys conv.{i}.{0,2}.*)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.ModuleList([
            nn.Sequential(nn.Conv2d(ci, ci, 3, padding=1), nn.SiLU(), nn.Conv2d(ci, co, 3, padding=1))
            for ci, co in zip(in_channels, out_channels)])

    def forward(self, feats):
        return [m(f) for m, f in zip(self.conv, feats)]


def mask_features(feats, mask):
    """/root/reference/src/utils/data_utils.py:4-16 -- nearest resizes are CHAINED (each starts from
    the previous result)."""
    out = []
    for f in feats:
        mask = F.interpolate(mask, size=f.shape[-2:])
        out.append(f * (1 - mask))
    return out


def prepare_mask_and_masked_image(image, mask):
    """diffusers 0.14 pipeline_stable_diffusion_inpaint.prepare_mask_and_masked_image, tensor branch
    (called at /root/reference/src/vto_pipelines/tryon_pipe.py:630).  Binarises `mask` IN PLACE."""
    if image.ndim == 3:
        image = image.unsqueeze(0)
    if mask.ndim == 2:
        mask = mask.unsqueeze(0).unsqueeze(0)
    if mask.ndim == 3:
        mask = mask.unsqueeze(0) if mask.shape[0] == 1 else mask.unsqueeze(1)
    assert image.ndim == 4 and mask.ndim == 4 and image.shape[-2:] == mask.shape[-2:] and image.shape[0] == mask.shape[0]
    if image.min() < -1 or image.max() > 1:
        raise ValueError("Image should be in [-1, 1] range")
    if mask.min() < 0 or mask.max() > 1:
        raise ValueError("Mask should be in [0, 1] range")
    mask[mask < 0.5] = 0
    mask[mask >= 0.5] = 1
    image = image.to(dtype=torch.float32)
    return mask, image * (mask < 0.5)


def randn_tensor(shape, generator=None, dtype=None):
    """diffusers.utils.randn_tensor (recalled, Appendix A.8), CPU: a LIST of generators draws each sample's (1, ...) slice from
    its own generator; a one-element list is unwrapped."""
    if isinstance(generator, (list, tuple)):
        if len(generator) == 1:
            generator = generator[0]
        else:
            return torch.cat([torch.randn((1,) + tuple(shape[1:]), generator=g, dtype=dtype) for g in generator], dim=0)
    return torch.randn(tuple(shape), generator=generator, dtype=dtype)


class _Step:
    def __init__(self, prev_sample, pred_original_sample):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


class DDIMScheduler:
    """diffusers 0.14 DDIMScheduler with the SD-2-inpainting scheduler config (Appendix A.6);
    created at /root/reference/src/inference.py:123-124, used at tryon_pipe.py:650-651,722,740."""
    order = 1
    init_noise_sigma = 1.0

    class _C(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1,
                 clip_sample=False, set_alpha_to_one=False):
        self.config = self._C(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                              beta_schedule="scaled_linear", steps_offset=steps_offset, clip_sample=clip_sample,
                              set_alpha_to_one=set_alpha_to_one, prediction_type="epsilon", skip_prk_steps=True)
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, n, device=None):
        self.num_inference_steps = n
        ratio = self.config.num_train_timesteps // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, eta=0.0, generator=None):
        t = int(timestep)
        tp = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[tp] if tp >= 0 else self.final_alpha_cumprod
        x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        if self.config.clip_sample:
            x0 = x0.clamp(-1, 1)
        # diffusers 0.14 DDIMScheduler._get_variance / step (recalled): sigma_t = eta * sqrt((1-a_prev)/(1-a_t) * (1 - a_t/a_prev));
        # the CLI never sets eta, so sigma_t = 0 on the benchmarked path
        variance = (1 - a_p) / (1 - a_t) * (1 - a_t / a_p)
        std = eta * variance ** 0.5
        direction = (1 - a_p - std ** 2) ** 0.5 * model_output
        prev = a_p ** 0.5 * x0 + direction
        if eta > 0:
            prev = prev + std * randn_tensor(model_output.shape, generator=generator, dtype=model_output.dtype)
        return _Step(prev, x0)


class ClipEncoderLayer(nn.Module):
    """transformers 4.27.3 CLIPEncoderLayer (pre-LN, GELU(erf) for ViT-H 'gelu'), written out; the
    reference instantiates it at /root/reference/src/models/inversion_adapter.py:9."""

    def __init__(self, dim, heads, mlp, eps=1e-5):
        super().__init__()
        self.heads = heads
        self.self_attn = nn.Module()
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            setattr(self.self_attn, n, nn.Linear(dim, dim))
        self.layer_norm1 = nn.LayerNorm(dim, eps=eps)
        self.mlp = nn.Module()
        self.mlp.fc1 = nn.Linear(dim, mlp)
        self.mlp.fc2 = nn.Linear(mlp, dim)
        self.layer_norm2 = nn.LayerNorm(dim, eps=eps)

    def forward(self, x):
        b, n, c = x.shape
        y = self.layer_norm1(x)
        sp = lambda t: t.view(b, n, self.heads, c // self.heads).transpose(1, 2)
        a = self.self_attn
        o = F.scaled_dot_product_attention(sp(a.q_proj(y)), sp(a.k_proj(y)), sp(a.v_proj(y)))
        x = x + a.out_proj(o.transpose(1, 2).reshape(b, n, c))
        return x + self.mlp.fc2(F.gelu(self.mlp.fc1(self.layer_norm2(x))))


class InversionAdapter(nn.Module):
    """/root/reference/src/models/inversion_adapter.py:5-28 with dims of hubconf.py:16-27
    (1280 -> 5120 -> 5120 -> 16384; one ViT-H encoder layer: 16 heads, mlp 5120)."""

    def __init__(self, input_dim=1280, hidden_dim=5120, output_dim=16384, heads=16, mlp_dim=5120, num_encoder_layers=1):
        super().__init__()
        self.encoder_layers = nn.ModuleList([ClipEncoderLayer(input_dim, heads, mlp_dim) for _ in range(num_encoder_layers)])
        self.post_layernorm = nn.LayerNorm(input_dim, eps=1e-5)
        self.layers = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.GELU(), nn.Dropout(0.5),
                                    nn.Linear(hidden_dim, hidden_dim), nn.GELU(), nn.Dropout(0.5),
                                    nn.Linear(hidden_dim, output_dim))

    def forward(self, x):
        for layer in self.encoder_layers:
            x = layer(x)
        return self.layers(self.post_layernorm(x[:, 0, :]))
