"""ORACLE (test infrastructure only) -- fp32 CPU restatement of the LaDI-VTON VAE fork.

Follows /root/reference/src/models/AutoencoderKL.py:145-188 (encode returns the encoder skip
list, decode adds EMASC features) and /root/reference/src/models/vae.py:99-119 (Encoder.forward),
:183-212 (Decoder.forward), :329-348 (DiagonalGaussianDistribution).  The blocks those files
obtain from diffusers==0.14.0 (`get_down_block`, `get_up_block`, `UNetMidBlock2D`, vae.py:21-23)
are restated from SURVEY.md Appendix A.5.

PARITY STATUS: the in-repo control flow is pinned bit-exactly against the reference's own
files executed on the shim (tests/golden/make_golden.py, run in the build container); the
diffusers blocks are unpinned upstream, pinned by the known-answer parameter count 83,653,863.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet import Downsample2D, ResnetBlock2D, Upsample2D, _Cfg

SD2_VAE = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
               latent_channels=4, norm_num_groups=32, scaling_factor=0.18215, sample_size=512)


class AttentionBlock(nn.Module):
    """Single-head spatial self-attention of the VAE mid block (Appendix A.5)."""

    def __init__(self, c, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.query = nn.Linear(c, c)
        self.key = nn.Linear(c, c)
        self.value = nn.Linear(c, c)
        self.proj_attn = nn.Linear(c, c)

    def forward(self, x):
        b, c, h, w = x.shape
        y = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.query(y), self.key(y), self.value(y)
        p = torch.softmax((q @ k.transpose(1, 2)).float() * (c ** -0.5), dim=-1).to(q.dtype)
        y = self.proj_attn(p @ v)
        return y.transpose(1, 2).reshape(b, c, h, w) + x


class VaeMidBlock(nn.Module):
    def __init__(self, c, groups=32, eps=1e-6):
        super().__init__()
        self.attentions = nn.ModuleList([AttentionBlock(c, groups, eps)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, None, groups, eps), ResnetBlock2D(c, c, None, groups, eps)])

    def forward(self, x, temb=None):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, eps, add_down):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, None, groups, eps) for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, 0)]) if add_down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, eps, add_up):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, None, groups, eps) for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, cin, cz, ch, layers, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(cin, ch[0], 3, padding=1)
        blocks, out = [], ch[0]
        for i, c in enumerate(ch):
            prev, out = out, c
            blocks.append(DownEncoderBlock2D(prev, out, layers, groups, 1e-6, i < len(ch) - 1))
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = VaeMidBlock(ch[-1], groups)
        self.conv_norm_out = nn.GroupNorm(groups, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * cz, 3, padding=1)

    def forward(self, x):
        # skip list = [input, conv_in out, input of every down block]  (vae.py:100-109)
        feats = [x]
        h = self.conv_in(x)
        feats.append(h)
        for blk in self.down_blocks:
            feats.append(h)
            h = blk(h)
        h = self.mid_block(h)
        return self.conv_out(F.silu(self.conv_norm_out(h))), feats


class Decoder(nn.Module):
    def __init__(self, cz, cout, ch, layers, groups):
        super().__init__()
        rch = list(reversed(ch))
        self.conv_in = nn.Conv2d(cz, rch[0], 3, padding=1)
        self.mid_block = VaeMidBlock(rch[0], groups)
        blocks, out = [], rch[0]
        for i, c in enumerate(rch):
            prev, out = out, c
            blocks.append(UpDecoderBlock2D(prev, out, layers + 1, groups, 1e-6, i < len(ch) - 1))
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(groups, ch[0], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[0], cout, 3, padding=1)

    def forward(self, z, feats=None, int_layers=None):
        h = self.mid_block(self.conv_in(z))
        if feats:
            feats.reverse()  # in place, like vae.py:190
            for blk, f in zip(self.up_blocks, feats):
                h = blk(h + f)
        else:
            for blk in self.up_blocks:
                h = blk(h)
        h = F.silu(self.conv_norm_out(h))
        if int_layers and 1 in int_layers:
            h = h + feats[len(int_layers) - 1 - int_layers.index(1)]
        h = self.conv_out(h)
        if int_layers and 0 in int_layers:
            h = h + feats[len(int_layers) - 1 - int_layers.index(0)]
        return h


class DiagonalGaussian:
    def __init__(self, moments):
        self.parameters = moments
        self.mean, logvar = moments.chunk(2, dim=1)
        self.logvar = logvar.clamp(-30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        eps = torch.randn(self.mean.shape, generator=generator, dtype=self.parameters.dtype)
        return self.mean + self.std * eps.to(self.mean.device)

    def mode(self):
        return self.mean


class _Enc:
    def __init__(self, d):
        self.latent_dist = d


class _Dec:
    def __init__(self, s):
        self.sample = s


class AutoencoderKL(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        cfg = dict(SD2_VAE)
        cfg.update(kw)
        self.config = _Cfg(cfg)
        ch, L, g, cz = cfg["block_out_channels"], cfg["layers_per_block"], cfg["norm_num_groups"], cfg["latent_channels"]
        self.encoder = Encoder(cfg["in_channels"], cz, ch, L, g)
        self.decoder = Decoder(cz, cfg["out_channels"], ch, L, g)
        self.quant_conv = nn.Conv2d(2 * cz, 2 * cz, 1)
        self.post_quant_conv = nn.Conv2d(cz, cz, 1)

    @property
    def dtype(self):
        return self.quant_conv.weight.dtype

    def encode(self, x):
        h, feats = self.encoder(x)
        return _Enc(DiagonalGaussian(self.quant_conv(h))), feats

    def decode(self, z, intermediate_features=None, int_layers=None):
        z = self.post_quant_conv(z)
        return _Dec(self.decoder(z, intermediate_features, int_layers) if intermediate_features else self.decoder(z))
