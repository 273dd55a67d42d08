"""ORACLE (test infrastructure only) -- fp32 CPU restatement of diffusers==0.14.0 `UNet2DConditionModel`.

The reference builds this class from the SD-2-inpainting config with `in_channels=31`
(/root/reference/hubconf.py:30-33) and calls it at
/root/reference/src/vto_pipelines/tryon_pipe.py:732.  diffusers itself is an un-vendored
third-party dependency (environment.yml:37) that is absent from this image, so this file
restates its published algorithm from the spec in SURVEY.md Appendix A.1-A.4.

PARITY STATUS: *unpinned against upstream diffusers* (no golden vectors exist in the reference,
SURVEY.md section 4).  Pinned instead by known answers: parameter counts 865,910,724 (in=4) /
865,988,484 (in=31) and the state-dict key set of Appendix A.7 (tests/test_oracle_pins.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (ladi_vton_b200/) never does.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

SD2_INPAINT_UNET = dict(
    in_channels=31, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024, norm_num_groups=32, norm_eps=1e-5,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    sample_size=64,
)


def timestep_embedding(t, dim=320, max_period=10000.0):
    """get_timestep_embedding with flip_sin_to_cos=True, freq_shift=0 (Appendix A.2 step 1)."""
    half = dim // 2
    freq = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    arg = t.float()[:, None] * freq[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


class ResnetBlock2D(nn.Module):
    """Appendix A.3.  temb_channels=None for the VAE flavour."""

    def __init__(self, cin, cout, temb_channels=1280, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, cout) if temb_channels else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class CrossAttention(nn.Module):
    def __init__(self, dim, ctx_dim, heads):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim)])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        b, n, c = x.shape
        hd = c // self.heads
        q = self.to_q(x).view(b, n, self.heads, hd).transpose(1, 2)
        k = self.to_k(ctx).view(b, -1, self.heads, hd).transpose(1, 2)
        v = self.to_v(ctx).view(b, -1, self.heads, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)  # softmax(q k^T / sqrt(hd)) v, no mask
        return self.to_out[0](o.transpose(1, 2).reshape(b, n, c))


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, 2 * inner)

    def forward(self, x):
        a, g = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(g)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, 4 * dim), nn.Identity(), nn.Linear(4 * dim, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, ctx_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = CrossAttention(dim, dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = CrossAttention(dim, ctx_dim, heads)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    """Appendix A.4 (use_linear_projection=True, one layer)."""

    def __init__(self, dim, heads, ctx_dim, groups=32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, ctx_dim)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, ctx):
        b, c, h, w = x.shape
        y = self.norm(x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        y = self.proj_in(y)
        y = self.transformer_blocks[0](y, ctx)
        y = self.proj_out(y)
        return y.reshape(b, h, w, c).permute(0, 3, 1, 2) + x


class Downsample2D(nn.Module):
    def __init__(self, c, padding=1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:  # VAE flavour: asymmetric pad right/bottom
            x = F.pad(x, (0, 1, 0, 1))
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, layers, temb, heads, ctx_dim, groups, eps, add_down, cross):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups, eps) for i in range(layers)])
        if cross:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, ctx_dim, groups) for _ in range(layers)])
        self.cross = cross
        if add_down:
            self.downsamplers = nn.ModuleList([Downsample2D(cout, 1)])
        self.add_down = add_down

    def forward(self, x, temb, ctx):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.cross:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.add_down:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class UpBlock(nn.Module):
    def __init__(self, cin, cout, cprev, layers, temb, heads, ctx_dim, groups, eps, add_up, cross):
        super().__init__()
        rs = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            rin = cprev if i == 0 else cout
            rs.append(ResnetBlock2D(rin + skip, cout, temb, groups, eps))
        self.resnets = nn.ModuleList(rs)
        if cross:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, ctx_dim, groups) for _ in range(layers)])
        self.cross = cross
        if add_up:
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])
        self.add_up = add_up

    def forward(self, x, skips, temb, ctx):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)  # current first, skip second
            x = r(x, temb)
            if self.cross:
                x = self.attentions[i](x, ctx)
        if self.add_up:
            x = self.upsamplers[0](x)
        return x


class MidBlock(nn.Module):
    def __init__(self, c, temb, heads, ctx_dim, groups, eps):
        super().__init__()
        self.attentions = nn.ModuleList([Transformer2DModel(c, heads, ctx_dim, groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, groups, eps), ResnetBlock2D(c, c, temb, groups, eps)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class _Out:
    def __init__(self, sample):
        self.sample = sample


class _Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class UNet2DConditionModel(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        cfg = dict(SD2_INPAINT_UNET)
        cfg.update(kw)
        self.config = _Cfg(cfg)
        ch = cfg["block_out_channels"]
        heads = cfg["attention_head_dim"]  # head COUNTS in SD-2 configs
        g, eps, ctx, L = cfg["norm_num_groups"], cfg["norm_eps"], cfg["cross_attention_dim"], cfg["layers_per_block"]
        temb = ch[0] * 4
        self.conv_in = nn.Conv2d(cfg["in_channels"], ch[0], 3, padding=1)
        self.time_embedding = nn.Module()
        self.time_embedding.linear_1 = nn.Linear(ch[0], temb)
        self.time_embedding.linear_2 = nn.Linear(temb, temb)
        self.down_blocks = nn.ModuleList()
        out = ch[0]
        for i, t in enumerate(cfg["down_block_types"]):
            cin, out = out, ch[i]
            self.down_blocks.append(
                DownBlock(cin, out, L, temb, heads[i], ctx, g, eps, i < len(ch) - 1, t.startswith("CrossAttn")))
        self.mid_block = MidBlock(ch[-1], temb, heads[-1], ctx, g, eps)
        self.up_blocks = nn.ModuleList()
        rch, rheads = list(reversed(ch)), list(reversed(heads))
        out = rch[0]
        for i, t in enumerate(cfg["up_block_types"]):
            prev, out = out, rch[i]
            cin = rch[min(i + 1, len(ch) - 1)]
            self.up_blocks.append(
                UpBlock(cin, out, prev, L + 1, temb, rheads[i], ctx, g, eps, i < len(ch) - 1, t.startswith("CrossAttn")))
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=eps)
        self.conv_out = nn.Conv2d(ch[0], cfg["out_channels"], 3, padding=1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    def time_emb(self, t, batch):
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.long)
        if t.ndim == 0:
            t = t[None]
        t = t.expand(batch)
        e = timestep_embedding(t, self.config.block_out_channels[0]).to(self.dtype)
        return self.time_embedding.linear_2(F.silu(self.time_embedding.linear_1(e)))

    def forward(self, sample, timestep, encoder_hidden_states, return_dict=True):
        emb = self.time_emb(timestep, sample.shape[0])
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips += outs
        x = self.mid_block(x, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, encoder_hidden_states)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return _Out(x) if return_dict else (x,)
