"""B200-native UNet2DConditionModel (SD-2-inpainting layout, in_channels=31) assembled from the sm_100a kernels in ops.py.

Drop-in for the object the reference builds in /root/reference/hubconf.py:30-37 and calls at
/root/reference/src/vto_pipelines/tryon_pipe.py:732: same constructor config keys, same state-dict key names (SURVEY.md
Appendix A.7), `unet(x_nchw, t, encoder_hidden_states=ctx).sample`.  Internally everything is NHWC bf16:

  * 3x3 convs: implicit GEMM on tcgen05 (ops.conv2d); the up-block skip concat and the ResnetBlock2D 1x1 conv_shortcut are
    extra K segments of the same accumulation (never materialised);
  * time embedding: the whole MLP + all 22 `time_emb_proj(silu(emb))` vectors depend only on the step index, so
    `plan_steps` tabulates `conv1.bias + time_emb_proj(...)` once per call and conv1's epilogue indexes it with the
    device-side step counter (graph-replay friendly);
  * cross-attention K/V of the text context are step-invariant: one GEMM per call for all 16 layers (`plan_context`);
  * attention: fused flash kernel reading per-head slices of the fused QKV projection in place;
  * LayerNorm (norm1/2/3 of every BasicTransformerBlock) never runs as a pass: the GEMM that produces the tensor also emits per-row
    partial sums, the GEMM that consumes it multiplies the raw tensor with W diag(gamma) and applies mean/rstd as a rank-1 correction
    in its epilogue (weights.fold_layernorm; env LADI_LN_FOLD=0 packs the stand-alone form for A/B timing);
  * Upsample2D: nearest-2x + conv3x3 as ONE sub-pixel convolution over the half-resolution tensor (4/9 of the MACs, no intermediate;
    env LADI_UP2X=0 for the materialised form).
"""
import math
import os

import torch

from . import engine as eng, lib, ops
from .weights import f32, fold_layernorm, interleave_geglu, pack_conv, pack_conv_up2x, pack_linear

LN_EPS = 1e-5  # nn.LayerNorm default (BasicTransformerBlock.norm1/2/3)

SD2_INPAINT_UNET = dict(
    in_channels=31, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024, norm_num_groups=32, norm_eps=1e-5,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    sample_size=64,
)


class _Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class _Sample:
    def __init__(self, sample):
        self.sample = sample


def unet_param_shapes(cfg):
    """Ordered {state-dict key: shape} of the architecture (the weight-format contract, Appendix A.7)."""
    cfg = {**SD2_INPAINT_UNET, **cfg}
    ch, heads, L, ctx = cfg["block_out_channels"], cfg["attention_head_dim"], cfg["layers_per_block"], cfg["cross_attention_dim"]
    temb = ch[0] * 4
    S = {}

    def conv(p, ci, co, k):
        S[p + ".weight"], S[p + ".bias"] = (co, ci, k, k), (co,)

    def lin(p, ci, co, bias=True):
        S[p + ".weight"] = (co, ci)
        if bias:
            S[p + ".bias"] = (co,)

    def norm(p, c):
        S[p + ".weight"], S[p + ".bias"] = (c,), (c,)

    def resnet(p, ci, co):
        norm(p + ".norm1", ci); conv(p + ".conv1", ci, co, 3); lin(p + ".time_emb_proj", temb, co)
        norm(p + ".norm2", co); conv(p + ".conv2", co, co, 3)
        if ci != co:
            conv(p + ".conv_shortcut", ci, co, 1)

    def transformer(p, c):
        norm(p + ".norm", c); lin(p + ".proj_in", c, c)
        b = p + ".transformer_blocks.0"
        norm(b + ".norm1", c)
        for a, kd in ((".attn1", c), (".attn2", ctx)):
            if a == ".attn2":
                norm(b + ".norm2", c)
            lin(b + a + ".to_q", c, c, False); lin(b + a + ".to_k", kd, c, False); lin(b + a + ".to_v", kd, c, False)
            lin(b + a + ".to_out.0", c, c)
        norm(b + ".norm3", c); lin(b + ".ff.net.0.proj", c, 8 * c); lin(b + ".ff.net.2", 4 * c, c)
        lin(p + ".proj_out", c, c)

    conv("conv_in", cfg["in_channels"], ch[0], 3)
    lin("time_embedding.linear_1", ch[0], temb); lin("time_embedding.linear_2", temb, temb)
    out = ch[0]
    for i, t in enumerate(cfg["down_block_types"]):
        cin, out = out, ch[i]
        for l in range(L):
            resnet(f"down_blocks.{i}.resnets.{l}", cin if l == 0 else out, out)
        if t.startswith("CrossAttn"):
            for l in range(L):
                transformer(f"down_blocks.{i}.attentions.{l}", out)
        if i < len(ch) - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", out, out, 3)
    transformer("mid_block.attentions.0", ch[-1])
    resnet("mid_block.resnets.0", ch[-1], ch[-1]); resnet("mid_block.resnets.1", ch[-1], ch[-1])
    rch = list(reversed(ch))
    out = rch[0]
    for i, t in enumerate(cfg["up_block_types"]):
        prev, out = out, rch[i]
        cin = rch[min(i + 1, len(ch) - 1)]
        for l in range(L + 1):
            resnet(f"up_blocks.{i}.resnets.{l}", (prev if l == 0 else out) + (cin if l == L else out), out)
        if t.startswith("CrossAttn"):
            for l in range(L + 1):
                transformer(f"up_blocks.{i}.attentions.{l}", out)
        if i < len(ch) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", out, out, 3)
    norm("conv_norm_out", ch[0]); conv("conv_out", ch[0], cfg["out_channels"], 3)
    return S


class UNet2DConditionModel:
    def __init__(self, **config):
        self.config = _Cfg({**SD2_INPAINT_UNET, **config})
        self.dtype = torch.bfloat16
        self.device = torch.device("cpu")
        self._sd = None
        self.P = None
        self._steps = None
        self._steps_key = None
        self._ctx = None
        self.pack_gen = 0  # bumped by every re-pack: captured graphs hold the packed weights' addresses (pipeline._graph_key)

    # ---- reference-compatible surface -------------------------------------------------------------------------------
    def eval(self):
        return self

    def requires_grad_(self, flag=False):
        return self

    def enable_xformers_memory_efficient_attention(self):  # inference.py:143-147: accepted, attention is already fused
        return self

    def load_state_dict(self, sd, strict=True):
        shapes = unet_param_shapes(self.config)
        missing = [k for k in shapes if k not in sd]
        unexpected = [k for k in sd if k not in shapes]
        bad = [k for k in shapes if k in sd and tuple(sd[k].shape) != tuple(shapes[k])]
        if strict and (missing or unexpected or bad):
            raise RuntimeError(f"UNet state_dict mismatch: missing={missing[:4]} unexpected={unexpected[:4]} shape={bad[:4]}")
        self._sd = {k: v.detach() for k, v in sd.items()}
        self.P = None
        if self.device.type == "cuda":
            self._pack()
        return self

    def to(self, device=None, dtype=None, **kw):
        if isinstance(device, torch.dtype):
            device, dtype = None, device
        if device is not None:
            device = torch.device(device)
            if device.type != "cuda":
                raise RuntimeError("ladi_vton_b200 UNet runs on CUDA (sm_100a) only; there is no CPU path")
            if self.P is not None and device == self.device:
                return self  # already packed on this device: re-packing would free weights a captured graph still points at
            self.device = device
            if self._sd is not None:
                self._pack()
            elif self.P is not None:
                raise RuntimeError("the state dict was released after packing; call load_state_dict again to move the UNet")
        return self

    def __call__(self, sample, timestep, encoder_hidden_states, return_dict=True):
        """Generic entry: NCHW sample (any float dtype), scalar timestep, ctx [B,77,ctx_dim]."""
        B, cin, h, w = sample.shape
        t = int(timestep)
        self.plan_steps([t])
        self.plan_context(encoder_hidden_states)
        x = torch.zeros((B, h, w, self.in_pitch), dtype=torch.bfloat16, device=self.device)
        ops.nchw_to_nhwc(sample.float().contiguous(), x)
        step = torch.zeros(2, dtype=torch.int32, device=self.device)
        eps = self.forward_nhwc(x, step)
        out = ops.nhwc_to_nchw(eps, self.config.out_channels).to(sample.dtype)
        return _Sample(out) if return_dict else (out,)

    # ---- packing ------------------------------------------------------------------------------------------------------
    def _pack(self):
        sd, dev, cfg = self._sd, self.device, self.config
        g = lambda k: sd[k].to(dev, torch.float32)
        P = {}
        ch, L = cfg.block_out_channels, cfg.layers_per_block
        self.in_pitch = (cfg.in_channels + 7) // 8 * 8
        self.resnets, self.transformers = [], []
        self.fold_ln = os.environ.get("LADI_LN_FOLD", "0") == "1"
        self.fuse_up = os.environ.get("LADI_UP2X", "1") != "0"

        def conv(p, srcs):
            P[p + ".w"] = pack_conv(g(p + ".weight"), srcs)
            P[p + ".b"] = f32(g(p + ".bias"))

        def resnet(p, srcs, co):
            ci = sum(srcs)
            P[p + ".n1"] = (f32(g(p + ".norm1.weight")), f32(g(p + ".norm1.bias")))
            P[p + ".n2"] = (f32(g(p + ".norm2.weight")), f32(g(p + ".norm2.bias")))
            P[p + ".w1"] = pack_conv(g(p + ".conv1.weight"), [ci])
            if ci != co:
                P[p + ".w2"] = pack_conv(g(p + ".conv2.weight"), [co], g(p + ".conv_shortcut.weight"), srcs)
                P[p + ".b2"] = f32(g(p + ".conv2.bias") + g(p + ".conv_shortcut.bias"))
            else:
                P[p + ".w2"] = pack_conv(g(p + ".conv2.weight"), [co])
                P[p + ".b2"] = f32(g(p + ".conv2.bias"))
            self.resnets.append((p, co))

        def transformer(p, c):
            b = p + ".transformer_blocks.0"
            P[p + ".norm"] = (f32(g(p + ".norm.weight")), f32(g(p + ".norm.bias")))
            for n in ("proj_in", "proj_out"):
                P[p + f".{n}.w"], P[p + f".{n}.b"] = pack_linear(g(p + f".{n}.weight")), f32(g(p + f".{n}.bias"))
            wqkv = torch.cat([g(b + ".attn1.to_q.weight"), g(b + ".attn1.to_k.weight"), g(b + ".attn1.to_v.weight")])
            wi, bi = interleave_geglu(g(b + ".ff.net.0.proj.weight"), g(b + ".ff.net.0.proj.bias"))
            if self.fold_ln and c % 64 == 0:  # LayerNorm folded into its consumer: (W diag(gamma), colsum, W beta + bias)
                ln = lambda i: (g(b + f".norm{i}.weight"), g(b + f".norm{i}.bias"))
                P[b + ".qkv"], P[b + ".qkv.cs"], P[b + ".qkv.b"] = fold_layernorm(wqkv, *ln(1))
                P[b + ".q2"], P[b + ".q2.cs"], P[b + ".q2.b"] = fold_layernorm(g(b + ".attn2.to_q.weight"), *ln(2))
                P[b + ".ff1.w"], P[b + ".ff1.cs"], P[b + ".ff1.b"] = fold_layernorm(wi, *ln(3), bi)
            else:
                for i in (1, 2, 3):
                    P[b + f".ln{i}"] = (f32(g(b + f".norm{i}.weight")), f32(g(b + f".norm{i}.bias")))
                P[b + ".qkv"], P[b + ".q2"] = pack_linear(wqkv), pack_linear(g(b + ".attn2.to_q.weight"))
                P[b + ".ff1.w"], P[b + ".ff1.b"] = pack_linear(wi), f32(bi)
            P[b + ".o1.w"], P[b + ".o1.b"] = pack_linear(g(b + ".attn1.to_out.0.weight")), f32(g(b + ".attn1.to_out.0.bias"))
            P[b + ".o2.w"], P[b + ".o2.b"] = pack_linear(g(b + ".attn2.to_out.0.weight")), f32(g(b + ".attn2.to_out.0.bias"))
            P[b + ".ff2.w"], P[b + ".ff2.b"] = pack_linear(g(b + ".ff.net.2.weight")), f32(g(b + ".ff.net.2.bias"))
            self.transformers.append((p, c))

        conv("conv_in", [cfg.in_channels])
        out = ch[0]
        for i, t in enumerate(cfg.down_block_types):
            cin, out = out, ch[i]
            for l in range(L):
                resnet(f"down_blocks.{i}.resnets.{l}", [cin if l == 0 else out], out)
                if t.startswith("CrossAttn"):
                    transformer(f"down_blocks.{i}.attentions.{l}", out)
            if i < len(ch) - 1:
                conv(f"down_blocks.{i}.downsamplers.0.conv", [out])
        resnet("mid_block.resnets.0", [ch[-1]], ch[-1])
        transformer("mid_block.attentions.0", ch[-1])
        resnet("mid_block.resnets.1", [ch[-1]], ch[-1])
        rch = list(reversed(ch))
        out = rch[0]
        for i, t in enumerate(cfg.up_block_types):
            prev, out = out, rch[i]
            cin = rch[min(i + 1, len(ch) - 1)]
            for l in range(L + 1):
                resnet(f"up_blocks.{i}.resnets.{l}", [prev if l == 0 else out, cin if l == L else out], out)
                if t.startswith("CrossAttn"):
                    transformer(f"up_blocks.{i}.attentions.{l}", out)
            if i < len(ch) - 1:
                p = f"up_blocks.{i}.upsamplers.0.conv"
                if self.fuse_up:
                    P[p + ".w"], P[p + ".b"] = pack_conv_up2x(g(p + ".weight"), [out]), f32(g(p + ".bias"))
                else:
                    conv(p, [out])
        P["conv_norm_out"] = (f32(g("conv_norm_out.weight")), f32(g("conv_norm_out.bias")))
        conv("conv_out", [ch[0]])
        # time-embedding MLP + one fused projection for all resnets: rows = [time_emb_proj_r ; ...], bias += conv1.bias
        # linear_1 reads the fp32 sinusoid table as a (hi, lo) pair of bf16 columns against [W | W]: input error 2^-17 instead of 2^-9
        w1 = g("time_embedding.linear_1.weight")
        P["te1.w"], P["te1.b"] = pack_linear(torch.cat([pack_linear(w1).float(), pack_linear(w1).float()], dim=1)), f32(g("time_embedding.linear_1.bias"))
        P["te2.w"], P["te2.b"] = pack_linear(g("time_embedding.linear_2.weight")), f32(g("time_embedding.linear_2.bias"))
        P["temb_all.w"] = pack_linear(torch.cat([g(p + ".time_emb_proj.weight") for p, _ in self.resnets]))
        P["temb_all.b"] = f32(torch.cat([g(p + ".time_emb_proj.bias") + g(p + ".conv1.bias") for p, _ in self.resnets]))
        self.temb_off, off = {}, 0
        for p, co in self.resnets:
            self.temb_off[p] = off
            off += co
        self.temb_total = off
        # cross-attention K/V projections of all layers as one [sum 2C, ctx] weight
        P["kv_all.w"] = pack_linear(torch.cat([torch.cat([g(p + ".transformer_blocks.0.attn2.to_k.weight"),
                                                          g(p + ".transformer_blocks.0.attn2.to_v.weight")]) for p, _ in self.transformers]))
        self.kv_off, off = {}, 0
        for p, c in self.transformers:
            self.kv_off[p] = off
            off += 2 * c
        self.kv_total = off
        self.P = P
        self.ws = ops.GroupNormWS(dev)
        self.pack_gen += 1
        self._steps = self._steps_key = self._ctx = None
        self.engine = None
        if eng.enabled() and not self.fold_ln and dev.type == "cuda":  # the C++ launch sequence (csrc/engine.cu); LADI_ENGINE=0 -> Python sequencing
            self.engine = eng.Engine(eng.flatten(P), **self.engine_config())
        if any(v.is_cuda for v in sd.values()):
            self._sd = None  # device-resident fp32 source weights (3.5 GB for the full UNet) are not kept beside the bf16 pack

    def engine_config(self):
        cfg = self.config
        return dict(unet_channels=cfg.block_out_channels, unet_heads=cfg.attention_head_dim, unet_layers_per_block=cfg.layers_per_block,
                    unet_down_attn=[int(t.startswith("CrossAttn")) for t in cfg.down_block_types],
                    unet_up_attn=[int(t.startswith("CrossAttn")) for t in cfg.up_block_types],
                    unet_in_channels=cfg.in_channels, unet_out_channels=cfg.out_channels, unet_norm_eps=cfg.norm_eps,
                    norm_groups=cfg.norm_num_groups, fuse_upsample=int(self.fuse_up))

    # ---- per-call planning (step-invariant work) ----------------------------------------------------
    def plan_steps(self, timesteps):
        """Tabulate conv1.bias + time_emb_proj(silu(time_embedding(t))) for every step: fp32 [steps, sum C_out]."""
        P, c0 = self.P, self.config.block_out_channels[0]
        key = tuple(int(x) for x in timesteps)
        if self._steps is not None and self._steps_key == key:
            return self._steps  # same schedule as the previous call: the table (and its address) stands
        keep = self._steps if (self._steps is not None and tuple(self._steps.shape) == (len(timesteps), self.temb_total)) else None
        if getattr(self, "engine", None) is not None and ops.PROFILE is None and lib.RECORD is None:
            # one ABI call (csrc/engine.cu unet_plan_steps): sinusoid table on the device, the time MLP and all 22 time_emb_proj as three GEMMs
            n = len(timesteps)
            tdev = torch.tensor([float(x) for x in timesteps], dtype=torch.float32).to(self.device)
            out = keep if keep is not None else torch.empty((n, self.temb_total), dtype=torch.float32, device=self.device)
            ws = self.engine.workspace(eng.MODULE_UNET_PLAN, n, 0, 0)
            lib.call("ladi_unet_plan_steps", self.engine.h, ops._ptr(tdev), n, ops._ptr(out), ops._ptr(ws), ws.numel(), ops._stream())
            self._steps, self._steps_key = out, key
            return self._steps
        half = c0 // 2
        t = torch.tensor([float(x) for x in timesteps], dtype=torch.float32)
        freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
        arg = t[:, None] * freq[None, :]
        emb = torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)  # fp32 input table (host-side)
        kp = P["te1.w"].shape[1] // 2
        hi = emb.to(torch.bfloat16)
        lo = (emb - hi.float()).to(torch.bfloat16)
        pair = torch.zeros((emb.shape[0], 2 * kp), dtype=torch.bfloat16)
        pair[:, :c0], pair[:, kp:kp + c0] = hi, lo
        e1 = ops.gemm(pair.to(self.device), P["te1.w"], P["te1.w"].shape[0], bias=P["te1.b"], act=ops.ACT_SILU)
        e2 = ops.gemm(e1, P["te2.w"], P["te2.w"].shape[0], bias=P["te2.b"], act=ops.ACT_SILU)  # = silu(emb)
        self._steps = ops.gemm(e2, P["temb_all.w"], self.temb_total, bias=P["temb_all.b"], out_fp32=True, out=keep)  # stable address
        self._steps_key = key
        return self._steps

    def plan_context(self, ctx, out=None):
        """K/V of the text context for all 16 cross-attention layers: bf16 [B', 77, sum 2C].  `out`: caller-owned buffer (the pipeline
        keeps one per session so captured graphs of different shapes never share or invalidate it)."""
        B, T, D = ctx.shape
        c = ctx.to(self.device, torch.bfloat16).contiguous().view(B * T, D)
        if out is None and self._ctx is not None and tuple(self._ctx.shape) == (B, T, self.kv_total):
            out = self._ctx
        if getattr(self, "engine", None) is not None and ops.PROFILE is None and lib.RECORD is None and D % 8 == 0:
            if out is None:
                out = torch.empty((B, T, self.kv_total), dtype=torch.bfloat16, device=self.device)
            assert out.is_contiguous()
            lib.call("ladi_unet_plan_context", self.engine.h, ops._ptr(c), B * T, D, ops._ptr(out), ops._stream())  # one GEMM, csrc/engine.cu
            self._ctx = out.view(B, T, self.kv_total)
            return self._ctx
        kv = ops.gemm(c, self.P["kv_all.w"], self.kv_total, out=None if out is None else out.view(B * T, self.kv_total))
        self._ctx = kv.view(B, T, self.kv_total)
        return self._ctx

    # ---- forward -------------------------------------------------------------------------------------------------------
    def _resnet(self, p, srcs, co, step):
        P, cfg = self.P, self.config
        g, eps = cfg.norm_num_groups, cfg.norm_eps
        hn = ops.groupnorm(srcs, *P[p + ".n1"], g, eps, self.ws, silu=True)
        off = self.temb_off[p]
        h = ops.conv2d([hn], P[p + ".w1"], co, bias=self._steps[:, off:off + co], bias_step_stride=self.temb_total, step_ptr=step)
        hn2 = ops.groupnorm([h], *P[p + ".n2"], g, eps, self.ws, silu=True)
        if sum(s.shape[3] for s in srcs) != co:
            return ops.conv2d([hn2], P[p + ".w2"], co, bias=P[p + ".b2"], shortcut=srcs)
        return ops.conv2d([hn2], P[p + ".w2"], co, bias=P[p + ".b2"], residual=srcs[0])

    def _transformer(self, p, x, heads):
        P, cfg = self.P, self.config
        B, h, w, C = x.shape
        M, N = B * h * w, h * w
        b = p + ".transformer_blocks.0"
        hn = ops.groupnorm([x], *P[p + ".norm"], cfg.norm_num_groups, 1e-6, self.ws, silu=False)
        off = self.kv_off[p]
        kc, vc = self._ctx[..., off:off + C], self._ctx[..., off + C:off + 2 * C]
        scale = (C // heads) ** -0.5
        if (b + ".qkv.cs") in P:
            # LayerNorm folded: each producer GEMM emits the row statistics its consumer's epilogue needs (no LayerNorm launches)
            stat = lambda: torch.empty((M, C // 32, 2), dtype=torch.float32, device=x.device)
            s1, s2, s3 = stat(), stat(), stat()
            t = ops.gemm(hn.view(M, C), P[p + ".proj_in.w"], C, bias=P[p + ".proj_in.b"], rowstat=s1)
            qkv = ops.gemm(t, P[b + ".qkv"], 3 * C, bias=P[b + ".qkv.b"], ln=(s1, P[b + ".qkv.cs"], LN_EPS)).view(B, N, 3 * C)
            a = ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads, scale)
            t = ops.gemm(a.view(M, C), P[b + ".o1.w"], C, bias=P[b + ".o1.b"], residual=t, rowstat=s2)
            q = ops.gemm(t, P[b + ".q2"], C, bias=P[b + ".q2.b"], ln=(s2, P[b + ".q2.cs"], LN_EPS)).view(B, N, C)
            a = ops.attention(q, kc, vc, heads, scale)
            t = ops.gemm(a.view(M, C), P[b + ".o2.w"], C, bias=P[b + ".o2.b"], residual=t, rowstat=s3)
            ff = ops.gemm(t, P[b + ".ff1.w"], 8 * C, bias=P[b + ".ff1.b"], act=ops.ACT_GEGLU, ln=(s3, P[b + ".ff1.cs"], LN_EPS))
        else:
            t = ops.gemm(hn.view(M, C), P[p + ".proj_in.w"], C, bias=P[p + ".proj_in.b"])
            qkv = ops.gemm(ops.layernorm(t, *P[b + ".ln1"]), P[b + ".qkv"], 3 * C).view(B, N, 3 * C)
            a = ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads, scale)
            t = ops.gemm(a.view(M, C), P[b + ".o1.w"], C, bias=P[b + ".o1.b"], residual=t)
            q = ops.gemm(ops.layernorm(t, *P[b + ".ln2"]), P[b + ".q2"], C).view(B, N, C)
            a = ops.attention(q, kc, vc, heads, scale)
            t = ops.gemm(a.view(M, C), P[b + ".o2.w"], C, bias=P[b + ".o2.b"], residual=t)
            ff = ops.gemm(ops.layernorm(t, *P[b + ".ln3"]), P[b + ".ff1.w"], 8 * C, bias=P[b + ".ff1.b"], act=ops.ACT_GEGLU)
        t = ops.gemm(ff, P[b + ".ff2.w"], C, bias=P[b + ".ff2.b"], residual=t)
        return ops.gemm(t, P[p + ".proj_out.w"], C, bias=P[p + ".proj_out.b"], residual=x.view(M, C)).view(B, h, w, C)

    def forward_nhwc(self, x_in, step):
        """x_in NHWC bf16 [B', h, w, in_pitch] (first in_channels valid), `step` = device int32[2] {index into the planned
        step table, 0} -> eps NHWC fp32 [B', h, w, 4]."""
        P, cfg = self.P, self.config
        ch, L, heads = cfg.block_out_channels, cfg.layers_per_block, cfg.attention_head_dim
        assert self._steps is not None and self._ctx is not None, "call plan_steps/plan_context first"
        if getattr(self, "engine", None) is not None and ops.PROFILE is None and lib.RECORD is None:
            # the product path: ONE ABI call, the launch sequence below lives in C++ (csrc/engine.cu unet_forward)
            B, h, w, _ = x_in.shape
            assert x_in.is_contiguous() and x_in.shape[3] == self.in_pitch and self._ctx.is_contiguous() and self._steps.is_contiguous()
            eps = torch.empty((B, h, w, 4), dtype=torch.float32, device=self.device)
            ws = self.engine.workspace(eng.MODULE_UNET, B, h, w)
            lib.call("ladi_unet_forward", self.engine.h, ops._ptr(x_in), ops._ptr(step), ops._ptr(self._steps), ops._ptr(self._ctx), B, h, w,
                     self._ctx.shape[1], ops._ptr(eps), ops._ptr(ws), ws.numel(), ops._stream())
            return eps
        x = ops.conv2d([x_in[..., :cfg.in_channels]], P["conv_in.w"], ch[0], bias=P["conv_in.b"])
        skips = [x]
        out = ch[0]
        for i, t in enumerate(cfg.down_block_types):
            out = ch[i]
            for l in range(L):
                x = self._resnet(f"down_blocks.{i}.resnets.{l}", [x], out, step)
                if t.startswith("CrossAttn"):
                    x = self._transformer(f"down_blocks.{i}.attentions.{l}", x, heads[i])
                skips.append(x)
            if i < len(ch) - 1:
                p = f"down_blocks.{i}.downsamplers.0.conv"
                x = ops.conv2d([x], P[p + ".w"], out, bias=P[p + ".b"], stride=2, pad_lo=1)
                skips.append(x)
        x = self._resnet("mid_block.resnets.0", [x], ch[-1], step)
        x = self._transformer("mid_block.attentions.0", x, heads[-1])
        x = self._resnet("mid_block.resnets.1", [x], ch[-1], step)
        rch, rheads = list(reversed(ch)), list(reversed(heads))
        for i, t in enumerate(cfg.up_block_types):
            out = rch[i]
            for l in range(L + 1):
                x = self._resnet(f"up_blocks.{i}.resnets.{l}", [x, skips.pop()], out, step)  # concat [current, skip] stays virtual
                if t.startswith("CrossAttn"):
                    x = self._transformer(f"up_blocks.{i}.attentions.{l}", x, rheads[i])
            if i < len(ch) - 1:
                p = f"up_blocks.{i}.upsamplers.0.conv"
                x = (ops.conv2d([x], P[p + ".w"], out, bias=P[p + ".b"], up2x=True) if self.fuse_up
                     else ops.conv2d([ops.upsample2x(x)], P[p + ".w"], out, bias=P[p + ".b"]))
        hn = ops.groupnorm([x], *P["conv_norm_out"], cfg.norm_num_groups, cfg.norm_eps, self.ws, silu=True)
        B, h, w, _ = hn.shape
        eps = torch.empty((B, h, w, 4), dtype=torch.float32, device=self.device)
        return ops.conv2d([hn], P["conv_out.w"], cfg.out_channels, bias=P["conv_out.b"], out=eps, out_fp32=True)
