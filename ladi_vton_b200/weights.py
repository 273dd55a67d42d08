"""Weight packing: reference state-dict tensors (diffusers 0.14 naming) -> the K-major bf16 layouts the
sm_100a kernels consume.  One-time, at model load; uses torch only for reshapes/casts.

Packed conv weight [c_out, K]: K order = for tap (ky, kx) row-major: for each source: channels zero-padded to a multiple of
64; then the optional fused 1x1 shortcut sources, each padded to 64 (matches ladi_conv2d_bf16's K-segment walk).
"""
import torch

BK = 64


def _pad_cols(w2d, to):
    if w2d.shape[1] == to:
        return w2d
    out = w2d.new_zeros((w2d.shape[0], to))
    out[:, : w2d.shape[1]] = w2d
    return out


def pad64(c):
    return (c + BK - 1) // BK * BK


def pack_conv(w, src_channels, sc_w=None, sc_channels=()):
    """w [c_out, sum(src_channels), kh, kw]; sc_w [c_out, sum(sc_channels), 1, 1] or None -> bf16 [c_out, K]."""
    co, ci, kh, kw = w.shape
    assert ci == sum(src_channels)
    cols = []
    for ky in range(kh):
        for kx in range(kw):
            off = 0
            for c in src_channels:
                cols.append(_pad_cols(w[:, off:off + c, ky, kx], pad64(c)))
                off += c
    if sc_w is not None:
        off = 0
        for c in sc_channels:
            cols.append(_pad_cols(sc_w[:, off:off + c, 0, 0], pad64(c)))
            off += c
    return torch.cat(cols, dim=1).to(torch.bfloat16).contiguous()


def pack_linear(w):
    """w [n_out, k] -> bf16 [n_out, pad64(k)]."""
    return _pad_cols(w, pad64(w.shape[1])).to(torch.bfloat16).contiguous()


def interleave_geglu(w, b):
    """GEGLU proj: rows [0:inner] = value, [inner:2*inner] = gate -> rows (v0, g0, v1, g1, ...) so that value/gate of one
    output land in adjacent accumulator columns (epilogue computes v * gelu(g))."""
    inner = w.shape[0] // 2
    wi = torch.stack([w[:inner], w[inner:]], dim=1).reshape(2 * inner, w.shape[1])
    bi = torch.stack([b[:inner], b[inner:]], dim=1).reshape(2 * inner)
    return wi, bi


def f32(t):
    return t.detach().to(torch.float32).contiguous()


def merge_up2x(w):
    """Sub-pixel decomposition of `conv3x3(nearest_upsample_2x(x))` (diffusers Upsample2D): output pixel (2i+py, 2j+px) reads the 2x2
    input pixels (i+py-1 .. i+py, j+px-1 .. j+px); the 3x3 taps that land on the same input pixel are summed (fp32).
    w [co, ci, 3, 3] -> [4 (parity py*2+px), co, ci, 2, 2]."""
    rows = {0: ([0], [1, 2]), 1: ([0, 1], [2])}  # parity -> kernel rows feeding tap 0 / tap 1
    out = w.new_zeros((4,) + tuple(w.shape[:2]) + (2, 2))
    for py in (0, 1):
        for px in (0, 1):
            for ty in (0, 1):
                for tx in (0, 1):
                    acc = 0
                    for ky in rows[py][ty]:
                        for kx in rows[px][tx]:
                            acc = acc + w[:, :, ky, kx]
                    out[py * 2 + px, :, :, ty, tx] = acc
    return out


def pack_conv_up2x(w, src_channels):
    """-> bf16 [4 * c_out, 4 taps * sum(pad64(c))]: the four parities' merged 2x2 weights stacked along the rows (the layout
    ladi_conv2d_bf16 reads with up2x = 1)."""
    m = merge_up2x(w.float())
    return torch.cat([pack_conv(m[p], src_channels) for p in range(4)], dim=0).contiguous()


def fold_layernorm(w, gamma, beta, bias=None):
    """LayerNorm folded into the linear that consumes it:  LN(x) W^T + b = rstd * (x W'^T - mean * colsum(W')) + (W beta + b)  with
    W' = W diag(gamma).  Returns (packed bf16 W', colsum fp32 [n_out] of the bf16-ROUNDED W' -- what the tensor core multiplies --,
    bias' fp32 [n_out])."""
    w, gamma, beta = w.float(), gamma.float(), beta.float()
    wp = pack_linear(w * gamma[None, :])
    colsum = wp.float().sum(dim=1).contiguous()
    b = w @ beta
    if bias is not None:
        b = b + bias.float()
    return wp, colsum, b.contiguous()
