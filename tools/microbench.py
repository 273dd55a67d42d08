"""Isolated device timing of the small kernels at the bench shapes (100 back-to-back launches between CUDA events, so launch
gaps of the eager profiler are excluded).  Prints us per launch and achieved GB/s or TFLOP/s."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ladi_vton_b200 import ops, weights  # noqa: E402

dev = torch.device("cuda:0")
N = 50


def timeit(fn, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(N):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / N * 1e3  # us


def rnd(*shape):
    return torch.randn(shape, device=dev).bfloat16()


print("== GroupNorm (stats / apply), bytes = activation size")
ws = ops.GroupNormWS(dev)
for n, hw, c0, c1 in [(16, 3072, 320, 0), (16, 3072, 640, 320), (16, 768, 640, 0), (16, 768, 1280, 640), (16, 192, 1280, 0), (16, 48, 1280, 1280),
                      (8, 196608, 128, 0), (8, 49152, 256, 0)]:
    h, w = hw // 48, 48
    x0 = rnd(n, h, w, c0)
    x1 = rnd(n, h, w, c1) if c1 else None
    C = c0 + c1
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    out = torch.empty((n, h, w, C), dtype=torch.bfloat16, device=dev)
    srcs = [x0] + ([x1] if c1 else [])
    wsb = ws.get(n, hw, 32)
    from ladi_vton_b200 import lib
    P = ops._ptr
    s = ops._stream()
    t_s = timeit(lambda: lib.call("ladi_groupnorm_stats", P(x0), c0, c0, P(x1), c1, c1, n, hw, 32, P(wsb), s))
    t_a = timeit(lambda: lib.call("ladi_groupnorm_apply", P(x0), c0, c0, P(x1), c1, c1, n, hw, 32, P(wsb), P(g), P(b), 1e-5, 1, P(None), 0, P(out), C, s))
    mb = n * hw * C * 2 / 1e6
    print(f"  n={n} hw={hw} C={C:5d}: stats {t_s:7.1f} us ({mb / t_s * 1e3 / 1e3:7.0f} GB/s)   apply {t_a:7.1f} us ({2 * mb / t_a:7.0f} GB/s)   [{mb:.1f} MB]")

print("== LayerNorm")
for rows, C in [(49152, 320), (12288, 640), (3072, 1280), (768, 1280)]:
    x = rnd(rows, C)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    out = torch.empty_like(x)
    t = timeit(lambda: ops.layernorm(x, g, b, out=out))
    mb = rows * C * 2 / 1e6
    print(f"  rows={rows} C={C}: {t:7.1f} us ({2 * mb / t:7.0f} GB/s)")

print("== GEMM / conv (TFLOP/s)")
for M, K, Nn, kw in [(49152, 320, 320, {}), (49152, 320, 320, {"res": 1}), (49152, 320, 960, {}), (49152, 320, 2560, {"geglu": 1}), (49152, 1280, 320, {"res": 1}),
                     (12288, 640, 640, {"res": 1}), (3072, 1280, 1280, {"res": 1}), (768, 1280, 1280, {"res": 1}), (1232, 1024, 24960, {})]:
    a = rnd(M, K)
    wt = weights.pack_linear(torch.randn(Nn, K, device=dev) * K ** -0.5)
    bias = torch.zeros(Nn, device=dev)
    res = rnd(M, Nn) if kw.get("res") else None
    out = torch.empty((M, Nn // 2 if kw.get("geglu") else Nn), dtype=torch.bfloat16, device=dev)
    t = timeit(lambda: ops.gemm(a, wt, Nn, bias=bias, residual=res, act=ops.ACT_GEGLU if kw.get("geglu") else 0, out=out))
    print(f"  gemm M={M} K={K} N={Nn} {kw}: {t:7.1f} us  {2 * M * K * Nn / t / 1e6:7.1f} TFLOP/s")
for n, h, w, cin, cout in [(16, 64, 48, 320, 320), (16, 32, 24, 640, 640), (16, 16, 12, 1280, 1280), (16, 8, 6, 1280, 1280), (16, 8, 6, 2560, 1280)]:
    x = rnd(n, h, w, cin)
    wt = weights.pack_conv(torch.randn(cout, cin, 3, 3, device=dev) * (9 * cin) ** -0.5, [cin])
    bias = torch.zeros(cout, device=dev)
    out = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=dev)
    t = timeit(lambda: ops.conv2d([x], wt, cout, bias=bias, out=out))
    print(f"  conv3x3 n={n} {h}x{w} {cin}->{cout}: {t:7.1f} us  {2 * n * h * w * cout * 9 * cin / t / 1e6:7.1f} TFLOP/s")

print("== attention")
for B, heads, nq, nkv in [(16, 5, 3072, 3072), (16, 10, 768, 768), (16, 20, 192, 192), (16, 20, 48, 48), (16, 5, 3072, 77), (16, 10, 768, 77), (16, 20, 192, 77)]:
    C = heads * 64
    q = rnd(B, nq, C); k = rnd(B, nkv, C); v = rnd(B, nkv, C)
    out = torch.empty_like(q)
    for var in ((0, 1, 4, 5) if nkv > 128 else (0,)):
        t = timeit(lambda: ops.attention(q, k, v, heads, 0.125, out=out, variant=var))
        print(f"  attn B={B} heads={heads} nq={nq} nkv={nkv} variant={var}: {t:7.1f} us  {4 * B * heads * nq * nkv * 64 / t / 1e6:7.1f} TFLOP/s")

print("== empty-ish launch floor")
a = rnd(8, 64); wt = weights.pack_linear(torch.randn(8, 64, device=dev)); out = torch.empty((8, 8), dtype=torch.bfloat16, device=dev)
print(f"  tiny gemm: {timeit(lambda: ops.gemm(a, wt, 8, out=out)):.1f} us;  tiny add: {timeit(lambda: ops.add(a, a, out=torch.empty_like(a))):.1f} us")
