"""Per-op device-time table of one eager UNet forward (and VAE/EMASC) at the bench shape: every ABI call bracketed by CUDA
events (ops.PROFILE), aggregated by (entry point, shape tag).  Writes gpurun_out/op_profile.txt."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ladi_vton_b200 import ops, synthetic as S  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (512, 384)
OUT = sys.argv[4] if len(sys.argv) > 4 else "op_profile.txt"
dev = torch.device("cuda:0")
pipe, _ = S.build_pipeline(dev)
inp = {k: v.to(dev) for k, v in S.synthetic_inputs(B, H, W).items()}
kw = dict(height=H, width=W, num_inference_steps=3, guidance_scale=7.5, output_type="pt")
pipe.use_cuda_graph = False
pipe(**inp, **kw)  # warm-up (packs nothing new, sets func attributes)
torch.cuda.synchronize()


def table(title, prof, f):
    agg = collections.OrderedDict()
    for name, e0, e1, fl, tag in prof:
        k = (name, tag)
        t = e0.elapsed_time(e1)
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] += fl
    tot = sum(a[1] for a in agg.values())
    f.write(f"\n== {title}: {len(prof)} launches, {tot:.3f} ms (event-bracketed, eager)\n")
    f.write(f"{'count':>5} {'ms':>9} {'%':>6} {'TFLOP/s':>9}  op\n")
    for (name, tag), (c, t, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{c:5d} {t:9.3f} {100 * t / tot:6.2f} {fl / (t * 1e-3) / 1e12 if t > 0 else 0:9.1f}  {name.replace('ladi_', '')} {tag}\n")
    fam = collections.defaultdict(float)
    for (name, tag), (c, t, fl) in agg.items():
        fam[name] += t
    for name, t in sorted(fam.items(), key=lambda kv: -kv[1]):
        f.write(f"   family {name:32s} {t:9.3f} ms {100 * t / tot:6.2f}%\n")


os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", OUT), "w") as f:
    s = next(iter(pipe._sessions.values()))
    pipe.unet._ctx = s.ctx_kv
    ops.PROFILE = []
    pipe.unet.forward_nhwc(s.unet_in, s.step)
    torch.cuda.synchronize()
    table(f"UNet forward, UNet batch {2 * B}, latents {H // 8}x{W // 8}", ops.PROFILE, f)
    ops.PROFILE = []
    mom, feats = pipe.vae.encode_nhwc(inp["image"])
    torch.cuda.synchronize()
    table(f"VAE encode, batch {B}, {H}x{W}", ops.PROFILE, f)
    ops.PROFILE = []
    sel = [feats[i] for i in range(1, 6)]
    inter = pipe.emasc(sel, [ops.inv_mask_rows(inp["mask_image"], H // t.shape[1]) for t in sel])
    torch.cuda.synchronize()
    table(f"EMASC, batch {B}", ops.PROFILE, f)
    ops.PROFILE = []
    pipe.vae.decode_nhwc(s.latents, inter, [1, 2, 3, 4, 5], scale=1 / 0.18215)
    torch.cuda.synchronize()
    table(f"VAE decode, batch {B}", ops.PROFILE, f)
    ops.PROFILE = None
print(open(os.path.join(ROOT, "gpurun_out", OUT)).read()[:9000])
