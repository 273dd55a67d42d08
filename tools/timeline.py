"""Kernel timeline of CUDA-graph-replayed denoising steps via torch.profiler (CUPTI): true per-kernel durations and the idle
gaps between consecutive kernels inside the graph.  Writes gpurun_out/timeline.txt."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from ladi_vton_b200 import synthetic as S  # noqa: E402

B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 512, 384
dev = torch.device("cuda:0")
pipe, _ = S.build_pipeline(dev, weights_on_device=True)
inp = {k: v.to(dev) for k, v in S.synthetic_inputs(B, H, W).items()}
kw = dict(height=H, width=W, num_inference_steps=6, guidance_scale=7.5, output_type="pt")
pipe(**inp, **kw)
pipe(**inp, **kw)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pipe(**inp, **kw)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
names = [e.name for e in evs]
# find ddim kernels -> step boundaries
idx = [i for i, n in enumerate(names) if "ddim_cfg_kernel" in n]
out = open(os.path.join(ROOT, "gpurun_out", "timeline.txt"), "w")
out.write(f"{len(evs)} kernel events; ddim kernels at {idx}\n")
if len(idx) >= 4:
    a, b = idx[2] + 1, idx[3] + 1  # one fully replayed step
    step = evs[a:b]
    t0, t1 = step[0].time_range.start, step[-1].time_range.end
    busy = sum(e.time_range.end - e.time_range.start for e in step)
    out.write(f"one replayed step: {len(step)} kernels, wall {(t1 - t0) / 1e3:.3f} ms, busy {busy / 1e3:.3f} ms, idle gaps {(t1 - t0 - busy) / 1e3:.3f} ms\n")
    agg = collections.OrderedDict()
    gaps = collections.defaultdict(float)
    prev_end = None
    for e in step:
        import re
        n = e.name.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        n = re.split(r"\(", n)[0]
        # exclusive time: with programmatic dependent launch a kernel's CTAs may start (and wait) while its predecessor drains;
        # count only the part after the predecessor ended
        st = e.time_range.start if prev_end is None else max(e.time_range.start, prev_end)
        d = max(0, e.time_range.end - st)
        a_ = agg.setdefault(n, [0, 0.0])
        a_[0] += 1; a_[1] += d
        if prev_end is not None:
            gaps[n] += max(0, e.time_range.start - prev_end)
        prev_end = max(e.time_range.end, prev_end or 0)
    out.write(f"{'count':>5} {'excl ms':>9} {'avg us':>8} {'gap-before ms':>13}  kernel\n")
    for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"{c:5d} {d / 1e3:9.3f} {d / c:8.1f} {gaps[n] / 1e3:13.3f}  {n[:90]}\n")
    out.write("\nfirst 60 kernels of the step (start us rel, dur us, gap us, name):\n")
    prev_end = None
    for e in step[:60]:
        g = 0 if prev_end is None else e.time_range.start - prev_end
        out.write(f"{(e.time_range.start - t0):9.1f} {(e.time_range.end - e.time_range.start):8.1f} {g:7.1f}  {e.name[:80]}\n")
        prev_end = e.time_range.end
out.close()
print(open(os.path.join(ROOT, "gpurun_out", "timeline.txt")).read()[:7000])
