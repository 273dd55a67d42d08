"""Times the cloth-warping front-end on cuda:0: generate_warped_cloth at 512x384 (TPS at 256x192 +
grid_sample + refinement U-Net), hub-constructor sizes, seeded random weights, CUDA events after warm-up; beside it the fp32 oracle on
the host cores (one call, batch 1, scaled).  Prints one JSON line.

    python tools/warp_bench.py [--batch 8] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    from ladi_vton_b200 import generate_warped_cloth, lib, synthetic as S
    from ladi_vton_b200.warp import ConvNet_TPS, UNetVanilla, control_points
    dev = torch.device("cuda:0")
    tps_sd = S.warp_state_dict(ConvNet_TPS(256, 192, 21, 3).param_shapes(), 11, ctrl_bias=torch.atanh(control_points()).view(-1))
    unet_sd = S.warp_state_dict(UNetVanilla(24, 3, True).param_shapes(), 12)
    tps = ConvNet_TPS(256, 192, 21, 3).load_state_dict(tps_sd).to(dev)
    net = UNetVanilla(24, 3, True).load_state_dict(unet_sd).to(dev)
    inp = S.warp_inputs(a.batch, 512, 384, seed=7)
    dinp = {k: v.to(dev) for k, v in inp.items()}
    fn = lambda: generate_warped_cloth(tps, net, dinp["cloth"], dinp["im_mask"], dinp["pose_map"])
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    n0 = lib.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    line = dict(stage="generate_warped_cloth 512x384", batch=a.batch, ms=round(ms, 3), images_per_s=round(a.batch / ms * 1e3, 1),
                gpu_launches=(lib.launches - n0) // 10)
    if not a.no_cpu:
        from ladi_oracle import warp as ow
        otps = ow.ConvNet_TPS(256, 192, 21, 3).eval()
        otps.load_state_dict(tps_sd, strict=False)
        onet = ow.UNetVanilla(24, 3, True).eval()
        onet.load_state_dict(unet_sd)
        torch.set_num_threads(min(64, os.cpu_count() or 1))
        one = {k: v[:1] for k, v in inp.items()}
        with torch.no_grad():
            ow.warp_cloth(otps, onet, one["cloth"], one["im_mask"], one["pose_map"])
            t0 = time.perf_counter()
            ow.warp_cloth(otps, onet, one["cloth"], one["im_mask"], one["pose_map"])
            dt = time.perf_counter() - t0
        line["cpu_oracle_images_per_s"] = round(1.0 / dt, 2)
        line["cpu_threads"] = torch.get_num_threads()
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
