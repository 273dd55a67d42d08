"""Times the conditioning front-end on cuda:0: CLIP text tower (encode_text_word_embedding, B prompts x 77
tokens, 16 pseudo-words) and CLIP ViT-H vision tower (B x 224 x 224), full-size random-init weights, CUDA events, after warm-up;
beside each the fp32 oracle on the host cores (one call).  Prints one JSON line per tower.

    python tools/frontend_bench.py [--batch 8] [--no-cpu]
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


def timed(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    from ladi_vton_b200 import CLIPTextModel, CLIPVisionModelWithProjection, encode_text_word_embedding, lib, synthetic as S
    dev = torch.device("cuda:0")
    B = a.batch
    g = torch.Generator().manual_seed(1)
    # ---- text tower
    txt = CLIPTextModel()
    sd = S.random_state_dict(txt.param_shapes(), 1, device="cuda", fast=True)
    txt.load_state_dict(sd).to(dev)
    ids = torch.randint(1, 250, (B, 77), generator=g)
    ids[:, 30:] = 49407
    ids[:, 10:26] = 259
    we = torch.randn((B, 16, 1024), generator=g).to(dev)
    n0 = lib.launches
    ms = timed(lambda: encode_text_word_embedding(txt, ids, we, 16))
    launches = (lib.launches - n0) // 13
    c = txt.config
    flops = B * 77 * 2 * c.num_hidden_layers * (4 * c.hidden_size ** 2 + 2 * c.hidden_size * c.intermediate_size) + \
        c.num_hidden_layers * 4 * B * c.num_attention_heads * 77 * 77 * 64
    line = dict(tower="clip_text", batch=B, ms=round(ms, 3), tflops=round(flops / ms / 1e9, 1), gpu_launches=launches,
                weights_gb=round(sum(v.numel() for v in sd.values()) * 2 / 1e9, 3))
    if not a.no_cpu:
        from ladi_oracle.clip import ClipTextEncoder, encode_text_word_embedding as ofn
        with torch.device("meta"):
            o = ClipTextEncoder()
        o.load_state_dict({k: v.cpu() for k, v in sd.items()}, assign=True)
        torch.set_num_threads(min(64, os.cpu_count() or 1))
        with torch.no_grad():
            ofn(o, ids[:1], we[:1].cpu(), 16)
            t0 = time.perf_counter()
            ofn(o, ids, we.cpu(), 16)
            line["cpu_oracle_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
            line["cpu_threads"] = torch.get_num_threads()
        del o
    print(json.dumps(line), flush=True)
    del txt, sd
    torch.cuda.empty_cache()
    # ---- vision tower
    vis = CLIPVisionModelWithProjection()
    sd = S.random_state_dict(vis.param_shapes(), 2, device="cuda", fast=True)
    vis.load_state_dict(sd).to(dev)
    px = torch.randn((B, 3, 224, 224), generator=g).to(dev)
    n0 = lib.launches
    ms = timed(lambda: vis(px))
    launches = (lib.launches - n0) // 13
    c = vis.config
    T = 257
    flops = B * T * 2 * c.num_hidden_layers * (4 * c.hidden_size ** 2 + 2 * c.hidden_size * c.intermediate_size) + \
        c.num_hidden_layers * 4 * B * c.num_attention_heads * T * T * 80 + B * 256 * 2 * 588 * c.hidden_size
    line = dict(tower="clip_vision", batch=B, ms=round(ms, 3), tflops=round(flops / ms / 1e9, 1), gpu_launches=launches,
                weights_gb=round(sum(v.numel() for v in sd.values()) * 2 / 1e9, 3))
    if not a.no_cpu:
        from ladi_oracle.clip import ClipVisionEncoder
        with torch.device("meta"):
            o = ClipVisionEncoder()
        o.load_state_dict({k: v.cpu() for k, v in sd.items()}, assign=True)
        with torch.no_grad():
            o(px[:1].cpu())
            t0 = time.perf_counter()
            o(px.cpu())
            line["cpu_oracle_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
            line["cpu_threads"] = torch.get_num_threads()
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
