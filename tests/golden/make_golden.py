"""Generates tests/golden/tryon_small.npz (and tryon_small_branches.npz: eta > 0, generator lists) by running the REFERENCE'S OWN FILES -- /root/reference/src/vto_pipelines/
tryon_pipe.py, src/models/AutoencoderKL.py, src/models/vae.py, src/models/emasc.py, src/utils/data_utils.py, imported
unmodified -- on the diffusers shim (oracle/shim), CPU fp32, with the seeded small-config weights and synthetic inputs
that the tests rebuild.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The fixture pins oracle/ladi_oracle (test_oracle_pins.py re-runs the restated oracle against it without /root/reference)
and is the end-to-end golden vector for the GPU parity tests.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "shim"), "/root/reference"):
    sys.path.insert(0, p)

from ladi_oracle.parts import DDIMScheduler  # noqa: E402
from ladi_oracle.unet import UNet2DConditionModel  # noqa: E402
from ladi_vton_b200 import synthetic as S  # noqa: E402
from src.models.AutoencoderKL import AutoencoderKL as RefVAE  # noqa: E402  (reference file)
from src.models.emasc import EMASC as RefEMASC  # noqa: E402  (reference file)
from src.vto_pipelines.tryon_pipe import StableDiffusionTryOnePipeline  # noqa: E402  (reference file)


class _TE(torch.nn.Module):  # stand-in exposing .dtype, as read at tryon_pipe.py:255
    dtype = torch.float32


def main():
    torch.set_num_threads(4)
    sds = S.build_state_dicts(S.SMALL_UNET, S.SMALL_VAE, seed=1234)
    unet = UNet2DConditionModel(**S.SMALL_UNET).eval()
    unet.load_state_dict(sds["unet"])
    vch = S.SMALL_VAE["block_out_channels"]
    vae = RefVAE(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
                 block_out_channels=vch, layers_per_block=2, latent_channels=4, norm_num_groups=32, sample_size=128).eval()
    vae.load_state_dict(sds["vae"])
    emasc = RefEMASC(*sds["emasc_channels"]).eval()
    emasc.load_state_dict(sds["emasc"])
    out = {}
    for tag, gs in (("cfg", 7.5), ("nocfg", 1.0)):
        inp = S.synthetic_inputs(2, 128, 64, seed=1234, ctx_dim=128)
        pipe = StableDiffusionTryOnePipeline(vae=vae, text_encoder=_TE(), tokenizer=None, unet=unet, scheduler=DDIMScheduler(),
                                             emasc=emasc, emasc_int_layers=[1, 2, 3, 4, 5])
        img = pipe(image=inp["image"], mask_image=inp["mask_image"], pose_map=inp["pose_map"], warped_cloth=inp["warped_cloth"],
                   prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"], height=128, width=64,
                   num_inference_steps=3, guidance_scale=gs, generator=torch.Generator().manual_seed(7), output_type="np").images
        out[f"image_{tag}"] = img.astype(np.float32)
    # component-level vectors from the reference VAE fork / EMASC
    with torch.no_grad():
        x = S.synthetic_inputs(1, 128, 64, seed=99, ctx_dim=128)["image"]
        enc, feats = vae.encode(x)
        out["vae_moments"] = enc.latent_dist.parameters.numpy()
        out["vae_skip3_sub"] = feats[3][:, ::4, ::4, ::4].contiguous().numpy()      # strided subsample keeps the fixture small
        out["emasc2_sub"] = emasc([f.clone() for f in feats[1:6]])[2][:, ::8, ::4, ::4].contiguous().numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tryon_small.npz"), **out)
    print({k: (v.shape, float(np.abs(v).mean())) for k, v in out.items()})
    # ---- the less-travelled branches of the reference `__call__` (round 2): stochastic DDIM (`eta` > 0: tryon_pipe.py:337-345,740) and a LIST of
    # per-sample generators (prepare_latents :412-416, prepare_mask_latents :445-450 -- with EMASC that branch of the reference indexes the per-sample
    # list with the layer indices (:452-456) and cannot run, so the list cases use emasc=None, where the decode takes no intermediate features)
    br = {}
    cases = (("eta_cfg", 0.6, False, True), ("list_noemasc", 0.0, True, False), ("list_eta_noemasc", 0.6, True, False))
    for tag, eta, as_list, with_emasc in cases:
        inp = S.synthetic_inputs(2, 128, 64, seed=1234, ctx_dim=128)
        pipe = StableDiffusionTryOnePipeline(vae=vae, text_encoder=_TE(), tokenizer=None, unet=unet, scheduler=DDIMScheduler(),
                                             emasc=emasc if with_emasc else None, emasc_int_layers=[1, 2, 3, 4, 5] if with_emasc else None)
        gen = [torch.Generator().manual_seed(5), torch.Generator().manual_seed(6)] if as_list else torch.Generator().manual_seed(7)
        img = pipe(image=inp["image"], mask_image=inp["mask_image"], pose_map=inp["pose_map"], warped_cloth=inp["warped_cloth"],
                   prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"], height=128, width=64,
                   num_inference_steps=3, guidance_scale=7.5, eta=eta, generator=gen, output_type="np").images
        br[f"image_{tag}"] = img.astype(np.float32)
    # ---- component level, reference classes only: (a) the train_emasc.py:388-403 forward (posterior of the image, skips of the masked image, EMASC,
    # mask_features, decode of the posterior SAMPLE); (b) decode with int_layers containing 0 and 1 (src/models/vae.py:204-210), six-scale EMASC
    from src.utils.data_utils import mask_features  # noqa: E402  (reference file)
    with torch.no_grad():
        g = torch.Generator().manual_seed(11)
        image = torch.rand((2, 3, 128, 64), generator=g) * 2 - 1
        mask = torch.zeros((2, 1, 128, 64)); mask[:, :, 30:100, 10:50] = 1
        int_layers = [1, 2, 3, 4, 5]
        posterior_im, _ = vae.encode(image)
        _, feats = vae.encode(image * (1 - mask))
        proc = mask_features(emasc([feats[i] for i in int_layers]), mask)
        torch.manual_seed(3)  # the reference samples from the global RNG here (train_emasc.py:400: latent_dist.sample() without a generator)
        lat = posterior_im.latent_dist.sample()
        rec = vae.decode(z=lat, intermediate_features=proc, int_layers=int_layers).sample
        br["train_emasc_latents"], br["train_emasc_rec"] = lat.numpy(), rec.numpy()
        ein, eout = S.emasc_channels(vch)
        em6 = RefEMASC([3] + ein, [3] + eout).eval()
        em6.load_state_dict(S.random_state_dict(S.emasc_param_shapes([3] + ein, [3] + eout), 77))
        x = torch.rand((2, 3, 128, 64), generator=torch.Generator().manual_seed(2)) * 2 - 1
        z = torch.randn((2, 4, 16, 8), generator=torch.Generator().manual_seed(21))
        layers = [0, 1, 2, 3, 4, 5]
        _, f6 = vae.encode(x)
        inter6 = mask_features(em6([f6[i] for i in layers]), mask)
        br["decode_layers0"] = vae.decode(z, list(inter6), layers).sample.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tryon_small_branches.npz"), **br)
    print({k: (v.shape, float(np.abs(v).mean())) for k, v in br.items()})


if __name__ == "__main__":
    main()
