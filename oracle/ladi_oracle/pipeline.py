"""ORACLE (test infrastructure only) -- CPU restatement of
`StableDiffusionTryOnePipeline.__call__` (/root/reference/src/vto_pipelines/tryon_pipe.py:494-765),
driven with `prompt_embeds`/`negative_prompt_embeds` (the tokenizer / text encoder are unavailable
offline).

PARITY STATUS: pinned bit-exactly (fp32, CPU) against the reference's own tryon_pipe.py /
AutoencoderKL.py / vae.py / emasc.py / data_utils.py executed on the diffusers shim in the build
container (tests/golden/make_golden.py writes the fixture; tests/test_oracle_pins.py re-checks it
without /root/reference).  The diffusers arithmetic underneath is "parity unpinned" upstream.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .parts import mask_features, prepare_mask_and_masked_image, randn_tensor


class OracleTryOnPipeline:
    def __init__(self, vae, unet, scheduler, emasc=None, emasc_int_layers=None):
        self.vae, self.unet, self.scheduler, self.emasc = vae, unet, scheduler, emasc
        self.emasc_int_layers = emasc_int_layers
        self.vae_scale_factor = 2 ** (len(vae.config.block_out_channels) - 1)  # tryon_pipe.py:138
        self.trace = None  # optional dict collecting intermediates for block-level parity tests

    @staticmethod
    def _sample(dist, generator):
        """DiagonalGaussianDistribution.sample (src/models/vae.py:341-347) = mean + std * randn_tensor(mean.shape, generator)."""
        return dist.mean + dist.std * randn_tensor(dist.mean.shape, generator=generator, dtype=dist.parameters.dtype)

    def _rec(self, k, v):
        if self.trace is not None:
            self.trace[k] = v.detach().clone() if torch.is_tensor(v) else v

    @torch.no_grad()
    def __call__(self, image, mask_image, pose_map, warped_cloth, prompt_embeds, negative_prompt_embeds=None,
                 height=None, width=None, num_inference_steps=50, guidance_scale=7.5, generator=None, latents=None,
                 output_type="np", cloth_cond_rate=1.0, no_pose=False, cloth_input_type="warped", eta=0.0):
        height = height or self.unet.config.sample_size * self.vae_scale_factor  # :584-585
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        if height % 8 or width % 8:  # :372-373
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        B = prompt_embeds.shape[0]  # :610
        cfg = guidance_scale > 1.0  # :617
        wdtype = self.unet.dtype
        ctx = prompt_embeds.to(wdtype)  # :255 (text_encoder.dtype == weight dtype in inference.py)
        if cfg:
            ctx = torch.cat([negative_prompt_embeds.to(wdtype), ctx])  # :315  [neg, pos]
        mask, masked_image = prepare_mask_and_masked_image(image, mask_image)  # :630
        pose = F.interpolate(pose_map, size=(pose_map.shape[2] // 8, pose_map.shape[3] // 8), mode="bilinear")  # :632
        if no_pose:
            pose = torch.zeros_like(pose)
        sf = self.vae.config.scaling_factor
        cloth = None
        if cloth_input_type == "warped":  # :639-647, RNG draw #1
            cloth = sf * self._sample(self.vae.encode(warped_cloth)[0].latent_dist, generator)
        elif cloth_input_type != "none":
            raise ValueError(f"Invalid cloth_input_type {cloth_input_type}")
        self.scheduler.set_timesteps(num_inference_steps)  # :650
        timesteps = self.scheduler.timesteps
        cloth_steps = (1 - cloth_cond_rate) * num_inference_steps  # :654
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        if latents is None:  # :410-425, RNG draw #2
            latents = randn_tensor((B, self.vae.config.latent_channels, h, w), generator=generator, dtype=ctx.dtype)
        latents = latents * self.scheduler.init_noise_sigma
        # prepare_mask_latents :427-492, RNG draw #3
        mask_l = F.interpolate(mask, size=(h, w)).to(ctx.dtype)
        post, feats = self.vae.encode(masked_image.to(ctx.dtype))
        masked_l = sf * self._sample(post.latent_dist, generator)  # :445-450 samples per image with generator[i]: the same draws
        self._rec("cloth_latents", cloth)
        self._rec("masked_latents", masked_l)
        inter = None
        if self.emasc is not None:
            inter = [feats[i] for i in self.emasc_int_layers]  # :460-461
            inter = mask_features(self.emasc(inter), mask_image)  # :684-685
            for i, f in enumerate(inter):
                self._rec(f"emasc{i}", f)
        if cfg:  # :482-485, :702-705
            mask_l, masked_l = torch.cat([mask_l] * 2), torch.cat([masked_l] * 2)
            pose = torch.cat([torch.zeros_like(pose), pose])
            if cloth is not None:
                cloth = torch.cat([torch.zeros_like(cloth), cloth])
        for i, t in enumerate(timesteps):  # :713-747
            x = torch.cat([latents] * 2) if cfg else latents
            if cloth is not None and i >= num_inference_steps - cloth_steps:
                cloth = torch.zeros_like(cloth)
            parts = [x, mask_l, masked_l, pose.to(mask_l.dtype)]
            if cloth is not None:
                parts.append(cloth.to(mask_l.dtype))
            x = torch.cat(parts, dim=1)  # latents4, mask1, masked4, pose18, cloth4
            eps = self.unet(x, t, encoder_hidden_states=ctx).sample
            if i == 0:
                self._rec("unet_in0", x)
                self._rec("eps0", eps)
            if cfg:
                e_u, e_t = eps.chunk(2)
                eps = e_u + guidance_scale * (e_t - e_u)
            latents = self.scheduler.step(eps, t, latents, eta=eta, generator=generator).prev_sample.to(self.vae.dtype)  # :337-345,740
        self._rec("final_latents", latents)
        z = 1 / sf * latents  # :349-359
        img = (self.vae.decode(z, inter, self.emasc_int_layers).sample if inter
               else self.vae.decode(z).sample)
        img = (img / 2 + 0.5).clamp(0, 1).cpu().permute(0, 2, 3, 1).float().numpy()
        if output_type == "uint8":
            img = (img * 255).round().astype(np.uint8)  # numpy_to_pil arithmetic, :760
        return img
