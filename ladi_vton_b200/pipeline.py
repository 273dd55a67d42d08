"""`StableDiffusionTryOnePipeline` -- same constructor, `.to()`, `__call__` signature/defaults/return types and error
behaviour as /root/reference/src/vto_pipelines/tryon_pipe.py:27-765, with the body re-built B200-first:

  * all model arithmetic runs in the hand-written sm_100a kernels (ops.py) -- NHWC bf16 activations, fp32 latents;
  * step-invariant work is hoisted out of the loop (27 of the 31 UNet input channels, text K/V, time-embedding tables);
  * one denoising step (UNet forward + CFG + DDIM update + re-assembly of the next UNet input) is captured ONCE as a CUDA
    graph per (batch, size, CFG) and replayed; the step index lives on the device, so there is no host sync in the loop;
  * RNG draws happen in the reference's order (cloth posterior -> initial latents -> masked-image posterior,
    tryon_pipe.py:640,419,458) with `randn_tensor` semantics, so a CPU generator reproduces the oracle's noise exactly.
"""
import inspect
from dataclasses import dataclass
from typing import Any

import torch

from . import lib, ops


@dataclass
class StableDiffusionPipelineOutput:
    images: Any
    nsfw_content_detected: Any


def _randn(shape, generator, device):
    """diffusers.utils.randn_tensor: sample on the generator's device, then move."""
    gdev = generator.device if generator is not None else device
    return torch.randn(shape, generator=generator, device=gdev, dtype=torch.float32).to(device)


class _Session:
    """Persistent device buffers + the captured step graph for one (B, h, w, cfg) shape."""

    def __init__(self, B, Bp, h, w, in_pitch, device):
        self.unet_in = torch.zeros((Bp, h, w, in_pitch), dtype=torch.bfloat16, device=device)
        self.latents = torch.zeros((B, 4, h, w), dtype=torch.float32, device=device)
        self.step = torch.zeros(2, dtype=torch.int32, device=device)
        self.coef = None
        self.graph = None
        self.guidance = None


class StableDiffusionTryOnePipeline:
    _optional_components = ["safety_checker"]

    def __init__(self, vae, text_encoder, tokenizer, unet, scheduler, safety_checker=None, feature_extractor=None,
                 requires_safety_checker: bool = False, emasc=None, emasc_int_layers=None):
        if safety_checker is not None and feature_extractor is None:
            raise ValueError("Make sure to define a feature extractor when loading {self.__class__} if you want to use the safety"
                             " checker. If you do not want to use the safety checker, you can pass `'safety_checker=None'` instead.")
        self.vae, self.text_encoder, self.tokenizer, self.unet, self.scheduler = vae, text_encoder, tokenizer, unet, scheduler
        self.safety_checker, self.feature_extractor = safety_checker, feature_extractor
        self.emasc, self.emasc_int_layers = emasc, emasc_int_layers
        self.vae_scale_factor = 2 ** (len(self.vae.config.block_out_channels) - 1)
        self.device = torch.device("cpu")
        self.use_cuda_graph = True
        self._sessions = {}

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("ladi_vton_b200 pipeline runs on CUDA (sm_100a) only; there is no CPU path")
        self.device = device
        for m in (self.vae, self.unet, self.emasc, self.text_encoder):
            if m is not None and hasattr(m, "to"):
                m.to(device)
        return self

    @property
    def _execution_device(self):
        return self.device

    # ---- copied semantics: tryon_pipe.py:362-407 ----------------------------------------------------------------------
    def check_inputs(self, prompt, height, width, callback_steps, negative_prompt=None, prompt_embeds=None,
                     negative_prompt_embeds=None):
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if (callback_steps is None) or (callback_steps is not None and (not isinstance(callback_steps, int) or callback_steps <= 0)):
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps} of type {type(callback_steps)}.")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `prompt`: {prompt} and `prompt_embeds`: {prompt_embeds}. Please make sure to"
                             " only forward one of the two.")
        elif prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and `prompt_embeds` undefined.")
        elif prompt is not None and (not isinstance(prompt, str) and not isinstance(prompt, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if negative_prompt is not None and negative_prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `negative_prompt`: {negative_prompt} and `negative_prompt_embeds`:"
                             f" {negative_prompt_embeds}. Please make sure to only forward one of the two.")
        if prompt_embeds is not None and negative_prompt_embeds is not None:
            if prompt_embeds.shape != negative_prompt_embeds.shape:
                raise ValueError("`prompt_embeds` and `negative_prompt_embeds` must have the same shape when passed directly, but"
                                 f" got: `prompt_embeds` {prompt_embeds.shape} != `negative_prompt_embeds` {negative_prompt_embeds.shape}.")

    def prepare_extra_step_kwargs(self, generator, eta):
        params = set(inspect.signature(self.scheduler.step).parameters.keys())
        kw = {}
        if "eta" in params:
            kw["eta"] = eta
        if "generator" in params:
            kw["generator"] = generator
        return kw

    def _encode_text(self, texts, max_length):
        if self.tokenizer is None or self.text_encoder is None:
            raise ValueError("no tokenizer/text_encoder available: pass `prompt_embeds` (and `negative_prompt_embeds` when "
                             "guidance_scale > 1) instead of strings")
        ids = self.tokenizer(texts, padding="max_length", max_length=max_length, truncation=True, return_tensors="pt").input_ids
        return self.text_encoder(ids.to(self.device))[0]

    def _encode_prompt(self, prompt, device, num_images_per_prompt, do_cfg, negative_prompt=None, prompt_embeds=None,
                       negative_prompt_embeds=None):
        """tryon_pipe.py:184-317 -> [B', 77, D] with the CFG batch ordered [negative, positive] (:315)."""
        if prompt is not None:
            batch_size = 1 if isinstance(prompt, str) else len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        if prompt_embeds is None:
            prompt_embeds = self._encode_text(prompt, self.tokenizer.model_max_length if self.tokenizer is not None else 77)
        prompt_embeds = prompt_embeds.to(device=device, dtype=torch.float32)
        bs, seq, _ = prompt_embeds.shape
        prompt_embeds = prompt_embeds.repeat(1, num_images_per_prompt, 1).view(bs * num_images_per_prompt, seq, -1)
        if do_cfg and negative_prompt_embeds is None:
            if negative_prompt is None:
                uncond = [""] * batch_size
            elif type(prompt) is not type(negative_prompt):
                raise TypeError(f"`negative_prompt` should be the same type to `prompt`, but got {type(negative_prompt)} !="
                                f" {type(prompt)}.")
            elif isinstance(negative_prompt, str):
                uncond = [negative_prompt]
            elif batch_size != len(negative_prompt):
                raise ValueError(f"`negative_prompt`: {negative_prompt} has batch size {len(negative_prompt)}, but `prompt`:"
                                 f" {prompt} has batch size {batch_size}. Please make sure that passed `negative_prompt` matches"
                                 " the batch size of `prompt`.")
            else:
                uncond = negative_prompt
            negative_prompt_embeds = self._encode_text(uncond, prompt_embeds.shape[1])
        if do_cfg:
            seq = negative_prompt_embeds.shape[1]
            negative_prompt_embeds = negative_prompt_embeds.to(device=device, dtype=torch.float32)
            negative_prompt_embeds = negative_prompt_embeds.repeat(1, num_images_per_prompt, 1).view(batch_size * num_images_per_prompt, seq, -1)
            prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds])
        return prompt_embeds

    @staticmethod
    def _prepare_mask_and_image(image, mask):
        """diffusers prepare_mask_and_masked_image, tensor branch (called at tryon_pipe.py:630): shape normalisation, range
        checks (raise ValueError) and IN-PLACE binarisation of the caller's mask at 0.5.  The `image * (mask < 0.5)` product
        itself is fused into the layout kernel (ops.nchw_to_nhwc gate)."""
        if not isinstance(image, torch.Tensor) or not isinstance(mask, torch.Tensor):
            raise TypeError("`image` and `mask_image` must be torch tensors (PIL inputs are not supported by this engine)")
        if image.ndim == 3:
            image = image.unsqueeze(0)
        if mask.ndim == 2:
            mask = mask.unsqueeze(0).unsqueeze(0)
        if mask.ndim == 3:
            mask = mask.unsqueeze(0) if mask.shape[0] == 1 else mask.unsqueeze(1)
        assert image.ndim == 4 and mask.ndim == 4, "Image and Mask must have 4 dimensions"
        assert image.shape[-2:] == mask.shape[-2:], "Image and Mask must have the same spatial dimensions"
        assert image.shape[0] == mask.shape[0], "Image and Mask must have the same batch size"
        if image.min() < -1 or image.max() > 1:
            raise ValueError("Image should be in [-1, 1] range")
        if mask.min() < 0 or mask.max() > 1:
            raise ValueError("Mask should be in [0, 1] range")
        mask[mask < 0.5] = 0
        mask[mask >= 0.5] = 1
        return mask, image

    def numpy_to_pil(self, images):
        from PIL import Image
        if images.ndim == 3:
            images = images[None]
        images = (images * 255).round().astype("uint8")
        return [Image.fromarray(im) for im in images]

    # ---- the denoising step that gets captured --------------------------------------------------------------------------
    def _step(self, s, cfg, guidance, advance=True):
        eps = self.unet.forward_nhwc(s.unet_in, s.step)
        ops.ddim_cfg_step(eps, s.latents, s.unet_in, cfg, guidance, s.coef, s.step, advance=advance)

    @torch.no_grad()
    def __call__(self, image, mask_image, pose_map, warped_cloth, prompt=None, height=None, width=None,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5, negative_prompt=None, num_images_per_prompt=1,
                 eta: float = 0.0, prompt_embeds=None, negative_prompt_embeds=None, generator=None, latents=None,
                 output_type="pil", return_dict: bool = True, callback=None, callback_steps=1, cloth_cond_rate: float = 1.0,
                 no_pose: bool = False, cloth_input_type: str = "warped", noise=None):
        """Same arguments as the reference `__call__` (tryon_pipe.py:495-520).  Extensions: output_type="pt" returns the
        device tensor; `noise=(cloth, latents, masked)` supplies the three RNG draws explicitly (batch-sharded runs slice
        one full-batch draw, distributed.draw_noise)."""
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("call .to('cuda') first: ladi_vton_b200 has no CPU path")
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        self.check_inputs(prompt, height, width, callback_steps, negative_prompt, prompt_embeds, negative_prompt_embeds)
        if image is None:
            raise ValueError("`image` input cannot be undefined.")
        if mask_image is None:
            raise ValueError("`mask_image` input cannot be undefined.")
        if eta != 0.0:
            raise NotImplementedError("eta != 0 (stochastic DDIM) is never used by the reference CLI and is not implemented")
        if isinstance(generator, list):
            raise NotImplementedError("per-sample generator lists are not supported; pass one generator")
        if cloth_input_type not in ("warped", "none"):
            raise ValueError(f"Invalid cloth_input_type {cloth_input_type}")
        if prompt is not None:
            batch_size = 1 if isinstance(prompt, str) else len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        cfg = guidance_scale > 1.0
        ctx = self._encode_prompt(prompt, dev, num_images_per_prompt, cfg, negative_prompt, prompt_embeds, negative_prompt_embeds)
        B = batch_size * num_images_per_prompt
        Bp = 2 * B if cfg else B
        sf = self.vae.config.scaling_factor
        vsf = self.vae_scale_factor
        h, w = height // vsf, width // vsf

        # 4. mask / image / pose preprocessing (validation on the tensors' own device, compute in fused layout kernels)
        mask, image = self._prepare_mask_and_image(image, mask_image)
        mask_d = mask.to(dev, torch.float32).contiguous()
        image_d = image.to(dev, torch.float32).contiguous()
        cin = self.unet.config.in_channels
        key = (B, Bp, h, w)
        s = self._sessions.get(key)
        if s is None:
            s = self._sessions[key] = _Session(B, Bp, h, w, self.unet.in_pitch, dev)
        s.unet_in.zero_()
        n_pose = pose_map.shape[1]
        c_mask, c_masked, c_pose, c_cloth = 4, 5, 9, 9 + n_pose
        assert cin == c_cloth + (4 if cloth_input_type == "warped" else 0), "UNet in_channels does not match the conditioning"
        cond = s.unet_in[B:] if cfg else s.unet_in  # conditional half of the CFG batch ([uncond, cond], tryon_pipe.py:315,702-705)
        if not no_pose:
            pose_d = ops.bilinear_down8(pose_map.to(dev, torch.float32).contiguous())  # :632-634
            ops.nchw_to_nhwc(pose_d, cond, c_off=c_pose)
        # 4b. warped cloth latents (RNG draw #1)
        if cloth_input_type == "warped":
            mom, _ = self.vae.encode_nhwc(warped_cloth)
            cloth = ops.posterior_sample(mom, noise[0].to(dev) if noise is not None else _randn((B, 4, h, w), generator, dev), sf)
            ops.nchw_to_nhwc(cloth, cond, c_off=c_cloth)
        # 5. timesteps, 6. latents (RNG draw #2)
        self.scheduler.set_timesteps(num_inference_steps, device=dev)
        ts = self.scheduler.timesteps_host
        cloth_steps = (1 - cloth_cond_rate) * num_inference_steps
        if latents is None:
            latents = noise[1] if noise is not None else _randn((B, 4, h, w), generator, dev)
        s.latents.copy_(latents.to(dev, torch.float32) * self.scheduler.init_noise_sigma)
        # 7. masked image -> latents + encoder skips (RNG draw #3); EMASC with mask_features fused
        masked = torch.zeros((B, height, width, 8), dtype=torch.bfloat16, device=dev)
        ops.nchw_to_nhwc(image_d, masked, gate=mask_d)  # image * (mask < 0.5)
        mom, feats = self.vae.encode_nhwc(masked, nhwc=True)
        masked_lat = ops.posterior_sample(mom, noise[2].to(dev) if noise is not None else _randn((B, 4, h, w), generator, dev), sf)
        inter = None
        if self.emasc:
            sel = [feats[i] for i in self.emasc_int_layers]  # :460-461
            inv = [ops.inv_mask_rows(mask_d, height // f.shape[1]) for f in sel]  # data_utils.py:9-14 (chained nearest == direct)
            inter = self.emasc(sel, inv)  # :684-685
        for half in ((s.unet_in[:B], s.unet_in[B:]) if cfg else (s.unet_in,)):  # :482-485
            ops.nchw_to_nhwc(mask_d, half, c_off=c_mask, f=vsf)  # nearest /8 (:434-436)
            ops.nchw_to_nhwc(masked_lat, half, c_off=c_masked)
            ops.nchw_to_nhwc(s.latents, half, c_off=0)
        # step-invariant UNet work
        self.unet.plan_context(ctx)
        self.unet.plan_steps(ts)
        coef = self.scheduler.coefficients()
        if s.coef is None or s.coef.shape != coef.shape:
            s.coef = coef.to(dev)
        else:
            s.coef.copy_(coef)
        s.step.zero_()
        # 9. denoising loop
        n_zero_from = num_inference_steps - cloth_steps  # :718-719
        use_graph = self.use_cuda_graph and callback is None
        for i in range(num_inference_steps):
            if cloth_input_type == "warped" and i >= n_zero_from and cloth_steps > 0:
                s.unet_in[..., c_cloth:c_cloth + 4].zero_()
            stale = s.graph is None or s.guidance != guidance_scale or s.graph_key != self._graph_key() + (s.coef.data_ptr(),)
            if not use_graph or (i == 0 and stale):
                self._step(s, cfg, guidance_scale)  # first step of a new shape runs eagerly: the warm-up before capture
            else:
                if stale:
                    s.graph = torch.cuda.CUDAGraph()
                    n0 = lib.launches
                    with torch.cuda.graph(s.graph):
                        self._step(s, cfg, guidance_scale)
                    s.graph_nodes = lib.launches - n0
                    lib.launches = n0  # capture records, it does not launch
                    s.guidance, s.graph_key = guidance_scale, self._graph_key() + (s.coef.data_ptr(),)
                    # capture records but does not execute: fall through to the replay below
                s.graph.replay()
                lib.launches += s.graph_nodes
            if callback is not None and i % callback_steps == 0:
                callback(i, ts[i], s.latents)
        # 11. decode with the EMASC skips, clamp, D2H
        img = self.vae.decode_nhwc(s.latents, inter, self.emasc_int_layers if inter is not None else None, scale=1.0 / sf)
        if output_type == "pil":  # numpy_to_pil's (x*255).round().astype(uint8) on the device: a quarter of the D2H bytes (:358-760)
            from PIL import Image
            u8 = ops.image_out_u8(img).cpu().numpy()
            pil = [Image.fromarray(im) for im in u8]
            return StableDiffusionPipelineOutput(images=pil, nsfw_content_detected=None) if return_dict else (pil, None)
        out = ops.image_out(img)  # [B, H, W, 3] fp32 in [0, 1]  (:356)
        if output_type == "pt":  # extension: leave the result on the device (used for device-resident timing / NCCL gather)
            return StableDiffusionPipelineOutput(images=out, nsfw_content_detected=None) if return_dict else (out, None)
        out = out.cpu().numpy()  # (:358)
        if output_type == "pil":
            out = self.numpy_to_pil(out)
        if not return_dict:
            return (out, None)
        return StableDiffusionPipelineOutput(images=out, nsfw_content_detected=None)

    def _graph_key(self):
        # the captured graph bakes in the addresses of the planned step table / text K/V buffers
        return (self.unet._steps.data_ptr(), self.unet._ctx.data_ptr(), self.unet._steps.shape[0])

    def _coef_ptr(self, s):
        return s.coef.data_ptr()
