"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

fp32 restatement of the batch body of the reference CLI, /root/reference/src/inference.py:226-312, composed from the other oracle modules:
warp the cloth (:236-263, warp.warp_cloth) -> CLIP pixel values (:265-268) -> vision tower last_hidden_state (:269-273) -> inversion
adapter -> [B, num_vstar, D] pseudo-word embeddings (:276-277) -> prompt strings (:279-286) -> tokenizer (:289-291) ->
encode_text_word_embedding(...).last_hidden_state (:294-295) -> try-on pipeline (:298-311), where the '' prompt of classifier-free
guidance is encoded through the same tokenizer + text encoder (tryon_pipe.py:284-301).

The CLIP image processor of :267 (`AutoProcessor` -> CLIPImageProcessor, transformers 4.27.3, not installable here) receives float
images that are already 224x224 in [0,1].  What it does to them is version behaviour, restated as three modes of `clip_pixel_values`:
  * "uint8" (default): 4.27.3's `resize` round-trips every non-PIL image through PIL -- `to_pil_image` multiplies [0,1] floats by 255 and
    truncates to uint8 -- and returns the uint8 array; `rescale(1/255)` and `normalize` follow: (floor(v*255)/255 - mean)/std.
    [recalled from the 4.27 source; pinned here only through the installed library's identical uint8 branch, tests/test_oracle_pins.py]
  * "double_rescale": later versions (and the installed 5.5, pinned in tests/test_oracle_pins.py) keep the float range through the
    resize and then apply `rescale(1/255)` to the already-[0,1] image: (v/255 - mean)/std.
  * "float": (v - mean)/std, no quantisation.
mean/std = `preprocessor_config.json` of laion/CLIP-ViT-H-14-laion2B-s32B-b79K.  The CLI takes the real processor object when its files
are local, which settles the question for a given installation.
"""
import torch
import torch.nn.functional as F

from .clip import encode_text_word_embedding
from .warp import warp_cloth

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
CATEGORY_TEXT = {'dresses': 'a dress', 'upper_body': 'an upper body garment', 'lower_body': 'a lower body garment'}  # :279-283


def clip_pixel_values(cloth, mode="uint8"):
    """src/inference.py:265-268."""
    x = F.interpolate((cloth.float() + 1) / 2, size=(224, 224), mode="bilinear", antialias=True, align_corners=False).clamp(0, 1)
    mean = torch.tensor(CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD).view(1, 3, 1, 1)
    if mode == "uint8":
        x = torch.floor(x * 255) / 255
    elif mode == "double_rescale":
        x = x / 255
    elif mode != "float":
        raise ValueError(mode)
    return (x - mean) / std


def prompts_for(categories, num_vstar):
    return [f'a photo of a model wearing {CATEGORY_TEXT[c]} {" $ " * num_vstar}' for c in categories]  # :285-286


@torch.no_grad()
def run_batch(batch, tps, refinement, vision_encoder, inversion_adapter, tokenizer, text_encoder, pipe, num_vstar, guidance_scale,
              num_inference_steps, generator, size=(512, 384), return_all=False):
    warped = warp_cloth(tps, refinement, batch["cloth"], batch["im_mask"], batch["pose_map"])
    feats = vision_encoder(clip_pixel_values(batch["cloth"])).last_hidden_state
    word = inversion_adapter(feats)
    word = word.reshape((word.shape[0], num_vstar, -1))
    ids = tokenizer(prompts_for(batch["category"], num_vstar), max_length=tokenizer.model_max_length, padding="max_length", truncation=True,
                    return_tensors="pt").input_ids
    ctx = encode_text_word_embedding(text_encoder, ids, word, num_vstar).last_hidden_state
    neg = None
    if guidance_scale > 1.0:  # tryon_pipe.py:284-301
        nid = tokenizer([""] * ids.shape[0], max_length=ctx.shape[1], padding="max_length", truncation=True, return_tensors="pt").input_ids
        neg = text_encoder(nid).last_hidden_state
    images = pipe(batch["image"].float(), batch["inpaint_mask"].float(), batch["pose_map"].float(), warped, ctx, neg, height=size[0],
                  width=size[1], num_inference_steps=num_inference_steps, guidance_scale=guidance_scale, generator=generator)
    return (images, warped, feats, word, ctx) if return_all else images
