"""CPU tests (-m "not gpu"): the oracle against its pins.

The reference ships no tests or golden vectors, and the arithmetic of its hot path lives in
diffusers==0.14.0 (not installed, not vendored).  The pins are therefore:
  1. exact parameter counts of the public SD-2 architectures (865,910,724 / 865,988,484 / 83,653,863 / 7,965,696);
  2. the state-dict key/shape contract (Appendix A.7) shared by oracle and engine;
  3. DDIM known answers (timesteps for N=50/20/100, alpha table endpoints);
  4. tests/golden/tryon_small.npz -- produced by the REFERENCE'S OWN tryon_pipe.py / AutoencoderKL.py / vae.py / emasc.py /
     data_utils.py running on the diffusers shim in the build container (tests/golden/make_golden.py): the restated oracle
     must reproduce it (bit-exact where the CPU matches; 1e-4 across machines).
"""
import math
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tryon_small.npz")


def _count(m):
    return sum(p.numel() for p in m.parameters())


def test_param_counts_known_answers():
    from ladi_oracle.parts import EMASC, InversionAdapter
    from ladi_oracle.unet import UNet2DConditionModel
    from ladi_oracle.vae import AutoencoderKL
    with torch.device("meta"):
        assert _count(UNet2DConditionModel(in_channels=4)) == 865_910_724
        assert _count(UNet2DConditionModel(in_channels=9)) == 865_925_124
        assert _count(UNet2DConditionModel()) == 865_988_484
        assert _count(AutoencoderKL()) == 83_653_863
        assert _count(EMASC([128, 128, 128, 256, 512], [128, 256, 512, 512, 512])) == 7_965_696
        assert _count(InversionAdapter()) == 136_360_704


def test_state_dict_contract_matches_engine():
    from ladi_oracle.unet import UNet2DConditionModel
    from ladi_oracle.vae import AutoencoderKL
    from ladi_vton_b200 import unet_param_shapes, vae_param_shapes
    with torch.device("meta"):
        for mod, shapes in ((UNet2DConditionModel(), unet_param_shapes({})), (AutoencoderKL(), vae_param_shapes({}))):
            sd = mod.state_dict()
            assert set(sd) == set(shapes)
            assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)
    assert sum(math.prod(v) for v in unet_param_shapes({}).values()) == 865_988_484
    for k in ("down_blocks.0.attentions.1.transformer_blocks.0.attn2.to_k.weight", "up_blocks.3.resnets.2.conv_shortcut.weight",
              "mid_block.attentions.0.proj_out.bias", "time_embedding.linear_2.weight", "up_blocks.2.upsamplers.0.conv.bias"):
        assert k in unet_param_shapes({})
    for k in ("encoder.mid_block.attentions.0.query.weight", "decoder.up_blocks.3.resnets.2.conv2.bias", "quant_conv.weight"):
        assert k in vae_param_shapes({})


def test_ddim_known_answers():
    from ladi_oracle.parts import DDIMScheduler as O
    from ladi_vton_b200 import DDIMScheduler as E
    for cls in (O, E):
        s = cls()
        s.set_timesteps(50)
        assert s.timesteps.tolist() == list(range(981, 0, -20))
        s.set_timesteps(20)
        assert s.timesteps.tolist() == list(range(951, 0, -50))
        s.set_timesteps(100)
        assert s.timesteps[0].item() == 991 and s.timesteps[-1].item() == 1
        assert abs(s.alphas_cumprod[0].item() - (1 - 0.00085)) < 1e-7
        assert abs(s.alphas_cumprod[-1].item() - 0.0046602) < 1e-5
    # engine coefficient table == oracle step arithmetic
    e, o = E(), O()
    e.set_timesteps(20); o.set_timesteps(20)
    coef = e.coefficients()
    x, eps = torch.randn(1, 4, 8, 6), torch.randn(1, 4, 8, 6)
    for i, t in enumerate(o.timesteps):
        ref = o.step(eps, t, x).prev_sample
        got = coef[i, 2] * ((x - coef[i, 1] * eps) * coef[i, 0]) + coef[i, 3] * eps
        assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)
    # eta > 0 (DDIMScheduler.step's variance noise): the table's sigma column and the sqrt(1 - a_prev - sigma^2) direction term, with the noise drawn
    # the way the step draws it (one randn_tensor of the model output's shape per step, from the caller's generator)
    for eta in (0.3, 1.0):
        coef = e.coefficients(eta=eta)
        assert coef.shape == (20, 8) and (coef[:, 4] > 0).all() and (coef[:, 5:] == 0).all()
        g_ref, g_eng = torch.Generator().manual_seed(7), torch.Generator().manual_seed(7)
        for i, t in enumerate(o.timesteps):
            ref = o.step(eps, t, x, eta=eta, generator=g_ref).prev_sample
            noise = torch.randn(eps.shape, generator=g_eng)
            got = coef[i, 2] * ((x - coef[i, 1] * eps) * coef[i, 0]) + coef[i, 3] * eps + coef[i, 4] * noise  # ddim_cfg_kernel's arithmetic
            assert torch.allclose(got, ref, rtol=1e-5, atol=2e-6), (eta, i)


def test_timestep_embedding_layout():
    from ladi_oracle.unet import timestep_embedding
    e = timestep_embedding(torch.tensor([0, 981]), 320)
    assert e.shape == (2, 320)
    assert torch.allclose(e[0, :160], torch.ones(160)) and torch.allclose(e[0, 160:], torch.zeros(160))  # [cos, sin]
    assert abs(e[1, 0].item() - math.cos(981.0)) < 1e-4 and abs(e[1, 160].item() - math.sin(981.0)) < 1e-4


def _oracle_small():
    from ladi_oracle.parts import EMASC, DDIMScheduler
    from ladi_oracle.pipeline import OracleTryOnPipeline
    from ladi_oracle.unet import UNet2DConditionModel
    from ladi_oracle.vae import AutoencoderKL
    from ladi_vton_b200 import synthetic as S
    sds = S.build_state_dicts(S.SMALL_UNET, S.SMALL_VAE, seed=1234)
    ou = UNet2DConditionModel(**S.SMALL_UNET).eval(); ou.load_state_dict(sds["unet"])
    ov = AutoencoderKL(**S.SMALL_VAE).eval(); ov.load_state_dict(sds["vae"])
    oe = EMASC(*sds["emasc_channels"]).eval(); oe.load_state_dict(sds["emasc"])
    return OracleTryOnPipeline(ov, ou, DDIMScheduler(), oe, [1, 2, 3, 4, 5]), ov, oe


@pytest.mark.parametrize("tag,gs", [("cfg", 7.5), ("nocfg", 1.0)])
def test_oracle_reproduces_reference_golden(tag, gs):
    from ladi_vton_b200 import synthetic as S
    gold = np.load(GOLD)
    pipe, _, _ = _oracle_small()
    inp = S.synthetic_inputs(2, 128, 64, seed=1234, ctx_dim=128)
    img = pipe(inp["image"], inp["mask_image"], inp["pose_map"], inp["warped_cloth"], inp["prompt_embeds"], inp["negative_prompt_embeds"],
               height=128, width=64, num_inference_steps=3, guidance_scale=gs, generator=torch.Generator().manual_seed(7))
    assert img.shape == gold[f"image_{tag}"].shape
    assert np.abs(img - gold[f"image_{tag}"]).max() < 1e-4


@pytest.mark.parametrize("tag,eta,as_list,with_emasc", [("eta_cfg", 0.6, False, True), ("list_noemasc", 0.0, True, False), ("list_eta_noemasc", 0.6, True, False)])
def test_oracle_reproduces_reference_golden_branches(tag, eta, as_list, with_emasc):
    """tests/golden/tryon_small_branches.npz: the REFERENCE'S OWN tryon_pipe.py run (make_golden.py) with stochastic DDIM (`eta` > 0) and with a list
    of per-sample generators (with emasc=None: the reference's list branch cannot run with EMASC, tryon_pipe.py:452-456) -- the restated oracle, which
    the GPU tests of these branches compare the engine with, must reproduce it."""
    from ladi_oracle.parts import DDIMScheduler
    from ladi_oracle.pipeline import OracleTryOnPipeline
    from ladi_vton_b200 import synthetic as S
    gold = np.load(os.path.join(os.path.dirname(GOLD), "tryon_small_branches.npz"))
    full, ov, oe = _oracle_small()
    pipe = full if with_emasc else OracleTryOnPipeline(ov, full.unet, DDIMScheduler(), None, None)
    inp = S.synthetic_inputs(2, 128, 64, seed=1234, ctx_dim=128)
    gen = [torch.Generator().manual_seed(5), torch.Generator().manual_seed(6)] if as_list else torch.Generator().manual_seed(7)
    img = pipe(inp["image"], inp["mask_image"], inp["pose_map"], inp["warped_cloth"], inp["prompt_embeds"], inp["negative_prompt_embeds"],
               height=128, width=64, num_inference_steps=3, guidance_scale=7.5, generator=gen, eta=eta)
    assert img.shape == gold[f"image_{tag}"].shape
    assert np.abs(img - gold[f"image_{tag}"]).max() < 1e-4


def test_oracle_train_emasc_forward_and_int_layers0_golden():
    """Component-level vectors written by the reference's own classes (make_golden.py): the train_emasc.py:388-403 forward and the decode with
    int_layers containing 0 and 1 (src/models/vae.py:204-210) -- the oracle modules the GPU tests `test_train_emasc_forward_small` and
    `test_vae_decode_int_layers_with_0` use as their reference must reproduce them."""
    from ladi_oracle.parts import EMASC as OE, mask_features
    from ladi_vton_b200 import synthetic as S
    gold = np.load(os.path.join(os.path.dirname(GOLD), "tryon_small_branches.npz"))
    _, ov, oe = _oracle_small()
    with torch.no_grad():
        g = torch.Generator().manual_seed(11)
        image = torch.rand((2, 3, 128, 64), generator=g) * 2 - 1
        mask = torch.zeros((2, 1, 128, 64)); mask[:, :, 30:100, 10:50] = 1
        layers = [1, 2, 3, 4, 5]
        post, _ = ov.encode(image)
        _, feats = ov.encode(image * (1 - mask))
        proc = mask_features(oe([feats[i] for i in layers]), mask)
        torch.manual_seed(3)
        lat = post.latent_dist.sample()
        assert np.abs(lat.numpy() - gold["train_emasc_latents"]).max() < 1e-5
        rec = ov.decode(z=lat, intermediate_features=list(proc), int_layers=layers).sample
        assert np.abs(rec.numpy() - gold["train_emasc_rec"]).max() < 1e-4
        ein, eout = S.emasc_channels(S.SMALL_VAE["block_out_channels"])
        em6 = OE([3] + ein, [3] + eout).eval()
        em6.load_state_dict(S.random_state_dict(S.emasc_param_shapes([3] + ein, [3] + eout), 77))
        x = torch.rand((2, 3, 128, 64), generator=torch.Generator().manual_seed(2)) * 2 - 1
        z = torch.randn((2, 4, 16, 8), generator=torch.Generator().manual_seed(21))
        l6 = [0, 1, 2, 3, 4, 5]
        _, f6 = ov.encode(x)
        inter6 = mask_features(em6([f6[i] for i in l6]), mask)
        out = ov.decode(z, list(inter6), l6).sample
        assert np.abs(out.numpy() - gold["decode_layers0"]).max() < 1e-4


def test_oracle_vae_emasc_golden():
    from ladi_vton_b200 import synthetic as S
    gold = np.load(GOLD)
    _, ov, oe = _oracle_small()
    with torch.no_grad():
        x = S.synthetic_inputs(1, 128, 64, seed=99, ctx_dim=128)["image"]
        enc, feats = ov.encode(x)
        assert np.abs(enc.latent_dist.parameters.numpy() - gold["vae_moments"]).max() < 1e-4
        assert np.abs(feats[3][:, ::4, ::4, ::4].numpy() - gold["vae_skip3_sub"]).max() < 1e-4
        e2 = oe(feats[1:6])[2][:, ::8, ::4, ::4].numpy()
        assert np.abs(e2 - gold["emasc2_sub"]).max() < 1e-4
    assert len(feats) == 6 and feats[1] is feats[2]  # vae.py:100-109: entries 1 and 2 are the same tensor


def test_mask_features_chained_equals_direct():
    """data_utils.py:9-14 resizes the mask in a chain; the engine uses direct nearest /f -- identical for powers of two."""
    import torch.nn.functional as F
    from ladi_oracle.parts import mask_features
    m = (torch.rand(2, 1, 64, 48) > 0.5).float()
    feats = [torch.ones(2, 1, 64, 48), torch.ones(2, 1, 64, 48), torch.ones(2, 1, 32, 24), torch.ones(2, 1, 16, 12), torch.ones(2, 1, 8, 6)]
    out = mask_features(feats, m)
    for o, f in zip(out, (1, 1, 2, 4, 8)):
        assert torch.equal(o, 1 - F.interpolate(m, size=(64 // f, 48 // f)))
        assert torch.equal(o[:, 0], 1 - m[:, 0, ::f, ::f])


# ---- conditioning front-end: oracle/ladi_oracle/clip.py -------------------------------------------------
def test_clip_oracle_matches_transformers():
    """Layer arithmetic + state-dict keys of the restated CLIP towers == the installed transformers classes (shared random weights)."""
    tr = pytest.importorskip("transformers")
    from ladi_oracle import clip as oc
    torch.manual_seed(0)
    cfg = tr.CLIPTextConfig(hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, vocab_size=1000,
                            max_position_embeddings=77, hidden_act="gelu", eos_token_id=999, bos_token_id=998, pad_token_id=0)
    hf = tr.CLIPTextModel(cfg).eval()
    o = oc.ClipTextEncoder(vocab=1000, dim=64, heads=2, layers=2, mlp=256, max_pos=77).eval()
    o.load_state_dict(hf.state_dict(), strict=True)
    ids = torch.randint(1, 900, (3, 77))
    ids[:, 20] = 999
    with torch.no_grad():
        a, b = hf(input_ids=ids), o(ids)
    assert (a.last_hidden_state - b.last_hidden_state).abs().max() < 1e-4
    assert (a.pooler_output - b.pooler_output).abs().max() < 1e-4
    vc = tr.CLIPVisionConfig(hidden_size=80, intermediate_size=320, num_hidden_layers=2, num_attention_heads=2, image_size=56, patch_size=14,
                             hidden_act="gelu")
    hv = tr.CLIPVisionModel(vc).eval()
    ov = oc.ClipVisionEncoder(dim=80, heads=2, layers=2, mlp=320, image=56, patch=14).eval()
    ov.load_state_dict(hv.state_dict(), strict=True)
    px = torch.randn(2, 3, 56, 56)
    with torch.no_grad():
        a, b = hv(pixel_values=px), ov(px)
    assert (a.last_hidden_state - b.last_hidden_state).abs().max() < 1e-4
    assert (a.pooler_output - b.pooler_output).abs().max() < 1e-4


def test_clip_text_golden_and_engine_key_contract():
    """tests/golden/clip_text_small.npz was produced by the reference's own encode_text_word_embedding.py running on the oracle text
    encoder (tests/golden/make_golden_clip.py); the restated function must reproduce it, and the engine's CLIP classes must accept
    exactly the oracle's (= transformers') state-dict keys and shapes."""
    import importlib.util
    from ladi_oracle import clip as oc
    spec = importlib.util.spec_from_file_location("make_golden_clip", os.path.join(os.path.dirname(GOLD), "make_golden_clip.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    gold = np.load(os.path.join(os.path.dirname(GOLD), "clip_text_small.npz"))
    enc, ids, we = mg.build()
    assert np.array_equal(ids.numpy(), gold["input_ids"]) and np.allclose(we.numpy(), gold["word_embeddings"])
    with torch.no_grad():
        out = oc.encode_text_word_embedding(enc, ids, we, 4)
    assert np.abs(out.last_hidden_state.numpy() - gold["last_hidden_state"]).max() < 1e-4
    assert np.abs(out.pooler_output.numpy() - gold["pooler_output"]).max() < 1e-4
    from ladi_vton_b200.clip import CLIPTextModel, CLIPVisionModelWithProjection
    with torch.device("meta"):
        ot = oc.ClipTextEncoder()
        ovis = oc.ClipVisionEncoder()
    assert {k: tuple(v.shape) for k, v in ot.state_dict().items()} == {k: tuple(v) for k, v in CLIPTextModel().param_shapes().items()}
    assert {k: tuple(v.shape) for k, v in ovis.state_dict().items()} == {k: tuple(v) for k, v in CLIPVisionModelWithProjection().param_shapes().items()}
    assert sum(p.numel() for p in ot.parameters()) == 340_387_840   # SD-2 text encoder (OpenCLIP ViT-H text tower, 23 layers kept)
    assert sum(p.numel() for p in ovis.parameters()) == 630_766_080  # CLIP ViT-H/14 vision tower without the projection


# ---- cloth-warping front-end: oracle/ladi_oracle/warp.py -------------------------------------------------
def _load_script(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(os.path.dirname(GOLD), name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_warp_oracle_reproduces_reference_golden():
    """tests/golden/warp_small.npz came from the reference's own ConvNet_TPS / UNetVanilla classes (make_golden_warp.py); the restated
    oracle, fed the same seeded weights and inputs, must reproduce control points, TPS grid, precomputed matrices and U-Net output."""
    from ladi_oracle import warp as ow
    mg = _load_script("make_golden_warp")
    gold = np.load(os.path.join(os.path.dirname(GOLD), "warp_small.npz"))
    tps_sd, unet_sd = mg.build_weights()
    tps = ow.ConvNet_TPS(256, 192, 21, 3).eval()
    missing = tps.load_state_dict(tps_sd, strict=False)
    assert sorted(missing.missing_keys) == ["gridGen.inverse_kernel", "gridGen.padding_matrix", "gridGen.target_coordinate_repr"]
    unet = ow.UNetVanilla(24, 3, True).eval()
    unet.load_state_dict(unet_sd)
    a, b, x = mg.build_inputs()
    with torch.no_grad():
        grid, pts = tps(a, b)
        y = unet(x)
    assert np.abs(tps.gridGen.inverse_kernel.numpy() - gold["inverse_kernel"]).max() < 1e-4
    assert np.abs(tps.gridGen.target_coordinate_repr[::97].numpy() - gold["repr_sub"]).max() < 1e-5
    assert np.abs(pts.numpy() - gold["points"]).max() < 1e-4
    assert np.abs(grid[:, ::8, ::8].numpy() - gold["grid_sub"]).max() < 1e-3
    assert np.abs(y.numpy() - gold["unet_out"]).max() < 1e-3 * max(1.0, float(np.abs(gold["unet_out"]).max()))


def test_warp_key_contract_and_counts():
    from ladi_oracle import warp as ow
    from ladi_vton_b200.warp import ConvNet_TPS, UNetVanilla
    o, e = ow.ConvNet_TPS(256, 192, 21, 3), ConvNet_TPS(256, 192, 21, 3)
    assert {k: tuple(v.shape) for k, v in o.state_dict().items()} == {k: tuple(v) for k, v in e.param_shapes().items()}
    ou, eu = ow.UNetVanilla(24, 3, True), UNetVanilla(24, 3, True)
    assert {k: tuple(v.shape) for k, v in ou.state_dict().items()} == {k: tuple(v) for k, v in eu.param_shapes().items()}
    assert sum(p.numel() for p in o.parameters()) == 19_056_626 and sum(p.numel() for p in ou.parameters()) == 17_275_203
    with pytest.raises(NotImplementedError):
        UNetVanilla(24, 3, bilinear=False)
    with pytest.raises(ValueError):
        ConvNet_TPS(250, 192, 21, 3)


def test_s2d_weight_equals_strided_conv():
    """The engine computes the 4x4 stride-2 pad-1 convolutions of ConvNet_TPS as 3x3 stride-1 convolutions over a space-to-depth
    tensor; the weight re-indexing is host logic and is checked here against torch's strided convolution."""
    import torch.nn.functional as F
    from ladi_vton_b200.warp import s2d_weight
    g = torch.Generator().manual_seed(0)
    x, w = torch.randn((2, 5, 8, 6), generator=g), torch.randn((7, 5, 4, 4), generator=g)
    xs = x.view(2, 5, 4, 2, 3, 2).permute(0, 3, 5, 1, 2, 4).reshape(2, 20, 4, 3)  # channel (sy*2+sx)*c + ch
    assert (F.conv2d(x, w, stride=2, padding=1) - F.conv2d(xs, s2d_weight(w), stride=1, padding=1)).abs().max() < 1e-4


def test_posemap_oracle_matches_reference_golden():
    """oracle/ladi_oracle/dataprep.py == the reference's own src/utils/posemap.py (golden written by tests/golden/make_posemap_golden.py
    from the unmodified file): in-map, missing (no coordinate > 0), half-missing, off-map and tie key-points."""
    import numpy as np
    from ladi_oracle.dataprep import numpy_to_uint8, pose_heatmaps
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "posemap.npz"))
    maps = pose_heatmaps(g["keypoints"], 64, 48, float(g["sigma"])).numpy()
    assert maps.dtype == np.float32 and maps.shape == (20, 64, 48)
    assert np.array_equal(maps, g["maps"])
    assert float(maps[10].max()) == 0.0 and float(maps[11].max()) == 0.0 and float(maps[12].max()) > 0.99  # (0,0), (-3,-1) missing; (0,17.5) kept
    x = np.array([[[[0.0, 0.5, 1.0], [0.00196, 0.00197, 0.49803922]]]], dtype=np.float32)
    assert numpy_to_uint8(x).tolist() == [[[[0, 128, 255], [0, 1, 127]]]]  # 127.5 -> 128 and 127.00000110 -> 127: round half to even


def test_clip_pixel_values_modes_pinned_to_installed_processor():
    """src/inference.py:265-268.  The installed transformers' CLIPImageProcessor (laion ViT-H preprocessor constants) pins two of the
    oracle's modes: [0,1] float input -> "double_rescale" (what versions >= 4.28 do), uint8 input -> "uint8" (the branch transformers
    4.27.3 reaches for float input, because its resize round-trips through PIL uint8)."""
    from transformers import CLIPImageProcessor
    from ladi_oracle.inference_body import CLIP_MEAN, CLIP_STD, clip_pixel_values
    proc = CLIPImageProcessor(do_resize=True, size={"shortest_edge": 224}, do_center_crop=True, crop_size={"height": 224, "width": 224},
                              do_rescale=True, rescale_factor=1 / 255, do_normalize=True, image_mean=list(CLIP_MEAN), image_std=list(CLIP_STD),
                              resample=3)
    g = torch.Generator().manual_seed(4)
    cloth = torch.rand((2, 3, 256, 192), generator=g) * 2.4 - 1.2  # exceeds [-1, 1]: the clamp matters
    x01 = torch.nn.functional.interpolate((cloth + 1) / 2, size=(224, 224), mode="bilinear", antialias=True).clamp(0, 1)
    lib_float = proc(images=x01, return_tensors="pt").pixel_values
    assert float((clip_pixel_values(cloth, "double_rescale") - lib_float).abs().max()) < 1e-6
    u8 = torch.floor(x01 * 255).to(torch.uint8)
    lib_u8 = proc(images=u8, return_tensors="pt").pixel_values
    assert float((clip_pixel_values(cloth, "uint8") - lib_u8).abs().max()) < 1e-5
    d = (clip_pixel_values(cloth, "uint8") - clip_pixel_values(cloth, "float")).abs().max()
    assert 0 < float(d) <= 1 / 255 / min(CLIP_STD) + 1e-6
    with pytest.raises(ValueError):
        clip_pixel_values(cloth, "nope")
