"""GPU parity of the I/O-edge kernels: pose heat-maps (src/utils/posemap.py, golden from the reference's own file)
and numpy_to_pil's uint8 conversion.  Tolerances: heat-maps fp32 vs the reference's float64 arithmetic 2e-6 absolute; uint8 bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pose_heatmaps_golden_and_full_size(cuda):
    from ladi_oracle.dataprep import pose_heatmaps as oracle
    from ladi_vton_b200 import ops
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "posemap.npz"))
    kp = torch.tensor(g["keypoints"], dtype=torch.float32)
    out = ops.pose_heatmaps(kp.to(cuda), 64, 48, float(g["sigma"])).cpu().numpy()
    assert out.shape == (20, 64, 48)
    assert float(np.abs(out - g["maps"]).max()) < 2e-6
    assert float(out[10].max()) == 0.0 and float(out[11].max()) == 0.0
    # the dataset's shape: batch 4 x 18 joints at 512x384, key-points drawn like OpenPose output scaled to the image (vitonhd.py:240-246)
    gen = torch.Generator().manual_seed(5)
    kp = torch.rand((4, 18, 2), generator=gen) * torch.tensor([384.0, 512.0])
    kp[1, 3] = 0.0  # undetected joint
    kp[2, 7] = torch.tensor([400.0, 100.0])  # just outside
    ref = oracle(kp.numpy(), 512, 384, 9.0).numpy()
    out = ops.pose_heatmaps(kp.to(cuda), 512, 384, 9.0).cpu().numpy()
    assert out.shape == ref.shape == (4, 18, 512, 384)
    assert float(np.abs(out - ref).max()) < 2e-6 and float(out[1, 3].max()) == 0.0
    with pytest.raises(RuntimeError):
        ops.pose_heatmaps(kp.to(cuda), 512, 384, 0.0)


def test_image_out_u8_bit_exact(cuda):
    from ladi_oracle.dataprep import numpy_to_uint8
    from ladi_vton_b200 import ops
    gen = torch.Generator().manual_seed(3)
    x = (torch.randn((2, 40, 24, 8), generator=gen) * 0.8).to(cuda)
    x[0, 0, 0, :3] = torch.tensor([0.0, -1.0, 1.0])  # 127.5 tie (-> 128), floor and ceiling
    for t in (x, x.bfloat16()):
        f = ops.image_out(t).cpu().numpy()
        u = ops.image_out_u8(t).cpu().numpy()
        assert u.dtype == np.uint8 and u.shape == (2, 40, 24, 3)
        assert np.array_equal(u, numpy_to_uint8(f))


def test_pipeline_pil_equals_numpy_to_pil(cuda):
    """output_type='pil' (device-side uint8) == numpy_to_pil(output_type='np') byte for byte."""
    from ladi_vton_b200 import synthetic as S
    pipe, _ = S.build_pipeline(cuda, S.SMALL_UNET, S.SMALL_VAE)
    inp = S.synthetic_inputs(2, 128, 64, ctx_dim=128)
    kw = dict(pose_map=inp["pose_map"], warped_cloth=inp["warped_cloth"], prompt_embeds=inp["prompt_embeds"],
              negative_prompt_embeds=inp["negative_prompt_embeds"], height=128, width=64, num_inference_steps=2, guidance_scale=7.5)
    a = pipe(image=inp["image"].clone(), mask_image=inp["mask_image"].clone(), generator=torch.Generator().manual_seed(1), output_type="pil", **kw).images
    b = pipe(image=inp["image"].clone(), mask_image=inp["mask_image"].clone(), generator=torch.Generator().manual_seed(1), output_type="np", **kw).images
    ref = pipe.numpy_to_pil(b)
    assert len(a) == 2 and a[0].size == (64, 128) and a[0].mode == "RGB"
    for x, y in zip(a, ref):
        assert np.array_equal(np.asarray(x), np.asarray(y))


def test_pose_map_from_keypoints(cuda):
    """vitonhd.py:240-287: BODY_25 json list -> rescale -> 18 COCO joints -> heat-maps, against the oracle loop."""
    from ladi_oracle.dataprep import kpoint_to_heatmap as oracle_one
    from ladi_vton_b200.data import get_coco_body25_mapping, kpoint_to_heatmap, pose_map_from_keypoints
    gen = torch.Generator().manual_seed(9)
    raw = torch.rand((2, 25, 3), generator=gen) * torch.tensor([768.0, 1024.0, 1.0])
    raw[0, 4, :2] = 0.0  # undetected wrist
    pm = pose_map_from_keypoints(raw.reshape(2, 75).tolist(), 512, 384).cpu()
    assert pm.shape == (2, 18, 512, 384)
    m = get_coco_body25_mapping()
    for b in range(2):
        pd = raw[b, :, :2].double().numpy().copy()
        pd[:, 0] *= 384 / 768
        pd[:, 1] *= 512 / 1024
        for i in (0, 4, 8, 17):
            ref = oracle_one(pd[m[i]], (512, 384), 9)
            assert float((pm[b, i] - ref).abs().max()) < 2e-6
    assert float(pm[0, 4].max()) == 0.0
    one = kpoint_to_heatmap([100.5, 200.25], (512, 384), 9).cpu()
    assert float((one - oracle_one([100.5, 200.25], (512, 384), 9)).abs().max()) < 2e-6
