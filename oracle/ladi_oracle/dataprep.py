"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

Restatement of the pose heat-map edge of the dataset tensorisation, /root/reference/src/utils/posemap.py:6-35 (`kpoint_to_heatmap`), as
the reference datasets call it per joint with sigma 9 (src/dataset/vitonhd.py:277-287), and of `numpy_to_pil`'s uint8 conversion
(diffusers DiffusionPipeline.numpy_to_pil, called at src/vto_pipelines/tryon_pipe.py:760).  Pinned against the reference's own
posemap.py (imported unmodified by tests/golden/make_posemap_golden.py -> tests/golden/posemap.npz)."""
import numpy as np
import torch


def kpoint_to_heatmap(kpoint, shape, sigma):
    """posemap.py:6-35: float64 numpy arithmetic, result cast to float32 by torch.Tensor(...)."""
    map_h, map_w = shape
    kpoint = np.asarray(kpoint, dtype=np.float64)
    if not np.any(kpoint > 0):  # :24 (a key-point at or left/above the origin in BOTH coordinates is "missing")
        return torch.zeros((map_h, map_w))
    x, y = kpoint
    ys, xs = np.meshgrid(np.arange(map_h, dtype=np.float64), np.arange(map_w, dtype=np.float64), indexing="ij")
    heat = np.exp(-((xs - x) ** 2 + (ys - y) ** 2) / sigma ** 2)  # :29 (xy_grid[y, x] = (x, y))
    heat /= heat.max() + np.finfo("float32").eps  # :30
    return torch.tensor(heat, dtype=torch.float32)


def pose_heatmaps(keypoints, h, w, sigma=9.0):
    """[..., 2] key-points -> [..., h, w] maps (the per-joint loop of vitonhd.py:277-287)."""
    k = np.asarray(keypoints, dtype=np.float64).reshape(-1, 2)
    maps = torch.stack([kpoint_to_heatmap(p, (h, w), sigma) for p in k])
    return maps.reshape(tuple(np.shape(keypoints)[:-1]) + (h, w))


def numpy_to_uint8(images):
    """numpy_to_pil's arithmetic: (images * 255).round().astype('uint8') on fp32 NHWC images in [0, 1]."""
    return (np.asarray(images, dtype=np.float32) * 255).round().astype("uint8")
