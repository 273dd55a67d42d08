"""Device-side pieces of the dataset tensorisation: the pose heat-maps that the reference datasets build on the
host, one numpy exp per joint and sample (/root/reference/src/dataset/vitonhd.py:236-287, src/utils/posemap.py:6-35).  Same function
names and argument meaning; the arithmetic runs in `ladi_pose_heatmaps` (csrc/pointwise.cu).  The label-map / PIL-drawing parts of the
datasets (parse masks, arm lines) are host code in the reference and stay out of scope."""
import torch

from . import ops


def get_coco_body25_mapping():
    """src/utils/posemap.py:37-58: COCO joint index -> BODY_25 joint index (BODY_25's joint 8, MidHip, is skipped)."""
    return {i: (i if i < 8 else i + 1) for i in range(18)}


def kpoint_to_heatmap(kpoint, shape, sigma, device="cuda"):
    """posemap.py:6-35 for ONE key-point: [x, y] -> fp32 [H, W] on `device`."""
    k = torch.as_tensor(kpoint, dtype=torch.float32).reshape(1, 2).to(device)
    return ops.pose_heatmaps(k, int(shape[0]), int(shape[1]), float(sigma))[0]


def pose_map_from_keypoints(pose_keypoints_2d, height=512, width=384, source_size=(1024, 768), sigma=9.0, device="cuda"):
    """vitonhd.py:240-287 for a batch: OpenPose BODY_25 `pose_keypoints_2d` lists ([B, 25*3] or [B, 25, 3], pixel coordinates in the
    `source_size` = (height, width) image) -> `pose_map` fp32 [B, 18, height, width]: drop the confidences, rescale to the working size
    (:245-246), pick the 18 COCO joints (:248-251) and render the Gaussians (:277-287)."""
    p = torch.as_tensor(pose_keypoints_2d, dtype=torch.float32)
    p = p.reshape(p.shape[0], -1, 3)[..., :2].clone()
    p[..., 0] *= width / source_size[1]
    p[..., 1] *= height / source_size[0]
    m = get_coco_body25_mapping()
    sel = p[:, [m[i] for i in range(len(m))]]
    return ops.pose_heatmaps(sel.to(device), height, width, sigma)
