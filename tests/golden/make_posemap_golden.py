"""Writes tests/golden/posemap.npz from the reference's OWN src/utils/posemap.py, imported unmodified (it only needs numpy + torch), run
in the build container where /root/reference exists:

    python tests/golden/make_posemap_golden.py
"""
import importlib.util
import os

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "posemap.npz")
spec = importlib.util.spec_from_file_location("ref_posemap", "/root/reference/src/utils/posemap.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

H, W, SIGMA = 64, 48, 9
rng = np.random.default_rng(1234)
kp = np.concatenate([
    rng.uniform(0, [W, H], size=(10, 2)),                       # inside the map, fractional
    np.array([[0.0, 0.0], [-3.0, -1.0], [0.0, 17.5], [12.25, 0.0],  # "missing" (no coordinate > 0) and half-missing key-points
              [W + 6.0, 10.0], [5.0, H + 20.0], [-4.0, 30.0],       # outside the map on one side
              [W - 1.0, H - 1.0], [0.4, 0.4], [23.5, 31.5]]),       # corners / exact half-pixel ties
]).astype(np.float64)
maps = np.stack([ref.kpoint_to_heatmap(p, (H, W), SIGMA).numpy() for p in kp])
np.savez_compressed(OUT, keypoints=kp, maps=maps, sigma=SIGMA)
print(maps.shape, maps.dtype, float(maps.max()), "->", OUT)
