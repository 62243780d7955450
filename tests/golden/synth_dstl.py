"""Synthetic DSTL-shaped .mat tree for the input-pipeline parity test: the same
seeded files are written for the reference run that produced tests/golden/input_pipeline.npz
(make_golden.make_input_pipeline_golden) and for the drop-in under test."""
import os

import numpy as np
import scipy.io as sio

N_TILES = 6
TRIND = np.array([0, 2, 3, 5])
LIM = 224


def write_synthetic_dstl(root, seed=42):
    rng = np.random.default_rng(seed)
    for sub in ("RGBs", "class06_mats", "all20Ch"):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    for t in range(N_TILES):
        name = "tile_%03d.mat" % t
        rgb = rng.integers(0, 2048, (LIM, LIM, 3)).astype(np.float64)          # DSTL digital numbers
        cube = rng.integers(0, 4096, (LIM, LIM, 20)).astype(np.float64)
        mask = (rng.random((LIM, LIM)) < 0.3).astype(np.float64)
        sio.savemat(os.path.join(root, "RGBs", name), {"inputPatch": rgb})
        sio.savemat(os.path.join(root, "class06_mats", name), {"inputPatch": mask})
        sio.savemat(os.path.join(root, "all20Ch", name), {"inputPatch": cube})
