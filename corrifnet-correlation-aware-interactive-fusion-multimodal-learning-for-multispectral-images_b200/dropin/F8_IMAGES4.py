"""Drop-in for the reference's F8_IMAGES4.py (``get_images4``, F8_IMAGES4.py:11-95): the DSTL input pipeline of
CorrIFNet - SURVEY.md section 8f row N4.

Reads ``inputPatch`` from the ``.mat`` files of three directories (RGB tiles, class-06 masks, 20-band tiles),
splits the 20-band cube into the NIR group (bands 9-11) and the SWIR group (bands 12-14), subtracts the
TRAINING-SET mean of every band from all tiles and stacks the three groups as modalities:
images [N, 3 modalities, 3 bands-as-depth, 224, 224] float32, masks [N, 3, 1, 224, 224] (the mask replicated per
modality, :88).  Returns (images, masks, trMeanR, trMeanG, trMeanB) like the reference.

The reference hard-codes ``C:/Users/Public/Server/data/DSTL`` (:20-32); here the root comes from
``CORRIF_DSTL_ROOT`` (default: the reference's literal path, so an unmodified setup keeps working).  File order
is ``os.listdir`` order, as in the reference.  This is one-time host work (numpy, like the reference); per step the
tensors it returns travel to the GPU through corrif_b200.staging.PinnedPipeline (F4_TRAIN.py)."""
import os

import numpy as np
import scipy.io as sio
import torch

LIM = 224
NIR_BANDS, SWIR_BANDS = (9, 10, 11), (12, 13, 14)


def _root():
    return os.environ.get("CORRIF_DSTL_ROOT", "C:/Users/Public/Server/data/DSTL")


def _load_dir(sub, names, **kw):
    return np.asarray([sio.loadmat(os.path.join(_root(), sub, n), **kw)["inputPatch"] for n in names], dtype=np.float32)


def get_images4(trainSetSize, fno, fsiz, tsind, trind, vlind, chindex):
    root = _root()
    names1 = os.listdir(os.path.join(root, "RGBs"))[0:trainSetSize]
    rgb = _load_dir("RGBs", names1)                                             # [N,224,224,3]
    masks = _load_dir("class06_mats", names1)                                   # [N,224,224]
    names2 = os.listdir(os.path.join(root, "all20Ch"))[0:trainSetSize]
    cube = _load_dir("all20Ch", names2, verify_compressed_data_integrity=False)  # [N,224,224,20]
    n = trainSetSize
    groups = [rgb.reshape(n, LIM, LIM, 3), cube[..., list(NIR_BANDS)].reshape(n, LIM, LIM, 3),
              cube[..., list(SWIR_BANDS)].reshape(n, LIM, LIM, 3)]
    groups = [np.moveaxis(g, 3, 1) for g in groups]                             # [N,3,224,224] views (:52-57)
    means = np.zeros((3, 3), np.float32)
    for m, g in enumerate(groups):
        for c in range(3):
            # same numpy float32 reduction over the same contiguous gather as the reference (:60-79): bit-identical
            means[m, c] = g[trind, c, :, :].mean()
            g[:, c, :, :] = g[:, c, :, :] - means[m, c]
    images = torch.stack([torch.from_numpy(g) for g in groups], dim=1)          # [N,3,3,224,224] (:87)
    masks_t = torch.from_numpy(masks.reshape(n, 1, LIM, LIM)).unsqueeze(1).repeat(1, 3, 1, 1, 1)   # :88
    print("image size", images.shape, "mask size", masks_t.shape)
    return images, masks_t, means[0, 0], means[0, 1], means[0, 2]
