"""Host-side pieces either side of the hot path (SURVEY.md section 8f row N4 and the a13 sharding), checked on the
CPU against fixtures produced by the unmodified reference (tests/golden/make_golden.py --input-pipeline):
``get_images4`` band split + training-mean subtraction (F8_IMAGES4.py:11-95), ``CrossVal`` (F6_CROSSVAL.py:5-37),
the 18-line config of F2_MAIN.py:61-83, and the micro-batch partition of the data-parallel step."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

DROPIN = os.path.join(ROOT, "corrifnet-correlation-aware-interactive-fusion-multimodal-learning-for-multispectral-images_b200", "dropin")
sys.path.insert(0, os.path.join(GOLDEN))


@pytest.fixture()
def dropin_path():
    sys.path.insert(0, DROPIN)
    names = ("F8_IMAGES4", "F6_CROSSVAL", "F2_MAIN", "F4_TRAIN", "F7_TEST2", "F3_DATASET", "mmvit4")
    for n in names:
        sys.modules.pop(n, None)
    yield
    sys.path.remove(DROPIN)
    for n in names:
        sys.modules.pop(n, None)


def test_get_images4_matches_reference_bit_for_bit(dropin_path, tmp_path, monkeypatch):
    from synth_dstl import N_TILES, TRIND, write_synthetic_dstl
    import F8_IMAGES4
    g = np.load(os.path.join(GOLDEN, "input_pipeline.npz"))
    root = str(tmp_path / "DSTL")
    write_synthetic_dstl(root)
    monkeypatch.setenv("CORRIF_DSTL_ROOT", root)
    real = os.listdir
    monkeypatch.setattr(os, "listdir", lambda p=".": sorted(real(p)))        # the fixture pinned the same order
    images, masks, r, gg, b = F8_IMAGES4.get_images4(N_TILES, 1, 5, None, TRIND, None, "x")
    assert tuple(images.shape) == tuple(g["im/shape"]) == (N_TILES, 3, 3, 224, 224)
    assert tuple(masks.shape) == tuple(g["im/mask_shape"]) == (N_TILES, 3, 1, 224, 224)
    assert images.dtype == torch.float32 and masks.dtype == torch.float32
    assert np.array_equal(np.array([r, gg, b], np.float32), g["im/means_rgb"])            # bit-exact means
    assert np.array_equal(images.reshape(-1)[::9973].numpy(), g["im/sample"])             # bit-exact tiles
    assert images.double().sum().item() == float(g["im/sum64"])
    assert np.array_equal(masks.reshape(-1)[::9973].numpy(), g["im/mask_sample"])
    assert masks.sum().item() == float(g["im/mask_sum"])
    # band groups: modality 1 = bands 9-11, modality 2 = bands 12-14 of the 20-band cube, each zero-mean on TRIND
    assert np.allclose(images[TRIND].double().mean(dim=(0, 3, 4)).numpy(), 0.0, atol=1e-3)
    assert torch.equal(masks[:, 0], masks[:, 1]) and torch.equal(masks[:, 0], masks[:, 2])


def test_crossval_matches_reference_folds(dropin_path, tmp_path, monkeypatch):
    import F6_CROSSVAL
    g = np.load(os.path.join(GOLDEN, "input_pipeline.npz"))
    with open(tmp_path / "randInd5985.txt", "w") as f:
        f.write("\n".join(str(int(v)) for v in g["cv/perm"]) + "\n")
    monkeypatch.chdir(tmp_path)
    for fno in (1, 2, 5):
        ts, tr, vl = F6_CROSSVAL.CrossVal(5985, fno, 5)
        assert np.array_equal(ts, g[f"cv/f{fno}/ts"]) and np.array_equal(tr, g[f"cv/f{fno}/tr"])
        assert np.array_equal(vl, g[f"cv/f{fno}/vl"])
        assert (len(tr), len(vl), len(ts)) == (4310, 478, 1197)                 # trind.txt / vlind.txt / tsind.txt
        assert len(set(ts) | set(tr) | set(vl)) == 5985


def test_f2_main_reads_the_18_line_config(dropin_path, tmp_path):
    import F2_MAIN
    lines = ["40", "1", "5", "0.1", "8", "2", "0.0001", "Adam", "BCEWithLogitsLoss", "BCEWithLogitsLoss", "Jaccard",
             "kaiming_normal_", "5", "0.9", "224", "MMVit4", "x", "notr"]
    p = tmp_path / "model0.txt"
    p.write_text("\n".join(lines) + "\n")
    cfg = F2_MAIN.read_config(str(p))
    assert cfg["trainSetSize"] == 40 and cfg["miniBatchSize"] == 8 and cfg["learnRate"] == 1e-4
    assert cfg["optimizerType"] == "Adam" and cfg["modeltype"] == "MMVit4" and cfg["lim"] == 224
    assert [n for n, _ in F2_MAIN.CONFIG_FIELDS][:3] == ["trainSetSize", "fno", "fsiz"] and len(F2_MAIN.CONFIG_FIELDS) == 18
    p.write_text("\n".join(lines[:10]))
    with pytest.raises(ValueError):
        F2_MAIN.read_config(str(p))
    im, ma = F2_MAIN.synthetic_tiles(3, 32, 0)
    assert im.shape == (3, 3, 3, 32, 32) and ma.shape == (3, 3, 1, 32, 32) and set(ma.unique().tolist()) <= {0.0, 1.0}


def test_shard_micro_batches_partition():
    from corrif_b200.train import shard_micro_batches
    # BASELINE configs[2]: 8 micro-batches per step; rank r takes r, r+G, ...
    for world in (1, 2, 4, 8):
        seen = []
        for rank in range(world):
            plan = shard_micro_batches(24, 8, rank, world)
            assert len(plan) == 3 and all(total == 8 for _, total in plan)
            assert plan[0][0] == list(range(rank, 8, world))
            seen += [j for mine, _ in plan for j in mine]
        assert sorted(seen) == list(range(24))                       # every micro-batch exactly once
    # ragged epoch: 11 batches, groups of 4 on 4 ranks -> last group has 3, rank 3 gets none
    plans = [shard_micro_batches(11, 4, r, 4) for r in range(4)]
    assert [p[-1] for p in plans] == [([8], 3), ([9], 3), ([10], 3), ([], 3)]
    with pytest.raises(ValueError):
        shard_micro_batches(10, 3, 0, 2)


def test_sequential_loader_is_resliced_by_index(dropin_path):
    import F4_TRAIN
    from F3_DATASET import satellitedata
    images = torch.arange(10.0).view(10, 1, 1, 1, 1).repeat(1, 3, 3, 2, 2)
    masks = torch.zeros(10, 3, 1, 2, 2)
    dl = torch.utils.data.DataLoader(satellitedata(images, masks), batch_size=4, shuffle=False)
    n, fetch = F4_TRAIN._batches(dl)
    assert n == 3
    ref = list(dl)
    for i in range(3):
        im, ma = fetch(i)
        assert torch.equal(im, ref[i][0]) and torch.equal(ma, ref[i][1])
    n2, fetch2 = F4_TRAIN._batches(ref)                               # any other iterable: walked once
    assert n2 == 3 and torch.equal(fetch2(2)[0], ref[2][0])
