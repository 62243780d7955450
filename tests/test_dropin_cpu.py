"""Drop-in boundary checks that need no GPU: the replacement ``mmvit4.MMVit4`` owns exactly the
reference's 1140 state_dict entries (names and shapes) and refuses to compute on the CPU."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "corrifnet-correlation-aware-interactive-fusion-multimodal-learning-for-multispectral-images_b200", "dropin")
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def dropin_mmvit4():
    sys.path.insert(0, DROPIN)
    try:
        sys.modules.pop("mmvit4", None)
        import mmvit4
        yield mmvit4
    finally:
        sys.path.remove(DROPIN)
        sys.modules.pop("mmvit4", None)


@pytest.fixture(scope="module")
def model(dropin_mmvit4):
    torch.manual_seed(0)
    return dropin_mmvit4.MMVit4(num_cls=1)


def test_state_dict_inventory_equals_reference(model):
    inv = json.load(open(os.path.join(GOLDEN, "mmvit4_state_dict_inventory.json")))
    sd = model.state_dict()
    assert len(sd) == len(inv) == 1140
    assert set(sd) == set(inv)
    for k, v in sd.items():
        assert list(v.shape) == inv[k], k
    assert sum(p.numel() for p in model.parameters()) == 85_479_624


def test_strict_load_of_reference_shaped_state(model):
    from oracle import corrif_oracle as O
    inv = json.load(open(os.path.join(GOLDEN, "mmvit4_state_dict_inventory.json")))
    missing, unexpected = model.load_state_dict(O.make_full_model_state(1, inv), strict=True)
    assert not missing and not unexpected


def test_fusion_parameters_are_the_hot_path_inventory(model):
    from oracle import corrif_oracle as O
    ps = model.fusion_parameters()
    assert sum(p.numel() for p in ps) == 10_310_336
    named = dict(model.named_parameters())
    assert [tuple(p.shape) for p in ps] == [tuple(O.param_shapes()[n]) for n in model._fusion_names]
    assert all(n in named for n in O.param_shapes())


def test_module_constants_match_reference(dropin_mmvit4):
    m = dropin_mmvit4
    assert (m.basic_dims, m.transformer_basic_dims, m.mlp_dim, m.num_heads, m.depth, m.num_modals,
            m.patch_size) == (8, 512, 512, 8, 1, 3, 8)


def test_forward_on_cpu_raises_no_fallback(model):
    with pytest.raises((RuntimeError, NotImplementedError, ValueError)):
        model(torch.zeros(1, 3, 3, 32, 32))


def test_jaccard_dropins_export_reference_names():
    sys.path.insert(0, DROPIN)
    try:
        for name in ("F5_JACCARD2", "F5_JACCARD", "F3_DATASET"):
            sys.modules.pop(name, None)
        import F5_JACCARD2, F5_JACCARD, F3_DATASET
        assert callable(F5_JACCARD2.Jaccard) and callable(F5_JACCARD2.Jaccard2) and callable(F5_JACCARD2.JaccardAndF1)
        assert callable(F5_JACCARD.Jaccard)
        ds = F3_DATASET.satellitedata(torch.zeros(4, 3, 3, 8, 8), torch.zeros(4, 3, 1, 8, 8))
        assert len(ds) == 4 and ds[1][0].shape == (3, 3, 8, 8) and ds[1][1].shape == (3, 1, 8, 8)
    finally:
        sys.path.remove(DROPIN)
        for name in ("F5_JACCARD2", "F5_JACCARD", "F3_DATASET"):
            sys.modules.pop(name, None)


def test_flat_adam_only_takes_over_a_stock_reference_style_adam():
    """TrainStep swaps optim.step() for the one-kernel FlatAdam only for torch.optim.Adam with the reference's
    settings (F2_MAIN.py:168-169) on CUDA fp32 parameters; anything else keeps its own step()."""
    import torch
    from corrif_b200 import train
    w = [torch.nn.Parameter(torch.zeros(4, 4))]
    assert not train.FlatAdam.eligible(torch.optim.Adam(w, 1e-3))                       # CPU parameters
    assert not train.FlatAdam.eligible(torch.optim.SGD(w, 1e-3))
    assert not train.FlatAdam.eligible(torch.optim.AdamW(w, 1e-3))
    step = train.TrainStep(torch.nn.Linear(2, 2), torch.optim.Adam(w, 1e-3, weight_decay=0.1))
    assert step.flat_adam is None

