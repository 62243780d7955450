"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/corrif.h declares; the ctypes prototypes cover the header; argument errors are
reported through the status code (no compute needs a GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "corrif.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(corrif_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import corrif_b200
    corrif_b200.build_library()
    from corrif_b200 import _lib
    return _lib.load()


def test_header_declares_expected_surface():
    syms = header_symbols()
    for must in ("corrif_gemm", "corrif_inter_corr_fwd", "corrif_inter_corr_bwd", "corrif_layernorm_fwd",
                 "corrif_layernorm_bwd", "corrif_softmax_fwd", "corrif_jaccard_sums", "corrif_loss_jaccard_fused", "corrif_dropout_colsum",
                 "corrif_confusion_counts", "corrif_adam_step"):
        assert must in syms


def test_library_exports_every_header_symbol(lib):
    for s in header_symbols():
        assert hasattr(lib, s), f"libcorrif_b200.so does not export {s}"


def test_ctypes_prototypes_cover_header(lib):
    from corrif_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == header_symbols()
    assert lib.corrif_abi_version() == 3


def test_gemm_desc_layout_matches_c_struct(lib):
    from corrif_b200 import _lib
    assert ctypes.sizeof(_lib.GemmDesc) == lib.corrif_sizeof_gemm_desc()
    assert _lib.GemmDesc.M.offset == 88 and _lib.GemmDesc.a_bo.offset == 120
    assert _lib.GemmDesc.alpha.offset == 180 and _lib.GemmDesc.drop_seed.offset % 8 == 0


def test_argument_errors_are_reported_without_a_gpu(lib):
    from corrif_b200 import _lib
    rc = lib.corrif_gemm(None, None)
    assert rc == -1 and b"null descriptor" in lib.corrif_last_error()
    rc = lib.corrif_layernorm_fwd(1, None, 0, 1, 1, None, 1, 1, 1, 8, 256, 0, None)
    assert rc == -1 and b"C must be 512" in lib.corrif_last_error()
    rc = lib.corrif_inter_corr_fwd(1, 1, 1, 7, 2, 8, 8, None)
    assert rc == -1 and b"M must be 2..6" in lib.corrif_last_error()
    with pytest.raises(_lib.CorrifError):
        _lib.check(rc, "inter_corr")


def test_no_cpu_fallback_in_ops():
    import torch
    from corrif_b200 import metrics
    with pytest.raises(ValueError):
        metrics.Jaccard2(torch.zeros(8, 1), torch.zeros(8, 1))


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "corrifnet-correlation-aware-interactive-fusion-multimodal-learning-for-multispectral-images_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or f == "README.md", (dirpath, f)
