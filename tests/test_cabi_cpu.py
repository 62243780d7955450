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


def _conv_desc(chans, cout, B, D, H, W, ksize=3, pad=1):
    from corrif_b200 import _lib
    d = _lib.Conv3dDesc()
    for i, c in enumerate(chans):
        d.src[i].p, d.src[i].C, d.src[i].ld = 0x10000 * (i + 1), c, c      # never dereferenced by the planners
    d.nsrc, d.B, d.D, d.H, d.W, d.Cin, d.Cout = len(chans), B, D, H, W, sum(chans), cout
    d.ksize, d.pad_mode = ksize, pad
    return d


def test_line_convolution_planner_accepts_the_decoder_shapes_and_rejects_the_rest(lib):
    """corrif_conv3d_tc_supported / _pack_floats are host-only: which shapes the tcgen05 line convolution takes
     can be checked without a GPU."""
    ok = lambda *a, **k: bool(lib.corrif_conv3d_tc_supported(ctypes.byref(_conv_desc(*a, **k))))   # noqa: E731
    # the decoder at micro-batch 8: forward shapes ...
    assert ok((24, 8), 8, 8, 128, 128, 128) and ok((16,), 8, 8, 128, 128, 128)
    assert ok((48, 16), 16, 8, 64, 64, 64) and ok((32,), 16, 8, 64, 64, 64) and ok((64,), 32, 8, 32, 32, 32)
    # ... and data-gradient shapes (channel roles swapped), incl. the ones whose weights are staged chunk by chunk
    assert ok((8,), 32, 8, 128, 128, 128, pad=2) and ok((32,), 128, 8, 32, 32, 32, pad=2) and ok((64,), 320, 8, 16, 16, 16, pad=2)
    assert ok((24,), 24, 8, 3, 64, 64, pad=0)                           # RFM1 3x3x3, zero padding
    # weights that do not fit in shared memory, even one output chunk at a time
    assert not ok((96, 32), 32, 8, 32, 32, 32) and not ok((128,), 128, 8, 16, 16, 16) and not ok((192,), 192, 8, 8, 8, 8)
    # geometry: line width must be 16 / 32 / 64 / 128 and the batch a multiple of 128 / W
    assert not ok((32,), 8, 1, 8, 16, 10) and not ok((32,), 8, 2, 8, 16, 16) and ok((32,), 8, 8, 8, 16, 16)
    assert not ok((32,), 8, 8, 16, 16, 16, ksize=1)
    assert not ok((12,), 8, 1, 4, 8, 128)                               # source channels must be a multiple of 8
    # the packed operand: 3 y-taps x [3 * NPAD rows x 128-byte rows] per 32-channel chunk, padded to 1 KB
    n = lib.corrif_conv3d_tc_pack_floats(ctypes.byref(_conv_desc((32,), 8, 1, 8, 8, 128)))
    assert n == 3 * (3 * 32 * 128) // 4
    assert lib.corrif_conv3d_tc_pack_floats(ctypes.byref(_conv_desc((12,), 8, 1, 4, 8, 128))) == 0


def test_tensor_core_weight_gradient_planner(lib):
    ok = lambda *a, **k: bool(lib.corrif_conv3d_wgrad_tc_supported(ctypes.byref(_conv_desc(*a, **k))))   # noqa: E731
    assert ok((24, 8), 8, 8, 128, 128, 128) and ok((16,), 8, 8, 128, 128, 128)
    assert ok((48, 16), 16, 8, 64, 64, 64) and ok((32,), 16, 8, 64, 64, 64)
    assert not ok((64,), 32, 8, 32, 32, 32)                             # lines shorter than 64 voxels, 32 output channels
    assert not ok((96,), 8, 1, 8, 8, 128) and not ok((32,), 8, 1, 8, 7, 128) and not ok((32,), 8, 1, 8, 8, 128, ksize=1)
    rc = lib.corrif_conv3d_wgrad_tc(ctypes.byref(_conv_desc((64,), 32, 8, 32, 32, 32)), 0x10000, 32, 0x20000, None)
    assert rc == -1 and b"not supported" in lib.corrif_last_error()
