"""CPU: the C-ABI library loads and exports every symbol include/*.h declares (no compute)."""
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ditherpie_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dp_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ("dp_palette_create", "dp_threshold_dither", "dp_halftone", "dp_error_diffusion",
                 "dp_ostromoukhov", "dp_hybrid", "dp_perceptual", "dp_adaptive_variance", "dp_resample_nearest", "dp_kmeans_accumulate",
                 "dp_kmeans_update", "dp_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from dither_pie_b200 import _capi
    from dither_pie_b200.build import build
    build()
    lib = _capi.load_library()
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    # and the binding covers the whole header
    assert set(declared_symbols()) == set(_capi.PROTOTYPES)
    assert lib.dp_version() >= 100


def test_product_fails_loudly_without_gpu_or_library(monkeypatch, tmp_path):
    import numpy as np
    import dither_pie_b200 as dp
    from dither_pie_b200 import _capi
    # missing library -> DitherPieError, never a silent CPU path
    with pytest.raises(_capi.DitherPieError):
        _capi.load_library(str(tmp_path / "nope.so"))
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        d = dp.ImageDitherer(palette=[(0, 0, 0), (255, 255, 255)])
        with pytest.raises(_capi.DitherPieError):
            d.apply_dithering_array(np.zeros((4, 4, 3), np.uint8))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dither_pie_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "").lower() or f == "synth.py" or \
                    "import oracle" not in src and "from oracle" not in src
                assert "from oracle" not in src and "import oracle" not in src
