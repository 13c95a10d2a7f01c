"""The C-ABI shared library loads and exports every symbol include/mdiff.h declares (no compute, no GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mdiff.h")).read()
    return sorted(set(re.findall(r"MD_API\s+[\w\s\*]+?\b(md_\w+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for s in ("md_create", "md_load_weights", "md_bind_sample", "md_denoise_step", "md_unet_forward",
              "md_spatial_volume", "md_frustum_feats", "md_voxelize", "md_op_conv_gemm", "md_last_error"):
        assert s in syms


def test_library_exports_all_declared_symbols(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.md_version() >= 100


def test_native_module_fails_loudly_without_library(tmp_path, monkeypatch):
    import importlib
    import morphablediffusion_b200._native as nat
    monkeypatch.setattr(nat, "LIB_PATH", tmp_path / "nope.so")
    try:
        nat._load()
        assert False, "expected MdiffError"
    except nat.MdiffError as e:
        assert "no CPU" in str(e) or "missing" in str(e)
    importlib.reload(nat)


def test_engine_refuses_cpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from morphablediffusion_b200 import _native as nat
    from morphablediffusion_b200.engine import Engine
    with pytest.raises(nat.MdiffError):
        Engine()
