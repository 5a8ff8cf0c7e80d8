"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol include/b2jpeg.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b2jpeg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"B2J_API\s+[^;(]*?\b(b2j_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for s in ("b2j_create", "b2j_destroy", "b2j_encode", "b2j_encode_device", "b2j_decode", "b2j_decode_device", "b2j_decode_finish", "b2j_diff", "b2j_psnr",
              "b2j_secondary", "b2j_strip_phase1", "b2j_strip_phase2", "b2j_strip_phase3", "b2j_strip_phase3_dev"):
        assert s in syms


def test_library_builds_and_exports_all_symbols():
    from nvjpeg_imagecompressor_b200 import _native as N
    path = N.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    syms = header_symbols()
    assert sorted(N.SYMBOLS) == syms, "python binding list and header disagree"
    for s in syms:
        assert hasattr(L, s), s
    assert b"sm_100a" in N.lib().b2j_version()


def test_peek_runs_without_gpu(golden):
    """b2j_peek is pure host code (header parse): usable on the CPU box."""
    import numpy as np
    from nvjpeg_imagecompressor_b200 import Engine
    for f in golden["files"]:
        jpg = np.fromfile(os.path.join(ROOT, "tests", "golden", f["file"]), np.uint8)
        assert Engine.peek(jpg) == (f["W"], f["H"], f["css"])
    with pytest.raises(RuntimeError):
        Engine.peek(np.zeros(100, np.uint8))


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nvjpeg_imagecompressor_b200 import B2JError, Engine
    with pytest.raises(B2JError):
        Engine(64, 64)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nvjpeg_imagecompressor_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "jpeg_oracle" not in txt, f
