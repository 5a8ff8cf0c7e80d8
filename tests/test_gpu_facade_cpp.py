"""The C++ drop-in facade (include/ImageCompressor.h + facade/ImageCompressor.cpp) driven by a port of the reference
demo's call order (tests/cpp/demo_main.cpp <- src/ImageCompressor/main.cpp), compiled with g++ against the cv::Mat shim."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp):
    exe = os.path.join(tmp, "demo")
    libdir = os.path.join(ROOT, "nvjpeg_imagecompressor_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-DB2J_USE_CV_SHIM", "-o", exe, os.path.join(ROOT, "tests/cpp/demo_main.cpp"),
                           os.path.join(ROOT, "facade/ImageCompressor.cpp"), "-L" + libdir, "-lb2jpeg", "-Wl,-rpath," + libdir])
    return exe


def test_facade_compiles_without_gpu(tmp_path):
    from nvjpeg_imagecompressor_b200 import _native as N
    N.build()
    assert os.path.exists(_build(str(tmp_path)))


@pytest.mark.gpu
@pytest.mark.parametrize("ngpus", [1, 2])
def test_cpp_demo_sequence(tmp_path, oracle, ngpus):
    """ngpus = 2: the same demo with compress() spread over two contexts of the one process (b2j_multi_*; both on this
    GPU here, B2J_MULTI_DEVICES=0,0 -- on a multi-GPU box the facade takes devices 0 .. ngpus-1)."""
    from nvjpeg_imagecompressor_b200 import _native as N
    N.lib()
    exe = _build(str(tmp_path))
    W, H = 640, 360
    a, b = oracle.synth(W, H, 1, 8), oracle.synth(W, H, 2, 8)
    a.tofile(tmp_path / "a.bgr")
    b.tofile(tmp_path / "b.bgr")
    env = dict(os.environ, B2J_MULTI_DEVICES="0,0")
    r = subprocess.run([exe, str(W), str(H), str(tmp_path / "a.bgr"), str(tmp_path / "b.bgr"), str(tmp_path), str(ngpus)],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("[INFO] Successful.") == 4
    for name, img in (("1", a), ("2", b)):
        jpg = np.fromfile(tmp_path / f"{name}.jpeg", np.uint8)
        want = oracle.encode(img, 1, 95, 1)
        assert np.array_equal(jpg, want)
        dec = np.fromfile(tmp_path / f"{name}.dec.bgr", np.uint8).reshape(H, W, 3)
        assert np.array_equal(dec, oracle.decode(want))
