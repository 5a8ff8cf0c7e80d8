"""The product's JPEG header parser (csrc/dec_parse.cpp) under AddressSanitizer: truncated, corrupted and spliced
streams (from golden files and checker-written ones, with and without restart markers) must be rejected or parsed
with in-range results, never read out of bounds. Host only."""
import glob
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parser_survives_mutated_streams(tmp_path, oracle):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "parse_fuzz")
    r = subprocess.run([nvcc, "-x", "cu", "-std=c++17", "-O1", "-g", "-arch=sm_100a", "--expt-relaxed-constexpr",
                        "-Xcompiler", "-fsanitize=address,-fno-omit-frame-pointer",
                        os.path.join(ROOT, "tests/cpp/parse_fuzz.cpp"),
                        os.path.join(ROOT, "nvjpeg_imagecompressor_b200/csrc/dec_parse.cpp"), "-o", exe, "-Xlinker", "-lasan"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cannot build with AddressSanitizer here: " + r.stderr[-300:])
    seeds = sorted(glob.glob(os.path.join(ROOT, "tests/golden/*.jpg")))
    for i, (W, H, css, q, opt, ri) in enumerate(((50, 70, 1, 95, 1, 0), (64, 48, 3, 90, 1, 5), (8, 8, 0, 100, 1, 0), (135, 121, 2, 60, 0, 3))):
        p = str(tmp_path / f"seed{i}.jpg")
        oracle.encode(oracle.synth(W, H, 1, 8), css, q, opt, ri).tofile(p)
        seeds.append(p)
    env = dict(os.environ, ASAN_OPTIONS="protect_shadow_gap=0:detect_leaks=0")
    r = subprocess.run([exe, "60000"] + seeds, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "accepted" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
