"""csrc/hostpipe.cpp (pageable-memory staging: copy threads + non-temporal stores) as a host-only C++ unit test."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_copy_pool_matches_memcpy(tmp_path):
    exe = str(tmp_path / "hostpipe_test")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I" + os.path.join(cuda, "include"), "-o", exe,
                           os.path.join(ROOT, "tests/cpp/hostpipe_test.cpp"),
                           os.path.join(ROOT, "nvjpeg_imagecompressor_b200/csrc/hostpipe.cpp"),
                           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + os.path.join(cuda, "lib64")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-2000:]
