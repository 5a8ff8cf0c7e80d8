"""The forward quantiser of k_fdct runs on the FMA pipes (one FADD + one FFMA per coefficient instead of jcdctmgr.c's
integer division). Its identity with the integer form is proven here exhaustively on the CPU for every divisor and
|c| <= 2^18, with the same IEEE operations (tests/cpp/quant_exhaustive.c). Host only."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fma_quantiser_equals_integer_division(tmp_path):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not found")
    exe = str(tmp_path / "quant_exhaustive")
    subprocess.check_call([gcc, "-O2", "-ffp-contract=off", os.path.join(ROOT, "tests/cpp/quant_exhaustive.c"), "-o", exe, "-lm"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "bad 0" in r.stdout, r.stdout[-800:]
