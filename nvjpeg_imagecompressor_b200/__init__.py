"""nvjpeg_imagecompressor_b200 -- B200-native JPEG encode/decode engine behind the Nvjpeg-ImageCompressor interface.

csrc/       hand-written sm_100a kernels + the extern "C" boundary (include/b2jpeg.h) -> libb2jpeg.so
engine.py   ctypes owner of one b2j_ctx
runner.py   NvjpegCompressRunner, the reference's facade (ImageCompressor.h) in Python
strips.py   multi-GPU MCU-row strip encoder over torch.distributed
"""
from ._native import CSS, build, lib  # noqa: F401
from .engine import B2JError, Engine, MultiEngine  # noqa: F401
from .runner import NvjpegCompressRunner  # noqa: F401
