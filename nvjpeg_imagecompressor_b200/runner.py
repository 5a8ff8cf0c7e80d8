"""NvjpegCompressRunner -- Python mirror of the reference's public facade (src/ImageCompressorDll/ImageCompressor.h:22-42,
ImageCompressor.cpp:14-101): same method names, argument meaning and failure convention (empty result + run_state 0),
numpy arrays standing in for cv::Mat / std::vector<uchar>. The C++ facade with the exact signatures is
include/ImageCompressor.h."""
import time

import numpy as np

from .engine import B2JError, Engine


class NvjpegCompressRunner:
    def __init__(self, width=8320, height=40000, quality=95, optimize=True, css="422", device=-1, verbose=True):
        self._args = dict(width=width, height=height, quality=quality, optimize=optimize, css=css, device=device)
        self._enc = None
        self._dec = None
        self._verbose = verbose

    # ImageCompressor.cpp:19-37
    def buildCompressEnv(self):
        if self._enc is None:
            self._enc = Engine(**self._args)

    def buildDecodeEnv(self):
        if self._dec is None:
            self._dec = self._enc if self._enc is not None else Engine(**self._args)

    def deleteCompressEnv(self):
        if self._enc is not None and self._enc is not self._dec:
            self._enc.close()
        self._enc = None

    def deleteDecodeEnv(self):
        if self._dec is not None and self._dec is not self._enc:
            self._dec.close()
        self._dec = None

    def _log(self, what, t0):
        if self._verbose:
            print(f"[INFO] NvjpegCompressRunner {what} Func Cost Time : {int((time.perf_counter() - t0) * 1000)} ms")

    # ImageCompressor.cpp:45-61 -- returns (obuffer, run_state)
    def compress(self, image):
        t0 = time.perf_counter()
        try:
            if self._enc is None:
                raise B2JError(-1, "buildCompressEnv() was not called")
            out = self._enc.encode(image)
        except (B2JError, AssertionError) as e:
            if self._verbose:
                print(f"[ERROR] {e}")
            out = np.empty(0, np.uint8)
        self._log("Compress", t0)
        return out, (0 if out.size == 0 else 1)

    # ImageCompressor.cpp:63-88 -- returns (image, run_state)
    def decode(self, image_path):
        try:
            jpg = np.fromfile(image_path, np.uint8)
        except OSError:
            print("Failed to open JPEG file.")
            return np.empty((0, 0, 3), np.uint8), 0
        t0 = time.perf_counter()
        try:
            if self._dec is None:
                raise B2JError(-1, "buildDecodeEnv() was not called")
            img = self._dec.decode(jpg)
        except B2JError as e:
            if self._verbose:
                print(f"[ERROR] {e}")
            img = np.empty((0, 0, 3), np.uint8)
        self._log("Decode", t0)
        return img, (0 if img.size == 0 else 1)

    # ImageCompressor.cpp:90-101
    def save(self, save_path, obuffer):
        try:
            np.asarray(obuffer, np.uint8).tofile(save_path)
        except OSError as e:
            print(f"Exception caught: {e}")

    # README.md:8 entry points
    def reconstruct(self, obuffer):
        eng = self._dec or self._enc
        return eng.decode(obuffer)

    def difference_map(self, a, b, mode="absdiff"):
        eng = self._enc or self._dec
        return eng.diff(a, b, 0 if mode == "absdiff" else 1)

    def psnr(self, a, b):
        eng = self._enc or self._dec
        return eng.psnr(a, b)[0]

    def secondary_compress(self, image, mode="offset128"):
        """-> (jpeg, jpeg_of_difference_map, reconstruction, psnr)"""
        return self._enc.secondary(image, 0 if mode == "absdiff" else 1)
