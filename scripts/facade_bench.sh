#!/bin/bash
# builds and runs tests/cpp/facade_bench.cpp (the C++ facade at the headline size on pageable memory); prints one JSON line
set -e
cd "$(dirname "$0")/.."
g++ -std=c++17 -O2 -DB2J_USE_CV_SHIM -o /tmp/facade_bench tests/cpp/facade_bench.cpp facade/ImageCompressor.cpp \
    -Lnvjpeg_imagecompressor_b200 -lb2jpeg -Wl,-rpath,$PWD/nvjpeg_imagecompressor_b200
B2J_QUIET=1 /tmp/facade_bench "$@"
