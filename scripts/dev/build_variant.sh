#!/bin/bash
# build a variant of libb2jpeg.so with extra -D flags into build_var/<name>.so  (development aid)
name=$1; shift
mkdir -p build_var/$name
for f in nvjpeg_imagecompressor_b200/csrc/*.cu nvjpeg_imagecompressor_b200/csrc/*.cpp; do
  b=$(basename $f)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr "$@" -x cu -c $f -o build_var/$name/$b.o 2>/dev/null &
done
wait
/usr/local/cuda/bin/nvcc -shared -o build_var/$name.so build_var/$name/*.o -cudart static -lpthread
