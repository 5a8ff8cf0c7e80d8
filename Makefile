# Builds libb2jpeg.so (sm_100a only) and the CPU oracle. `make` = both.
NVCC ?= /usr/local/cuda/bin/nvcc
CSRC := nvjpeg_imagecompressor_b200/csrc
LIB  := nvjpeg_imagecompressor_b200/libb2jpeg.so
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden \
           --expt-relaxed-constexpr -Xptxas -v
SRCS := $(wildcard $(CSRC)/*.cu) $(wildcard $(CSRC)/*.cpp)
OBJS := $(patsubst $(CSRC)/%,build/%.o,$(SRCS))
HDRS := $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/b2jpeg.h

all: $(LIB) oracle

build/%.cu.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

build/%.cpp.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) -cudart static -lpthread

oracle:
	$(MAKE) -C oracle -s

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
