# Builds the C-ABI CUDA library (sm_100a only) and the oracle helpers.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
SRC := $(wildcard henbun_b200/csrc/*.cu)
OBJ := $(patsubst henbun_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := henbun_b200/libhenbun_b200.so

all: $(LIB)

build/%.o: henbun_b200/csrc/%.cu $(wildcard henbun_b200/csrc/*.cuh) include/henbun_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart -ldl

clean:
	rm -rf build $(LIB)
.PHONY: all clean
