# Builds libgpz_b200.so (sm_100a only) and nothing else.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v
# CUTLASS/CuTe header tree vendored in the image (flashinfer); only i8gemm_cutlass.cu uses it
CUTLASS_DIR := $(shell python -c "import os,site; p=[os.path.join(s,'flashinfer','data','cutlass') for s in site.getsitepackages()]; p=[x for x in p if os.path.isdir(os.path.join(x,'include','cutlass'))]; print(p[0] if p else '')" 2>/dev/null)
ifneq ($(CUTLASS_DIR),)
CUTLASS_FLAGS := -DGPZ_HAVE_CUTLASS --expt-relaxed-constexpr -I$(CUTLASS_DIR)/include -I$(CUTLASS_DIR)/tools/util/include
endif
SRC := $(wildcard gpz_b200/csrc/*.cu)
HDR := $(wildcard gpz_b200/csrc/*.cuh) include/gpz_b200.h
OBJ := $(patsubst gpz_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := gpz_b200/libgpz_b200.so

all: $(LIB)

build/i8gemm_cutlass.o: gpz_b200/csrc/i8gemm_cutlass.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $(CUTLASS_FLAGS) -c $< -o $@ 2> build/i8gemm_cutlass.log || (cat build/i8gemm_cutlass.log; false)

build/%.o: gpz_b200/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -ldl

clean:
	rm -rf build $(LIB)
.PHONY: all clean
