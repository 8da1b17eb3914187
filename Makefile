# Builds libgpz_b200.so (sm_100a only) and nothing else.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v
SRC := $(wildcard gpz_b200/csrc/*.cu)
HDR := $(wildcard gpz_b200/csrc/*.cuh) include/gpz_b200.h
OBJ := $(patsubst gpz_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := gpz_b200/libgpz_b200.so

HARNESS := tests/mex_stub/libgpz_mex_harness.so

all: $(LIB) $(HARNESS)

# test infrastructure: the MEX gateway linked against a minimal libmx mock, so that it can be executed without MATLAB
$(HARNESS): matlab/gpz_b200_mex.cpp tests/mex_stub/mex_mock.cpp tests/mex_stub/mex.h include/gpz_b200.h $(LIB)
	g++ -std=c++17 -O1 -fPIC -shared -Wall -Iinclude -Itests/mex_stub matlab/gpz_b200_mex.cpp tests/mex_stub/mex_mock.cpp \
	    -Lgpz_b200 -lgpz_b200 -Wl,-rpath,'$$ORIGIN/../../gpz_b200' -o $@

build/%.o: gpz_b200/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -ldl -lpthread

clean:
	rm -rf build $(LIB) $(HARNESS)
.PHONY: all clean
