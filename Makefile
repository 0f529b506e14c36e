# Builds the C-ABI library (hand-written sm_100a CUDA) and the C oracle pieces.
# `python -c "import __graft_entry__ as g; g.build()"` drives this.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo --use_fast_math -Xcompiler -fPIC,-Wall,-Wno-unused-function \
             -Xptxas -v --cudart static -Iinclude
CSRC      := xpretrain_b200/csrc
LIBDIR    := xpretrain_b200/lib
OBJDIR    := build/obj
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
LIB       := $(LIBDIR)/libxpretrain_b200.so

all: $(LIB)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.inc) $(wildcard $(CSRC)/*.h) include/xpretrain_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)
	@grep -E "error|warning|spill" $(OBJDIR)/$*.ptxas.log | grep -v "0 bytes spill" | head -20 || true

# retrieval.cu feeds integer rank logic from float comparisons: IEEE exp / division and denormals (no flush-to-zero), so
# that tiny DSL weights stay distinct exactly as in the reference's numpy code
$(OBJDIR)/retrieval.o: NVFLAGS := $(filter-out --use_fast_math,$(NVFLAGS))

$(LIB): $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared --cudart static -o $@ $(OBJS)

clean:
	rm -rf build $(LIB)

# per-kernel SASS mnemonic counts (UTCHMMA / UTMALDG / LDTM / STTM / HMMA ...) -> profiles/r02_sass.md
sass: $(LIB)
	python tools/sass_census.py $(LIB) > profiles/r02_sass.md

.PHONY: all clean sass

# standalone GPU self-tests (no torch); run on the GPU box
TOOLS := build/gemm_selftest
tools: $(LIB) $(TOOLS)
build/gemm_selftest: tools/gemm_selftest.cu $(LIB)
	$(NVCC) $(ARCH) -O2 -std=c++17 --cudart static -Iinclude $< -o $@ -L$(LIBDIR) -lxpretrain_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../$(LIBDIR)'
