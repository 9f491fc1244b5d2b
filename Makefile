# Build of the B200 (sm_100a) hot-path library, the host tools and the test-only oracle.
#   make            -> videovector_b200/lib/libvv_b200.so  (product: CUDA kernels + C-ABI + host C++)
#   make oracle     -> oracle/_build/libvv_oracle.so        (TEST INFRASTRUCTURE ONLY)
#   make ref        -> oracle/_ref/...                      (reference layer sources, when /root/reference exists)
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Iinclude -Ivideovector_b200/csrc \
             --expt-relaxed-constexpr -Xptxas -v
LIBDIR    := videovector_b200/lib
OBJDIR    := build/obj
CU_SRCS   := $(wildcard videovector_b200/csrc/*.cu) $(wildcard videovector_b200/csrc/host/*.cu)
CPP_SRCS  := $(wildcard videovector_b200/csrc/host/*.cpp) $(wildcard videovector_b200/csrc/host/caffe_compat/*.cpp)
CU_OBJS   := $(patsubst %.cu,$(OBJDIR)/%.o,$(CU_SRCS))
CPP_OBJS  := $(patsubst %.cpp,$(OBJDIR)/%.o,$(CPP_SRCS))
HDRS      := include/vv_b200.h $(wildcard videovector_b200/csrc/host/caffe_compat/caffe/proto/*.inc) $(wildcard videovector_b200/csrc/host/caffe_compat/caffe/*.hpp) $(wildcard videovector_b200/csrc/host/caffe_compat/caffe/proto/*.hpp) $(wildcard videovector_b200/csrc/*.cuh) $(wildcard videovector_b200/csrc/host/*.h) $(wildcard videovector_b200/csrc/host/*.hpp)

all: $(LIBDIR)/libvv_b200.so build/vv_caffe

$(OBJDIR)/%.o: %.cu $(HDRS)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(OBJDIR)/%.o: %.cpp $(HDRS)
	@mkdir -p $(dir $@)
	$(CXX) -O2 -std=c++17 -fPIC -Wall -Iinclude -Ivideovector_b200/csrc -Ivideovector_b200/csrc/host/caffe_compat -I/usr/local/cuda/include -c $< -o $@

$(LIBDIR)/libvv_b200.so: $(CU_OBJS) $(CPP_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $^ -ldl

tools: build/vv_caffe
build/vv_caffe: tools/vv_caffe.cpp $(LIBDIR)/libvv_b200.so
	$(CXX) -O2 -std=c++17 -Iinclude -Ivideovector_b200/csrc/host/caffe_compat -I/usr/local/cuda/include $< -o $@ -L$(LIBDIR) -lvv_b200 -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)'

oracle: oracle/_build/libvv_oracle.so
oracle/_build/libvv_oracle.so: oracle/vv_oracle.cpp
	@mkdir -p oracle/_build
	$(CXX) -O2 -std=c++14 -Wno-deprecated-declarations -fPIC -shared -o $@ $< -ldl

ref:
	@if [ -d /root/reference/src/caffe ] && [ -f oracle/ref_shim/build_ref.sh ]; then bash oracle/ref_shim/build_ref.sh; else echo "reference tree or shim absent: skipping oracle/_ref"; fi

clean:
	rm -rf build $(LIBDIR)/libvv_b200.so oracle/_build

.PHONY: all oracle ref clean tools
