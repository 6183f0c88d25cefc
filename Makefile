# Builds the product: libspb200.so (CUDA kernels + C ABI, sm_100a only) and the spb200 host binary.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
ARCH      := -gencode arch=compute_100a,code=sm_100a
PHASEFLAG := $(if $(PHASE),-DSPB_PHASE_TIMING,)
NVFLAGS   := $(PHASEFLAG) $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-fopenmp --expt-extended-lambda -ccbin $(CXX)
CSRC      := spacepharer_b200/csrc
LIBDIR    := spacepharer_b200/lib
CU        := $(CSRC)/capi.cu $(CSRC)/index.cu $(CSRC)/prefilter.cu $(CSRC)/align.cu $(CSRC)/frames.cu $(CSRC)/nuclaln.cu
OBJ       := $(patsubst $(CSRC)/%.cu,build/%.o,$(CU))
HOSTLIBSRC := $(CSRC)/host/tantan.cpp
HOSTLIBOBJ := $(patsubst $(CSRC)/host/%.cpp,build/host_%.o,$(HOSTLIBSRC))
HOSTSRC   := $(filter-out $(HOSTLIBSRC),$(wildcard $(CSRC)/host/*.cpp))
HOSTOBJ   := $(patsubst $(CSRC)/host/%.cpp,build/host_%.o,$(HOSTSRC))

ifneq ($(HOSTSRC),)
all: $(LIBDIR)/libspb200.so $(LIBDIR)/spb200
else
all: $(LIBDIR)/libspb200.so
endif

build/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh include/spb200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIBDIR)/libspb200.so: $(OBJ) $(HOSTLIBOBJ)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -cudart static -Xcompiler -fopenmp -ccbin $(CXX) -o $@ $(OBJ) $(HOSTLIBOBJ)

build/host_%.o: $(CSRC)/host/%.cpp $(wildcard $(CSRC)/host/*.h) include/spb200.h
	@mkdir -p build
	$(CXX) -O2 -g -std=c++17 -Wall -fPIC -fopenmp -ffp-contract=off -Iinclude -c $< -o $@

$(LIBDIR)/spb200: $(HOSTOBJ) $(LIBDIR)/libspb200.so
	$(CXX) -fopenmp -o $@ $(HOSTOBJ) -L$(LIBDIR) -lspb200 -Wl,-rpath,'$$ORIGIN' -ldl -lpthread -lrt

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIBDIR)/libspb200.so $(LIBDIR)/spb200

.PHONY: all oracle clean
