# Build of the B200-native gargammel hot path: CUDA library (C-ABI), host tools, oracle.
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      ?= g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v
CSRC     := gargammel_b200/csrc
LIB      := gargammel_b200/libgargammel_b200.so
BINS     := bin/fragSim bin/deamSim bin/adptSim bin/gargammel-b200

all: $(LIB) $(BINS) oracle

$(LIB): $(CSRC)/gg_api.cu $(CSRC)/gg_kernels.cu $(CSRC)/gg_host.cpp $(CSRC)/gg_kernels.cuh $(CSRC)/gg_device.cuh $(CSRC)/gg_host.h $(CSRC)/gg_lz.h include/gargammel_b200.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/gg_api.cu $(CSRC)/gg_kernels.cu $(CSRC)/gg_host.cpp -lz 2> build_ptxas.log || (cat build_ptxas.log; exit 1)
	@grep -E "registers|spill" build_ptxas.log | sort | uniq -c | sort -rn | head -20

bin/%: $(CSRC)/cli_%.cpp $(CSRC)/cli_common.h $(CSRC)/gg_host.h include/gargammel_b200.h $(LIB)
	@mkdir -p bin
	$(CXX) -std=c++11 -O2 -Wall -o $@ $< -Lgargammel_b200 -lgargammel_b200 -Wl,-rpath,'$$ORIGIN/../gargammel_b200' -lz -lpthread

bin/gargammel-b200: $(CSRC)/cli_gargammel.cpp $(CSRC)/cli_common.h $(CSRC)/gg_host.h include/gargammel_b200.h $(LIB)
	@mkdir -p bin
	$(CXX) -std=c++11 -O2 -Wall -o $@ $< -Lgargammel_b200 -lgargammel_b200 -Wl,-rpath,'$$ORIGIN/../gargammel_b200' -lz -lpthread

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(LIB) $(BINS) build_ptxas.log; $(MAKE) -C oracle clean
.PHONY: all oracle clean
