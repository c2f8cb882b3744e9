# Build recipe of the B200 alignment core.
#   make lib      -> mpibwa_b200/libmpibwa_b200.so  (the product: C-ABI + host orchestration + sm_100a kernels)
#   make driver   -> tools/b200_driver              (single-process stand-in for the mpiBWA chunk loop)
#   make hostemu  -> tests/_build/*                 (TEST ONLY: host orchestration over CPU-looped stage bodies)
#   make oracle   -> oracle/liboracle.so + oracle/_ref/* (TEST ONLY)
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      ?= g++
CC       ?= gcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the finish stages reproduce the reference's double/float decisions; a*b+c must stay two IEEE operations
NVFLAGS  := -O3 -std=c++17 $(ARCH) -lineinfo -fmad=false -Xcompiler -fPIC,-O3,-pthread -Xptxas -v --expt-relaxed-constexpr
CXXFLAGS := -O3 -std=c++17 -fPIC -pthread -Wall -Wno-unused-function
CSRC     := mpibwa_b200/csrc
HOSTSRC  := $(CSRC)/capi.cpp $(CSRC)/host_align.cpp $(CSRC)/pipeline.cpp $(CSRC)/hostshim.cpp
HDRS     := $(wildcard $(CSRC)/*.h) include/mpibwa_b200.h
LIB      := mpibwa_b200/libmpibwa_b200.so
B        := build

all: lib driver

lib: $(LIB)

$(B)/%.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p $(B)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(B)/%.o: $(CSRC)/%.cu $(HDRS) $(wildcard $(CSRC)/*.cuh)
	@mkdir -p $(B)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(B)/$*.ptxas.log || (cat $(B)/$*.ptxas.log; false)

$(LIB): $(B)/capi.o $(B)/host_align.o $(B)/pipeline.o $(B)/hostshim.o $(B)/stages_cuda.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static -lpthread

driver: tools/b200_driver
tools/b200_driver: tools/b200_driver.c $(LIB)
	$(CC) -O2 -Wall -Iinclude tools/b200_driver.c -o $@ -Lmpibwa_b200 -lmpibwa_b200 -Wl,-rpath,'$$ORIGIN/../mpibwa_b200'

hostemu: tests/_build/libmpibwa_b200_hostemu.so tests/_build/b200_driver_hostemu
tests/_build/libmpibwa_b200_hostemu.so: $(HOSTSRC) tests/hostemu/stages_emu.cpp tests/hostemu/emu_bodies.h $(HDRS)
	@mkdir -p tests/_build
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOSTSRC) tests/hostemu/stages_emu.cpp
tests/_build/b200_driver_hostemu: tools/b200_driver.c tests/_build/libmpibwa_b200_hostemu.so
	$(CC) -O2 -Wall -Iinclude tools/b200_driver.c -o $@ -Ltests/_build -lmpibwa_b200_hostemu -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(B) $(LIB) tools/b200_driver tests/_build
.PHONY: all lib driver hostemu oracle clean
