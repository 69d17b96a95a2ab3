# Top-level build: libwrp.so (CUDA kernels + C ABI, sm_100a), libwrphost.so (C++ host mirror of
# the reference's Dimension/Sector/RadarProcessor API on top of the C ABI), the oracle.
PKG      := weather-radar-processing_b200
CSRC     := $(PKG)/csrc
HOST     := $(PKG)/host
NVCC     ?= nvcc
PYTHON   ?= python3
CXX      := g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -std=c++17 -O3 -lineinfo -Xcompiler -fPIC
LIB      := $(PKG)/libwrp.so
OBJS     := $(CSRC)/wrp_fused.o $(CSRC)/wrp_persistent.o $(CSRC)/wrp_stream.o $(CSRC)/wrp_staged.o $(CSRC)/wrp_api.o $(CSRC)/wrp_tables.o $(CSRC)/wrp_volume.o

HOSTLIB  := $(PKG)/libwrphost.so
HOSTSRC  := $(HOST)/dimension.cpp $(HOST)/sector.cpp $(HOST)/floats.c $(HOST)/radar_processor.cpp $(HOST)/stage_dump.cpp
HOSTBINS := $(HOST)/wrp_chain $(HOST)/host_selftest

all: $(LIB) $(HOSTLIB) $(HOSTBINS) oracle

# C++ host mirror of the reference's API, on top of the C ABI only
$(HOSTLIB): $(HOSTSRC) $(HOST)/*.h include/wrp.h $(LIB)
	$(CXX) -O2 -std=c++17 -fPIC -shared -o $@ $(HOST)/dimension.cpp $(HOST)/sector.cpp -x c++ $(HOST)/floats.c -x none \
	    $(HOST)/radar_processor.cpp $(HOST)/stage_dump.cpp -L$(PKG) -lwrp -Wl,-rpath,'$$ORIGIN'

$(HOST)/wrp_chain: $(HOST)/wrp_chain.cpp $(HOSTLIB)
	$(CXX) -O2 -std=c++17 -o $@ $< -L$(PKG) -lwrphost -lwrp -Wl,-rpath,'$$ORIGIN/..'

$(HOST)/host_selftest: $(HOST)/host_selftest.cpp $(HOSTLIB)
	$(CXX) -O2 -std=c++17 -o $@ $< -L$(PKG) -lwrphost -lwrp -Wl,-rpath,'$$ORIGIN/..'

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/wrp_internal.h $(CSRC)/wrp_fft.cuh $(CSRC)/wrp_ptx.cuh $(CSRC)/wrp_chain_params.h $(CSRC)/wrp_stream.h include/wrp.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(CSRC)/wrp_tables.o: $(CSRC)/wrp_tables.cpp $(CSRC)/wrp_internal.h include/wrp.h
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(CSRC)/wrp_volume.o: $(CSRC)/wrp_volume.cpp include/wrp.h
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart static
	$(PYTHON) tools/check_sass.py $@

oracle:
	$(MAKE) -C oracle liboracle.so

# Checked build of the streaming kernels (WRP_CHECK range checks that trap): run the GPU tests against it with
#   WRP_LIB=$$PWD/tools/libwrp_checked.so python -m pytest tests -m gpu
checked: $(LIB)
	$(NVCC) $(NVFLAGS) -DWRP_CHECKED -c $(CSRC)/wrp_stream.cu -o $(CSRC)/wrp_stream_checked.o
	$(NVCC) $(ARCH) -shared -o tools/libwrp_checked.so $(filter-out $(CSRC)/wrp_stream.o,$(OBJS)) $(CSRC)/wrp_stream_checked.o -cudart static

clean:
	rm -f $(OBJS) $(LIB) $(HOSTLIB) $(HOSTBINS)

.PHONY: all oracle clean checked
