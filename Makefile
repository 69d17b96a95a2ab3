# Top-level build: libwrp.so (CUDA kernels + C ABI, sm_100a), libwrphost.so (C++ host mirror of
# the reference's Dimension/Sector/RadarProcessor API on top of the C ABI), the oracle.
PKG      := weather-radar-processing_b200
CSRC     := $(PKG)/csrc
HOST     := $(PKG)/host
NVCC     ?= nvcc
CXX      := g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -std=c++17 -O3 -lineinfo -Xcompiler -fPIC
LIB      := $(PKG)/libwrp.so
OBJS     := $(CSRC)/wrp_fused.o $(CSRC)/wrp_persistent.o $(CSRC)/wrp_unified.o $(CSRC)/wrp_staged.o $(CSRC)/wrp_api.o $(CSRC)/wrp_tables.o

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

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/wrp_internal.h $(CSRC)/wrp_fft.cuh $(CSRC)/wrp_ptx.cuh $(CSRC)/wrp_chain_params.h include/wrp.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(CSRC)/wrp_tables.o: $(CSRC)/wrp_tables.cpp $(CSRC)/wrp_internal.h include/wrp.h
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart static
	python tools/check_sass.py $@

oracle:
	$(MAKE) -C oracle liboracle.so

# Experimental, not validated on a GPU: the pair kernel (csrc/experimental/wrp_pair.cu) linked in
# place of wrp_unified.o.  A/B it with  WRP_LIB=$$PWD/tools/libwrp_pair.so python tools/ab.py ...
pair: $(LIB)
	$(NVCC) $(NVFLAGS) -c $(CSRC)/experimental/wrp_pair.cu -o $(CSRC)/experimental/wrp_pair.o
	$(NVCC) $(ARCH) -shared -o tools/libwrp_pair.so $(filter-out $(CSRC)/wrp_unified.o,$(OBJS)) $(CSRC)/experimental/wrp_pair.o -cudart static
	python tools/check_sass.py tools/libwrp_pair.so

clean:
	rm -f $(OBJS) $(LIB) $(HOSTLIB) $(HOSTBINS)

# Experimental variants of chain_unified_kernel, none validated on a GPU yet (the default build is
# byte-identical without the guarded code): each becomes tools/libwrp_<name>.so for tools/ab.py.
#   split      exchange rendezvous as a split-phase mbarrier, Doppler arithmetic in between
#   x2last     x2 hand-off stores carry an L2 evict-last policy
#   rowsfirst  ring rows are read with an L2 evict-first policy (dead after the read)
#   l2both     x2last + rowsfirst
VARIANTS := split x2last rowsfirst l2both
FLAGS_split     := -DWRP_UNI_SPLIT_BARRIER
FLAGS_x2last    := -DWRP_UNI_X2_EVICT_LAST
FLAGS_rowsfirst := -DWRP_UNI_ROWS_EVICT_FIRST
FLAGS_l2both    := -DWRP_UNI_X2_EVICT_LAST -DWRP_UNI_ROWS_EVICT_FIRST
variants: $(addprefix tools/libwrp_,$(addsuffix .so,$(VARIANTS)))
tools/libwrp_%.so: $(LIB) $(CSRC)/wrp_unified.cu
	$(NVCC) $(NVFLAGS) $(FLAGS_$*) -c $(CSRC)/wrp_unified.cu -o $(CSRC)/experimental/wrp_unified_$*.o
	$(NVCC) $(ARCH) -shared -o $@ $(filter-out $(CSRC)/wrp_unified.o,$(OBJS)) $(CSRC)/experimental/wrp_unified_$*.o -cudart static
	python tools/check_sass.py $@

.PHONY: all oracle clean pair variants
