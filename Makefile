# Builds libmodsgpu.so (sm_100a only) in-tree, plus the oracle (test infrastructure).
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH  = -gencode arch=compute_100a,code=sm_100a
PKG   = mods_light_zmq_b200
SRC   = $(PKG)/csrc
OBJ   = $(SRC)/build
NVFLAGS = $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off -Iinclude
# detect.cu / sampler.cu are the bit-exact integer/float paths: no FMA contraction
EXACT = --fmad=false

OBJS = $(OBJ)/api.o $(OBJ)/detect.o $(OBJ)/sampler.o $(OBJ)/cnn.o $(OBJ)/match.o $(OBJ)/ransac.o $(OBJ)/npz.o $(OBJ)/mods_host.o

all: $(PKG)/libmodsgpu.so oracle

$(OBJ)/detect.o: $(SRC)/detect.cu $(SRC)/common.cuh include/modsgpu.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(OBJ)/sampler.o: $(SRC)/sampler.cu $(SRC)/common.cuh include/modsgpu.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(OBJ)/ransac.o: $(SRC)/ransac.cu $(SRC)/common.cuh include/modsgpu.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(OBJ)/%.o: $(SRC)/%.cu $(SRC)/common.cuh include/modsgpu.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(OBJ)/mods_host.o: $(SRC)/host/mods_host.cpp $(SRC)/host/mods_host.h include/modsgpu.h
	@mkdir -p $(OBJ)
	g++ -O2 -std=c++17 -fPIC -ffp-contract=off -c $< -o $@
$(OBJ)/npz.o: $(SRC)/npz.cpp
	@mkdir -p $(OBJ)
	g++ -O2 -std=c++17 -fPIC -c $< -o $@

$(PKG)/libmodsgpu.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lz -lcudart_static -lpthread -ldl -lrt

oracle:
	$(MAKE) -C oracle -s all

clean:
	rm -rf $(OBJ) $(PKG)/libmodsgpu.so
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
