# Builds libmodsgpu.so (sm_100a only) in-tree, plus the oracle (test infrastructure).
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH  = -gencode arch=compute_100a,code=sm_100a
PKG   = mods_light_zmq_b200
SRC   = $(PKG)/csrc
OBJ   = $(SRC)/build
NVFLAGS = $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off -Iinclude
# detect.cu / sampler.cu are the bit-exact integer/float paths: no FMA contraction
EXACT = --fmad=false

OBJS = $(OBJ)/api.o $(OBJ)/detect.o $(OBJ)/sampler.o $(OBJ)/cnn.o $(OBJ)/match.o $(OBJ)/ransac.o $(OBJ)/ransac_f.o $(OBJ)/synth.o $(OBJ)/classic.o $(OBJ)/chain.o $(OBJ)/npz.o $(OBJ)/mods_host.o

all: $(PKG)/libmodsgpu.so $(PKG)/libmodsgpu_degensac.so oracle example

$(OBJ)/detect.o: $(SRC)/detect.cu $(SRC)/common.cuh include/modsgpu.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(OBJ)/sampler.o: $(SRC)/sampler.cu $(SRC)/common.cuh include/modsgpu.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(OBJ)/ransac.o: $(SRC)/ransac.cu $(SRC)/ransac_common.cuh $(SRC)/ransac_h.cuh $(SRC)/common.cuh include/modsgpu.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(OBJ)/synth.o: $(SRC)/synth.cu $(SRC)/common.cuh include/modsgpu.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(OBJ)/chain.o: $(SRC)/chain.cu $(SRC)/common.cuh include/modsgpu.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(OBJ)/classic.o: $(SRC)/classic.cu $(SRC)/common.cuh include/modsgpu.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) $(EXACT) -c $< -o $@
$(OBJ)/ransac_f.o: $(SRC)/ransac_f.cu $(SRC)/ransac_common.cuh $(SRC)/ransac_h.cuh $(SRC)/common.cuh include/modsgpu.h
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

# link-time drop-in for the reference's degensac target (exp_ransacHcustom / exp_ransacFcustom)
$(PKG)/libmodsgpu_degensac.so: $(SRC)/compat/degensac_compat.cpp include/modsgpu.h $(PKG)/libmodsgpu.so
	g++ -O2 -std=c++17 -fPIC -shared -o $@ $< -L$(PKG) -lmodsgpu -Wl,-rpath,'$$ORIGIN'

# smallest host program over the C ABI (plain C): one pair, PPM / PGM in, correspondences out
example: examples/mods_pair
examples/mods_pair: examples/mods_pair.c include/modsgpu.h $(PKG)/libmodsgpu.so
	gcc -O2 -std=c11 -Wall -Iinclude -o $@ $< -L$(PKG) -lmodsgpu -Wl,-rpath,'$$ORIGIN/../$(PKG)'

oracle:
	$(MAKE) -C oracle -s all

clean:
	rm -rf $(OBJ) $(PKG)/libmodsgpu.so $(PKG)/libmodsgpu_degensac.so examples/mods_pair
	$(MAKE) -C oracle clean

.PHONY: all oracle clean example
