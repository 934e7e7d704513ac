# Plain-make build for C / C++ consumers (the Python layer builds the same library through cover_b200/_build.py).
#   make            -> cover_b200/libcover_b200.so   (nvcc, sm_100a only)
#   make oracle     -> oracle/liboracle.so (+ oracle/_ref when /root/reference exists)   [test infrastructure]
#   make host_test  -> build/host_api_test (the reference's unit tests through include/cover_b200.hpp; needs a B200 to run)
NVCC ?= nvcc
CXX ?= g++
NVCCFLAGS ?= -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -cudart static
LIB := cover_b200/libcover_b200.so

all: $(LIB)

$(LIB): cover_b200/csrc/cover_b200.cu cover_b200/csrc/device_core.cuh cover_b200/csrc/area_rows.cuh cover_b200/csrc/lattice_words.cuh cover_b200/csrc/polygon_split.hpp include/cover_b200.h
	$(NVCC) $(NVCCFLAGS) -o $@ cover_b200/csrc/cover_b200.cu

oracle:
	$(MAKE) -C oracle liboracle.so ref

host_test: $(LIB)
	mkdir -p build
	$(CXX) -std=c++17 -O1 -Wall -Wextra -Iinclude tests/cpp/host_api_test.cpp -o build/host_api_test -Lcover_b200 -lcover_b200 -Wl,-rpath,$(abspath cover_b200)

clean:
	rm -f $(LIB); rm -rf build

.PHONY: all oracle host_test clean
