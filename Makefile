# Builds the CUDA library (sm_100a only) and the CPU oracle's C restatement.
NVCC ?= /usr/local/cuda/bin/nvcc
NVCCFLAGS := -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a \
             -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr -cudart static
SRC := cvpack_b200/csrc
OBJ := build/obj
LIB := cvpack_b200/libcvpack_b200.so
SOURCES := $(wildcard $(SRC)/*.cu)
OBJECTS := $(patsubst $(SRC)/%.cu,$(OBJ)/%.o,$(SOURCES))

all: $(LIB) oracle

$(OBJ)/%.o: $(SRC)/%.cu $(wildcard $(SRC)/*.cuh) include/cvpack_b200.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVCCFLAGS) $(EXTRA) -c $< -o $@

$(LIB): $(OBJECTS)
	$(NVCC) -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o $@ $^

oracle: oracle/liboracle.so

oracle/liboracle.so: oracle/cv_oracle.c
	gcc -O3 -march=x86-64-v3 -fopenmp -fPIC -shared -o $@ $< -lm

clean:
	rm -rf build $(LIB) oracle/liboracle.so

.PHONY: all oracle clean
