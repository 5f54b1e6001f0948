# Builds the C-ABI shared library for sm_100a (in-tree, so it travels to the GPU box with the snapshot).
NVCC ?= nvcc
CSRC := easyfea_b200/csrc
SRCS := $(CSRC)/api.cu $(CSRC)/elem_kernels.cu $(CSRC)/pf_kernels.cu $(CSRC)/csr_kernels.cu $(CSRC)/pcg_kernels.cu $(CSRC)/post_kernels.cu $(CSRC)/fused_kernels.cu $(CSRC)/fused_mma.cu
OBJS := $(SRCS:.cu=.o)
LIB := easyfea_b200/libeasyfea_b200.so
NVFLAGS := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v

all: $(LIB)

$(CSRC)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/easyfea_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) -cudart static

clean:
	rm -f $(OBJS) $(CSRC)/*.ptxas.log $(LIB)
