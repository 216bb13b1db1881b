# Builds the C-ABI library (sm_100a only), the C oracle pieces and, when the reference checkout is
# present, the reference oracle binary and the resql-b200 host front end.
NVCC ?= nvcc
NVCCFLAGS = -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
CSRC = resql_b200/csrc
LIB = resql_b200/libresql_b200.so

all: $(LIB)

$(LIB): $(CSRC)/engine.cu $(CSRC)/engine_exec.inl $(CSRC)/scan_kernel.cuh $(CSRC)/device_util.cuh $(CSRC)/hash_kernels.cuh \
        $(CSRC)/sort_kernels.cuh $(CSRC)/exchange_kernels.cuh $(CSRC)/host_narrow.h $(CSRC)/rq_internal.h $(CSRC)/dist.h include/resql_b200.h
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC)/engine.cu -ldl

ptxas-info:
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -shared -o /tmp/rq_ptxas.so $(CSRC)/engine.cu -ldl

oracle-ref:
	bash oracle/ref_build/build_ref.sh

host: $(LIB)
	bash resql_b200/host/build_host.sh

clean:
	rm -f $(LIB) resql_b200/host/resql-b200

.PHONY: all ptxas-info oracle-ref host clean
