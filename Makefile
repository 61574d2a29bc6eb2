# Builds libspike_b200.so (hand-written sm_100a CUDA + C ABI), the PETSc-shaped host glue and the
# CPU oracle.  `make` here is what __graft_entry__.build() runs.
NVCC    ?= /usr/local/cuda/bin/nvcc
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v $(NVEXTRA)
CSRC    := spike_petsc_b200/csrc
LIBDIR  := spike_petsc_b200/lib
OBJS    := $(LIBDIR)/layout.o $(LIBDIR)/lu.o $(LIBDIR)/tips.o $(LIBDIR)/solve.o $(LIBDIR)/msweep.o $(LIBDIR)/krylov.o $(LIBDIR)/peer.o $(LIBDIR)/wide_lu.o $(LIBDIR)/wide_sweep.o $(LIBDIR)/wide.o $(LIBDIR)/awbm.o $(LIBDIR)/capi.o

HOST    := spike_petsc_b200/host

all: $(LIBDIR)/libspike_b200.so $(LIBDIR)/libspike_petsc.so oracle

$(LIBDIR)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/lu_dev.cuh $(CSRC)/wide.cuh include/spike_b200.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/$*.ptxas.log || (cat $(LIBDIR)/$*.ptxas.log; exit 1)

$(LIBDIR)/libspike_b200.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

# PETSc-shaped host glue (C): PCBANDED / KSPREORDER / MatCreateSubMatrixBanded on the C ABI
HOSTSRC := $(HOST)/pcbanded.c $(HOST)/kspreorder.c $(HOST)/matbanded_type.c $(HOST)/ordering.c $(HOST)/petscshim.c $(HOST)/matio.c
$(LIBDIR)/libspike_petsc.so: $(HOSTSRC) $(HOST)/petscshim.h $(HOST)/petsc_access.h $(HOST)/spike_petsc.h $(LIBDIR)/libspike_b200.so
	/usr/bin/gcc -O2 -fPIC -Wall -shared -o $@ $(HOSTSRC) -L$(LIBDIR) -lspike_b200 -Wl,-rpath,'$$ORIGIN' -lm

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(LIBDIR)/*.o $(LIBDIR)/*.so $(LIBDIR)/*.log
	$(MAKE) -C oracle clean
.PHONY: all oracle clean
