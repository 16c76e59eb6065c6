# Builds the C-ABI shared library (sm_100a only) in-tree so it travels to the GPU box.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v $(EXTRA)
CSRC      := spectral_petsc_b200/csrc
OBJDIR    := build
LIB       := spectral_petsc_b200/libspectral_b200.so
CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CPP_SRCS  := $(wildcard $(CSRC)/*.cpp)
HOST      := spectral_petsc_b200/host
HOST_SRCS := $(wildcard $(HOST)/*.cpp)
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(CU_SRCS)) $(patsubst $(CSRC)/%.cpp,$(OBJDIR)/%.o,$(CPP_SRCS)) $(patsubst $(HOST)/%.cpp,$(OBJDIR)/host_%.o,$(HOST_SRCS))
HDRS      := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)

DRIVER    := tests/cpp/ref_api_driver
DRIVER2   := tests/cpp/ref_api_driver2
APP_ELL   := apps/elliptic
APP_STK   := apps/stokes
APP_CHEB  := apps/cheb

all: $(LIB) $(DRIVER) $(DRIVER2) $(APP_ELL) $(APP_STK) $(APP_CHEB)

$(DRIVER): tests/cpp/ref_api_driver.cpp $(LIB) $(HDRS)
	g++ -O2 -std=c++17 -o $@ $< -Lspectral_petsc_b200 -lspectral_b200 -Wl,-rpath,'$$ORIGIN/../../spectral_petsc_b200'

$(DRIVER2): tests/cpp/ref_api_driver2.cpp $(LIB) $(HDRS)
	g++ -O2 -std=c++17 -o $@ $< -Lspectral_petsc_b200 -lspectral_b200 -Wl,-rpath,'$$ORIGIN/../../spectral_petsc_b200'

$(APP_ELL): apps/elliptic.cpp apps/common.h $(LIB) $(HDRS)
	g++ -O2 -std=c++17 -o $@ $< -Lspectral_petsc_b200 -lspectral_b200 -Wl,-rpath,'$$ORIGIN/../spectral_petsc_b200'

$(APP_STK): apps/stokes.cpp apps/common.h $(LIB) $(HDRS)
	g++ -O2 -std=c++17 -o $@ $< -Lspectral_petsc_b200 -lspectral_b200 -Wl,-rpath,'$$ORIGIN/../spectral_petsc_b200'

$(APP_CHEB): apps/cheb.cpp apps/common.h $(LIB) $(HDRS)
	g++ -O2 -std=c++17 -o $@ $< -Lspectral_petsc_b200 -lspectral_b200 -Wl,-rpath,'$$ORIGIN/../spectral_petsc_b200'

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)

$(OBJDIR)/%.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p $(OBJDIR)
	g++ -O2 -std=c++17 -fPIC -c $< -o $@

$(OBJDIR)/host_%.o: $(HOST)/%.cpp $(HDRS)
	@mkdir -p $(OBJDIR)
	g++ -O2 -std=c++17 -fPIC -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

# diagnostic build with the ablation / timeline switches compiled in (SB200_XFLAGS is read only by this library; the tools that
# need it load it with SB200_ABLATE_LIB=1)
ablate:
	$(MAKE) OBJDIR=build_ablate LIB=spectral_petsc_b200/libspectral_b200_ablate.so EXTRA=-DSB200_ABLATE spectral_petsc_b200/libspectral_b200_ablate.so

clean:
	rm -rf build_ablate spectral_petsc_b200/libspectral_b200_ablate.so
	rm -rf $(OBJDIR) $(LIB) $(DRIVER) $(DRIVER2) $(APP_ELL) $(APP_STK) $(APP_CHEB)

.PHONY: all clean ablate
