# Builds libipoke_b200.so (sm_100a only) and the oracle helpers.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v --expt-relaxed-constexpr
SRC := ipoke_b200/csrc
OBJDIR := build/obj
LIBDIR := ipoke_b200/lib
SOURCES := $(wildcard $(SRC)/*.cu)
OBJECTS := $(patsubst $(SRC)/%.cu,$(OBJDIR)/%.o,$(SOURCES))
HEADERS := $(wildcard $(SRC)/*.cuh) include/ipoke_b200.h

all: $(LIBDIR)/libipoke_b200.so

$(OBJDIR)/%.o: $(SRC)/%.cu $(HEADERS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIBDIR)/libipoke_b200.so: $(OBJECTS)
	@mkdir -p $(LIBDIR)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJECTS) -lcudart -ldl

clean:
	rm -rf build $(LIBDIR)
.PHONY: all clean
