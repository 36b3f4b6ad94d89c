# Builds the product: libssb200.so (sm_100a CUDA behind a C ABI) and the two drop-in C mains.
#   make            -> stochasticsim_b200/lib/{libssb200.so,tncCountsProfile,stochasticSpike} + tools
#   make oracle     -> CPU restatements (test infrastructure) in oracle/_build/
#   make ref        -> unmodified reference compiled into oracle/_ref/ (authoring container only)
NVCC      ?= /usr/local/cuda/bin/nvcc
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v $(EXTRA_NVFLAGS)
LIBDIR    := stochasticsim_b200/lib
CSRC      := stochasticsim_b200/csrc
HOST      := stochasticsim_b200/host
CU        := $(wildcard $(CSRC)/*.cu)
OBJ       := $(patsubst $(CSRC)/%.cu,$(LIBDIR)/obj/%.o,$(CU))
HOSTSRC   := $(wildcard $(HOST)/*.c)
BINS      := $(patsubst $(HOST)/%.c,$(LIBDIR)/%,$(HOSTSRC))

all: $(LIBDIR)/libssb200.so $(BINS) tools

$(LIBDIR)/obj/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/ssb200.h
	@mkdir -p $(LIBDIR)/obj
	$(NVCC) $(NVFLAGS) -Iinclude -c $< -o $@ 2> $(LIBDIR)/obj/$*.ptxas.log || (cat $(LIBDIR)/obj/$*.ptxas.log; false)
	@grep -E "error|warning" $(LIBDIR)/obj/$*.ptxas.log | grep -v "ptxas info" || true

$(LIBDIR)/libssb200.so: $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart -ldl

$(LIBDIR)/%: $(HOST)/%.c $(LIBDIR)/libssb200.so include/ssb200.h $(wildcard $(HOST)/*.h)
	$(CC) -std=c99 -O2 -Wall -D_GNU_SOURCE -Iinclude -o $@ $< -L$(LIBDIR) -lssb200 -Wl,-rpath,'$$ORIGIN' -lm -lz -lpthread

tools: tools/_build/gen_synth tools/_build/libsynth.so

tools/_build/gen_synth: tools/gen_synth.c
	@mkdir -p tools/_build
	$(CC) -std=c99 -O2 -Wall -DGEN_SYNTH_MAIN -o $@ $< -lm

tools/_build/libsynth.so: tools/gen_synth.c
	@mkdir -p tools/_build
	$(CC) -std=c99 -O2 -Wall -fPIC -shared -o $@ $< -lm

oracle:
	$(MAKE) -C oracle

ref:
	$(MAKE) -C oracle ref

clean:
	rm -rf $(LIBDIR) tools/_build oracle/_build

.PHONY: all tools oracle ref clean
