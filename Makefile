# Build of the B200-native node-depth library, CLI, oracle and tools.
# Everything CUDA is compiled for sm_100a only.
NVCC      ?= nvcc
CXX       ?= g++
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function
CXXFLAGS  := -O3 -std=c++17 -fPIC -Wall -I/usr/local/cuda/include
CSRC      := pollen_b200/csrc
LIBDIR    := pollen_b200/lib
OBJDIR    := build/obj

LIB_OBJS  := $(OBJDIR)/depth_device.o $(OBJDIR)/tokenize.o $(OBJDIR)/interval_device.o $(OBJDIR)/depth_multi.o $(OBJDIR)/ops_depth.o $(OBJDIR)/ops_window_depth.o $(OBJDIR)/flatbed.o $(OBJDIR)/file.o $(OBJDIR)/parse.o $(OBJDIR)/print.o $(OBJDIR)/capi.o

all: $(LIBDIR)/libflatgfa.so $(LIBDIR)/libflatgfa.a $(LIBDIR)/libfgfa_synth.so bin/fgfa oracle tools build/depth_example

$(OBJDIR)/depth_device.o: $(CSRC)/depth_device.cu $(CSRC)/depth_kernels.cuh $(CSRC)/window_kernels.cuh include/fgfa_depth.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJDIR)/tokenize.o: $(CSRC)/tokenize.cu $(CSRC)/tokenize_kernels.cuh include/fgfa_depth.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJDIR)/interval_device.o: $(CSRC)/interval_device.cu $(CSRC)/interval_kernels.cuh include/fgfa_depth.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJDIR)/%.o: $(CSRC)/%.cpp $(wildcard $(CSRC)/*.hpp) include/fgfa_depth.h include/flatgfa.h
	@mkdir -p $(OBJDIR)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(LIBDIR)/libflatgfa.so: $(LIB_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $(LIB_OBJS) -cudart static -lpthread -ldl

# flatgfa-c/Cargo.toml:6-8 builds both a cdylib and a staticlib; link the archive with
#   g++ app.o libflatgfa.a -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread   (libnccl is dlopen'ed on first multi-GPU use)
$(LIBDIR)/libflatgfa.a: $(LIB_OBJS)
	@mkdir -p $(LIBDIR)
	rm -f $@ && ar rcs $@ $(LIB_OBJS)

$(LIBDIR)/libfgfa_synth.so: $(CSRC)/synth.cpp
	@mkdir -p $(LIBDIR)
	$(CXX) $(CXXFLAGS) -shared -o $@ $< -lpthread

bin/fgfa: $(CSRC)/fgfa_main.cpp $(LIBDIR)/libflatgfa.so
	@mkdir -p bin
	$(CXX) $(CXXFLAGS) -o $@ $< -L$(LIBDIR) -lflatgfa -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)'

# plain C client of the depth additions (compiled as C: the headers must stay C-clean)
build/depth_example: examples/depth.c include/flatgfa.h $(LIBDIR)/libflatgfa.so
	@mkdir -p build
	$(CC) -std=c11 -Wall -O2 -I include examples/depth.c -L$(LIBDIR) -lflatgfa -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)' -o $@

oracle:
	$(MAKE) -C oracle

tools: build/ubench build/sort_dedup_probe build/ubench_win build/ubench_smem build/ubench_ld
build/ubench_ld: tools/ubench_ld.cu $(CSRC)/window_kernels.cuh $(CSRC)/depth_kernels.cuh build/ubench
	$(NVCC) $(ARCH) -lineinfo -O3 -std=c++17 tools/ubench_ld.cu build/synth.o -o $@
build/ubench_win: tools/ubench_win.cu tools/experimental_window.cuh $(CSRC)/window_kernels.cuh $(CSRC)/depth_kernels.cuh build/ubench
	$(NVCC) $(ARCH) -lineinfo -O3 -std=c++17 tools/ubench_win.cu build/depth_oracle.o build/synth.o -o $@
build/ubench_smem: tools/ubench_smem.cu
	$(NVCC) $(ARCH) -lineinfo -O3 -std=c++17 tools/ubench_smem.cu -o $@
build/sort_dedup_probe: tools/sort_dedup_probe.cu $(CSRC)/synth.cpp oracle/depth_oracle.c build/ubench
	$(NVCC) $(ARCH) -O3 -std=c++17 tools/sort_dedup_probe.cu build/depth_oracle.o build/synth.o -o $@
build/ubench: tools/ubench.cu tools/experimental_kernels.cuh $(CSRC)/depth_kernels.cuh $(CSRC)/synth.cpp oracle/depth_oracle.c
	@mkdir -p build
	$(CC) -O3 -c oracle/depth_oracle.c -o build/depth_oracle.o
	$(CXX) -O3 -std=c++17 -c $(CSRC)/synth.cpp -o build/synth.o
	$(NVCC) $(ARCH) -lineinfo -O3 -std=c++17 tools/ubench.cu build/depth_oracle.o build/synth.o -o $@

# kernel-A variants for tools/ubench.cu (not part of `all`): packed depth counters, with and without
# the in-warp merge of same-word lanes; steps-per-thread.  See DESIGN.md section 5.
experiments: build/ubench
	for v in 16 8; do \
	  $(NVCC) -DFGFA_DEPTH_PACK=$$v $(ARCH) -lineinfo -O3 -std=c++17 tools/ubench.cu build/depth_oracle.o build/synth.o -o build/ubench_p$$v; \
	  $(NVCC) -DFGFA_DEPTH_PACK=$$v -DFGFA_PACK_MERGE=1 $(ARCH) -lineinfo -O3 -std=c++17 tools/ubench.cu build/depth_oracle.o build/synth.o -o build/ubench_p$${v}m; \
	done

# host code with AddressSanitizer + UndefinedBehaviorSanitizer (system g++ so that the runtime matches
# libasan.so.8); the CUDA objects are linked as they are.  Used by tools/asan_host.sh.
ASAN_SRCS := ops_depth ops_window_depth flatbed file parse print capi
asan: $(LIBDIR)/libflatgfa.so
	@mkdir -p build/asan
	for f in $(ASAN_SRCS); do /usr/bin/g++ -O1 -g -std=c++17 -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -I/usr/local/cuda/include -c $(CSRC)/$$f.cpp -o build/asan/$$f.o || exit 1; done
	$(NVCC) -ccbin /usr/bin/g++ $(ARCH) -shared -o build/asan/libflatgfa_asan.so $(OBJDIR)/depth_device.o $(OBJDIR)/tokenize.o $(OBJDIR)/interval_device.o $(addprefix build/asan/,$(addsuffix .o,$(ASAN_SRCS))) -cudart static -lpthread -Xlinker /usr/lib/x86_64-linux-gnu/libasan.so.8 -Xlinker /usr/lib/x86_64-linux-gnu/libubsan.so.1

TSAN_SRCS := ops_depth ops_window_depth flatbed file parse print
tsan: $(LIBDIR)/libflatgfa.so
	@mkdir -p build/tsan
	for f in $(TSAN_SRCS); do /usr/bin/g++ -O1 -g -std=c++17 -fPIC -fsanitize=thread -I/usr/local/cuda/include -c $(CSRC)/$$f.cpp -o build/tsan/$$f.o || exit 1; done
	/usr/bin/g++ -O1 -g -std=c++17 -fsanitize=thread -I/usr/local/cuda/include -c tools/tsan_host.cpp -o build/tsan/main.o
	$(NVCC) -ccbin /usr/bin/g++ $(ARCH) -o build/tsan/tsan_host build/tsan/main.o $(addprefix build/tsan/,$(addsuffix .o,$(TSAN_SRCS))) $(OBJDIR)/depth_device.o $(OBJDIR)/tokenize.o $(OBJDIR)/interval_device.o -cudart static -lpthread -Xlinker -ltsan

clean:
	rm -rf build bin $(LIBDIR)/*.so $(LIBDIR)/*.a oracle/*.so

.PHONY: all oracle tools clean experiments asan tsan
