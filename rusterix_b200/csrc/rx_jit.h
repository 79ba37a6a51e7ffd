// rx_jit.h -- interface of the run-time kernel compilation (rx_jit.cu: batch shaders as straight-line code, raster kernels recompiled
// with a scene's constants, the analysis of which programs can observe a carried Execution) towards rx_api.cu.  Not part of the C ABI.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "rxcuda.h"

struct RxJit;
// C++ for the programs the translator accepts (vm_run_jit + one function per program); jit_index[i] = i or 0xFFFFFFFF.
// Returns false when no program was accepted.
bool rxj_generate(const rxc_program* progs, uint32_t n, std::string* source, std::vector<uint32_t>* jit_index);
// out[i]: 0 = program i cannot tell the reference's per-tile Execution from the device's fresh one, 1 = it can (see rx_jit.cu),
// 2 = not analysable.  usage[i]: bit 0 = bound to a 3D batch, bit 1 = to a 2D batch (nullptr: both).
void rxj_state_report(const rxc_program* progs, uint32_t n, const uint8_t* usage, bool scene_has_3d, uint32_t* out);
// mode: 0 off, 1 compile in the background (frames use the interpreter until the kernel is ready), 2 compile synchronously
RxJit* rxj_create(const std::string& generated, int mode);   // one per context; the kernels it loads outlive the scenes
void rxj_set_programs(RxJit* j, const std::string& generated, int mode);   // a new scene's programs (may be empty)
void rxj_destroy(RxJit* j);
// cudaKernel_t of k_raster<sample_mode, planes, raster_mode> compiled with `defines` (space-separated -D switches: the scene's
// constants) and, for raster mode 2, the generated code -- sample_mode -1: of k_vm_execute -- or nullptr (not ready / failed / off)
void* rxj_kernel(RxJit* j, int sample_mode, bool planes, int raster_mode, const std::string& defines, std::string* failed);
int rxj_idle(RxJit* j);
void rxj_shutdown();
size_t rxj_compile_offline(const std::string& generated, int sample_mode, bool planes, int raster_mode, const std::string& defines, std::string* log);
void rxj_stats(RxJit* j, uint64_t* compiled, uint64_t* used);
