// rx_internal.h -- what the translation units of librxcuda.so share besides the kernels' launch interface:
// the few accessors rx_mgpu.cu needs into the context that rx_api.cu owns.  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "rxcuda.h"

struct RxMgpu;  // multi-GPU state of a context (rx_mgpu.cu)

// rx_api.cu
cudaStream_t rxi_stream(rxc_ctx* ctx);
int rxi_device(rxc_ctx* ctx);
int32_t rxi_fail(rxc_ctx* ctx, int32_t code, const std::string& msg);
RxMgpu** rxi_mgpu_slot(rxc_ctx* ctx);
void rxi_count_launch(rxc_ctx* ctx, uint32_t n);
// n frames into device memory, asynchronously on the context's stream; pitch_bytes = 0: tight rows
int32_t rxi_rasterize_device(rxc_ctx* ctx, const rxc_frame* frames, uint32_t n_frames, uint8_t* d_pixels, uint64_t stride, uint64_t pitch_bytes);

// rx_mgpu.cu
void rxi_mgpu_destroy(rxc_ctx* ctx);
