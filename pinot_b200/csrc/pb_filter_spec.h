// pb_filter_spec.h — plan-time specialised instantiations of pb_filter_kernel (pb_filter_spec.cu), reached through a small
// dispatch table so that pb_engine.cu does not have to compile them.  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
struct DevQuery;
// PK: 0 = dictId range, 1 = IN / NOT IN membership LUT in shared memory
bool pb_filter_spec_available(int width, int pred_kind);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) on the current device + resident CTAs per SM for `smem` bytes
cudaError_t pb_filter_spec_prepare(int width, int pred_kind, size_t smem, int* ctas_per_sm);
cudaError_t pb_filter_spec_launch(int width, int pred_kind, int grid, size_t smem, cudaStream_t st, const DevQuery* q);
