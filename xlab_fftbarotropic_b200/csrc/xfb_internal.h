// xfb_internal.h -- launcher interface between the C-ABI layer and the kernel translation units.
#pragma once
#include <cuda_runtime.h>

#include "xfb_col.cuh"
#include "xfb_row.cuh"

namespace xfb {

// column tile width used for a given x length (0 = size not served by the fused kernels)
int col_tile_width(int nx);
bool row_size_ok(int ny);

// returns cudaError_t as int
int launch_row(int ny, int mode, const RowParams &p, cudaStream_t st);
int launch_col(int nx, int mode, const ColParams &p, int batch, cudaStream_t st);

}  // namespace xfb
