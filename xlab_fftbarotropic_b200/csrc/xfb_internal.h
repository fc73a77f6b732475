// xfb_internal.h -- launcher interface between the C-ABI layer and the kernel translation units.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

#include "xfb_col.cuh"
#include "xfb_row.cuh"

namespace xfb {

// column tile width used for a given x length (0 = size not served by the fused kernels)
int col_tile_width(int nx);
// true if the stepper's K-COL for this x length is the two-level kernel (xfb_col2l.cuh)
bool col_two_level(int nx);
bool row_size_ok(int ny);

// returns cudaError_t as int
int launch_row(int ny, int mode, const RowParams &p, cudaStream_t st);
int launch_col(int nx, int mode, const ColParams &p, int batch, cudaStream_t st);

// Launch configuration that must exist once PER DEVICE: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the
// number of resident CTAs of a persistent kernel.  xfb_create(..., device) allows handles on several GPUs of one
// process and several host threads may launch at once, so the cache is indexed by the current device and written
// atomically; the initialiser is idempotent (two racing first calls compute the same value).
struct PerDeviceInt {
    std::atomic<int> v[64];
    // init() -> value > 0, or <= 0 on failure with *err set to the cudaError_t
    template <class F>
    int get(F init, int *err)
    {
        int d = 0;
        cudaGetDevice(&d);
        d &= 63;
        int x = v[d].load(std::memory_order_acquire);
        if (x <= 0) {
            x = init(err);
            if (x > 0) v[d].store(x, std::memory_order_release);
        }
        return x;
    }
};

// occupancy-limited number of CTAs of a persistent kernel on the current device, after raising its shared-memory limit
template <class K>
static inline int resident_ctas(K kernel, int threads, int smem, int tmem_cols, int *err)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { *err = (int)e; return 0; }
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (tmem_cols > 0 && per_sm * tmem_cols > 512) per_sm = 512 / tmem_cols;     // TMEM columns are a per-SM resource too
    return sms * (per_sm > 0 ? per_sm : 1);
}

}  // namespace xfb
