// xfb_row.cu -- instantiations and launcher of the K-ROW kernels.
#include "xfb_internal.h"

namespace xfb {

bool row_size_ok(int ny)
{
    switch (ny) {
    case 256: case 512: case 1024: case 2048: case 4096: case 8192: case 16384: return true;
    default: return false;
    }
}

template <int NY, int MODE, bool DIST>
static int launch_row_t(const RowParams &p, cudaStream_t st)
{
    typedef RowCfg<NY> C;
    constexpr int smem = (MODE == ROW_JAC) ? C::SMEM_JAC : C::SMEM;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(row_kernel<NY, MODE, DIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int blocks = (p.nrows + C::LPC - 1) / C::LPC;
    row_kernel<NY, MODE, DIST><<<blocks, C::THREADS, smem, st>>>(p);
    return (int)cudaGetLastError();
}

template <int NY>
static int launch_row_n(int mode, const RowParams &p, cudaStream_t st)
{
    const bool dist = p.cw > 0;
    switch (mode) {
    case ROW_R2C: return dist ? launch_row_t<NY, ROW_R2C, true>(p, st) : launch_row_t<NY, ROW_R2C, false>(p, st);
    case ROW_C2R: return dist ? launch_row_t<NY, ROW_C2R, true>(p, st) : launch_row_t<NY, ROW_C2R, false>(p, st);
    case ROW_JAC: return dist ? launch_row_t<NY, ROW_JAC, true>(p, st) : launch_row_t<NY, ROW_JAC, false>(p, st);
    }
    return (int)cudaErrorInvalidValue;
}

int launch_row(int ny, int mode, const RowParams &p, cudaStream_t st)
{
    switch (ny) {
    case 256: return launch_row_n<256>(mode, p, st);
    case 512: return launch_row_n<512>(mode, p, st);
    case 1024: return launch_row_n<1024>(mode, p, st);
    case 2048: return launch_row_n<2048>(mode, p, st);
    case 4096: return launch_row_n<4096>(mode, p, st);
    case 8192: return launch_row_n<8192>(mode, p, st);
    case 16384: return launch_row_n<16384>(mode, p, st);
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace xfb
