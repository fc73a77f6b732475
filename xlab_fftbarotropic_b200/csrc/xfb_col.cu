// xfb_col.cu -- instantiations and launcher of the K-COL kernels.
#include "xfb_internal.h"

namespace xfb {

#include <cstdlib>

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// Column tile width W (adjacent complex columns per CTA).  A tile of NX x W complex values must fit
// in shared memory: NX x 4 at 8192 would need 278 KB.  XFB_COL_W overrides the default where two
// instantiations exist (tuning knob).
int col_tile_width(int nx)
{
    static const int forced = env_int("XFB_COL_W", 0);
    switch (nx) {
    case 256: case 512: case 1024: case 2048: return 4;
    case 4096: return (forced == 2 || forced == 4) ? forced : 2;
    case 8192: return (forced == 1 || forced == 2) ? forced : 2;
    case 16384: return 1;
    default: return 0;
    }
}

template <int NX, int W, int MODE>
static int launch_col_t(const ColParams &p, int batch, cudaStream_t st)
{
    typedef ColCfg<NX, W> C;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(col_kernel<NX, W, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    dim3 grid(p.pitch / W, batch);
    col_kernel<NX, W, MODE><<<grid, C::THREADS, C::SMEM, st>>>(p);
    return (int)cudaGetLastError();
}

template <int NX, int W>
static int launch_col_n(int mode, const ColParams &p, int batch, cudaStream_t st)
{
    switch (mode) {
    case COL_FWD: return launch_col_t<NX, W, COL_FWD>(p, batch, st);
    case COL_INV: return launch_col_t<NX, W, COL_INV>(p, batch, st);
    case COL_STEP: return launch_col_t<NX, W, COL_STEP>(p, batch, st);
    case COL_PRO: return launch_col_t<NX, W, COL_PRO>(p, batch, st);
    }
    return (int)cudaErrorInvalidValue;
}

int launch_col(int nx, int mode, const ColParams &p, int batch, cudaStream_t st)
{
    switch (nx) {
    case 256: return launch_col_n<256, 4>(mode, p, batch, st);
    case 512: return launch_col_n<512, 4>(mode, p, batch, st);
    case 1024: return launch_col_n<1024, 4>(mode, p, batch, st);
    case 2048: return launch_col_n<2048, 4>(mode, p, batch, st);
    case 4096:
        return col_tile_width(4096) == 2 ? launch_col_n<4096, 2>(mode, p, batch, st) : launch_col_n<4096, 4>(mode, p, batch, st);
    case 8192:
        return col_tile_width(8192) == 1 ? launch_col_n<8192, 1>(mode, p, batch, st) : launch_col_n<8192, 2>(mode, p, batch, st);
    case 16384: return launch_col_n<16384, 1>(mode, p, batch, st);
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace xfb
