// probe_tma_tile.cu -- how fast can one SM move narrow column tiles (W complex columns x many rows) of a
// row-major [NX][pitch] complex array with TMA (cp.async.bulk.tensor.2d), compared with plain LDG/STG?
// This is the access pattern of the transposed pass (K-COL).  One elected thread per CTA drives the TMA
// engine through a ring of shared-memory stages: load chunk -> store chunk to a second array.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_tma_tile probe_tma_tile.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled get_encode()
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    return (EncodeTiled)fn;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int x, int y, const void *src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(x), "r"(y), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// chunk = CR rows x W columns (BOXR rows per TMA op).  Chunks are numbered over (column tile, row chunk).
template <int W, int CR, int BOXR, int STAGES>
__global__ void __launch_bounds__(128) k_tma_copy(const __grid_constant__ CUtensorMap src, const __grid_constant__ CUtensorMap dst,
                                                  int nx, int ntiles)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[STAGES];
    constexpr int CHUNK_BYTES = CR * W * 8;
    const int chunks_per_tile = nx / CR;
    const int nchunks = ntiles * chunks_per_tile;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    // this CTA's chunks: c = blockIdx.x, += gridDim.x
    int issued = 0, done = 0;
    uint32_t phase_bits = 0;
    const int first = blockIdx.x, stride = gridDim.x;
    const int mine = (nchunks - first + stride - 1) / stride;
    auto issue = [&](int n) {
        const int c = first + n * stride;
        const int tile = c / chunks_per_tile, rc = c % chunks_per_tile;
        const int s = n % STAGES;
        mbar_expect_tx(&full[s], CHUNK_BYTES);
        for (int b = 0; b < CR / BOXR; ++b)
            tma_load_2d(smem + (size_t)s * CHUNK_BYTES + (size_t)b * BOXR * W * 8, &src, tile * W, rc * CR + b * BOXR, &full[s]);
    };
    for (; issued < STAGES && issued < mine; ++issued) issue(issued);
    for (; done < mine; ++done) {
        const int s = done % STAGES;
        mbar_wait(&full[s], (phase_bits >> s) & 1);
        phase_bits ^= 1u << s;
        const int c = first + done * stride;
        const int tile = c / chunks_per_tile, rc = c % chunks_per_tile;
        for (int b = 0; b < CR / BOXR; ++b)
            tma_store_2d(&dst, tile * W, rc * CR + b * BOXR, smem + (size_t)s * CHUNK_BYTES + (size_t)b * BOXR * W * 8);
        tma_commit();
        if (issued < mine) {
            tma_wait_read<0>();          // the stage being refilled is the one just stored when STAGES == 1; keep it simple
            issue(issued++);
        }
    }
    tma_wait_read<0>();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// LSU baseline: 512 threads copy an NX x W tile, 8 bytes per thread per access, 16 accesses in flight
template <int W>
__global__ void __launch_bounds__(512) k_lsu_copy(const float2 *src, float2 *dst, int nx, int pitch)
{
    const int j0 = blockIdx.x * W;
    const int c = threadIdx.x % W, t = threadIdx.x / W;
    const int rows_per_pass = 512 / W;
    for (int r0 = 0; r0 < nx; r0 += rows_per_pass * 16) {
        float2 v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = src[(size_t)(r0 + t + k * rows_per_pass) * pitch + j0 + c];
#pragma unroll
        for (int k = 0; k < 16; ++k) dst[(size_t)(r0 + t + k * rows_per_pass) * pitch + j0 + c] = v[k];
    }
}

static CUtensorMap make_map(EncodeTiled enc, void *base, int nx, int pitch, int w, int boxr)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)nx};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 8};
    cuuint32_t box[2] = {(cuuint32_t)w, (cuuint32_t)boxr};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (w=%d boxr=%d)\n", (int)r, w, boxr); exit(1); }
    return m;
}

template <typename F>
static float time_ms(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) f();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 3;
}

template <int W, int CR, int BOXR, int STAGES>
static void run_tma(EncodeTiled enc, float2 *src, float2 *dst, int nx, int pitch, int ctas_per_sm)
{
    CUtensorMap ms = make_map(enc, src, nx, pitch, W, BOXR), md = make_map(enc, dst, nx, pitch, W, BOXR);
    const int smem = STAGES * CR * W * 8;
    CK(cudaFuncSetAttribute(k_tma_copy<W, CR, BOXR, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int ntiles = (pitch - 4) / W;      // whole tiles only
    CK(cudaMemset(dst, 0, sizeof(float2) * (size_t)nx * pitch));
    const float t = time_ms([&] { k_tma_copy<W, CR, BOXR, STAGES><<<148 * ctas_per_sm, 128, smem>>>(ms, md, nx, ntiles); });
    // verify a few elements
    float2 a, b;
    const size_t probe = (size_t)(nx - 3) * pitch + (size_t)ntiles * W - 1;
    CK(cudaMemcpy(&a, src + probe, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&b, dst + probe, 8, cudaMemcpyDeviceToHost));
    const double bytes = 2.0 * 8.0 * nx * (double)ntiles * W;
    printf("TMA  W=%d chunk=%5d rows box=%3d stages=%d ctas/sm=%d smem=%6d : %7.3f ms  %7.1f GB/s (r+w)  %s\n", W, CR, BOXR, STAGES,
           ctas_per_sm, smem, t, bytes / t * 1e-6, (a.x == b.x && a.y == b.y) ? "ok" : "MISMATCH");
}

template <int W>
static void run_lsu(float2 *src, float2 *dst, int nx, int pitch)
{
    const int ntiles = (pitch - 4) / W;
    const float t = time_ms([&] { k_lsu_copy<W><<<ntiles, 512>>>(src, dst, nx, pitch); });
    const double bytes = 2.0 * 8.0 * nx * (double)ntiles * W;
    printf("LSU  W=%d (512 thr, 16 x 8 B in flight per thread)               : %7.3f ms  %7.1f GB/s (r+w)\n", W, t, bytes / t * 1e-6);
}

int main()
{
    const int nx = 8192, pitch = 4100;
    float2 *src, *dst;
    CK(cudaMalloc(&src, sizeof(float2) * (size_t)nx * pitch));
    CK(cudaMalloc(&dst, sizeof(float2) * (size_t)nx * pitch));
    {
        float2 *h = (float2 *)malloc(sizeof(float2) * (size_t)nx * pitch);
        for (size_t i = 0; i < (size_t)nx * pitch; ++i) h[i] = make_float2((float)(i % 9973), (float)(i % 7919));
        CK(cudaMemcpy(src, h, sizeof(float2) * (size_t)nx * pitch, cudaMemcpyHostToDevice));
        free(h);
    }
    EncodeTiled enc = get_encode();
    run_lsu<2>(src, dst, nx, pitch);
    run_lsu<4>(src, dst, nx, pitch);
    run_lsu<8>(src, dst, nx, pitch);
    // W = 2 (16-byte box rows): what K-COL needs at 8192
    run_tma<2, 2048, 256, 2>(enc, src, dst, nx, pitch, 1);
    run_tma<2, 2048, 256, 2>(enc, src, dst, nx, pitch, 2);
    run_tma<2, 2048, 256, 3>(enc, src, dst, nx, pitch, 2);
    run_tma<2, 1024, 256, 4>(enc, src, dst, nx, pitch, 2);
    run_tma<2, 1024, 256, 4>(enc, src, dst, nx, pitch, 3);
    run_tma<2, 2048, 128, 2>(enc, src, dst, nx, pitch, 2);
    run_tma<4, 1024, 256, 2>(enc, src, dst, nx, pitch, 1);
    run_tma<4, 1024, 256, 2>(enc, src, dst, nx, pitch, 2);
    run_tma<4, 1024, 256, 3>(enc, src, dst, nx, pitch, 2);
    run_tma<8, 512, 256, 3>(enc, src, dst, nx, pitch, 2);
    run_tma<16, 256, 256, 3>(enc, src, dst, nx, pitch, 2);
    CK(cudaDeviceSynchronize());
    return 0;
}
