#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void tmem_st32(unsigned taddr, const float (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        :: "r"(taddr), "f"(r[0]),"f"(r[1]),"f"(r[2]),"f"(r[3]),"f"(r[4]),"f"(r[5]),"f"(r[6]),"f"(r[7]),"f"(r[8]),"f"(r[9]),"f"(r[10]),"f"(r[11]),"f"(r[12]),"f"(r[13]),"f"(r[14]),"f"(r[15]),"f"(r[16]),"f"(r[17]),"f"(r[18]),"f"(r[19]),"f"(r[20]),"f"(r[21]),"f"(r[22]),"f"(r[23]),"f"(r[24]),"f"(r[25]),"f"(r[26]),"f"(r[27]),"f"(r[28]),"f"(r[29]),"f"(r[30]),"f"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=f"(r[0]),"=f"(r[1]),"=f"(r[2]),"=f"(r[3]),"=f"(r[4]),"=f"(r[5]),"=f"(r[6]),"=f"(r[7]),"=f"(r[8]),"=f"(r[9]),"=f"(r[10]),"=f"(r[11]),"=f"(r[12]),"=f"(r[13]),"=f"(r[14]),"=f"(r[15]),"=f"(r[16]),"=f"(r[17]),"=f"(r[18]),"=f"(r[19]),"=f"(r[20]),"=f"(r[21]),"=f"(r[22]),"=f"(r[23]),"=f"(r[24]),"=f"(r[25]),"=f"(r[26]),"=f"(r[27]),"=f"(r[28]),"=f"(r[29]),"=f"(r[30]),"=f"(r[31]) : "r"(taddr) : "memory");
}
__global__ void __launch_bounds__(512) k(float *out, int reps)
{
    __shared__ unsigned base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"((unsigned)__cvta_generic_to_shared(&base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned base = base_s;
    // warp w: lanes 32*(w%4) .., columns (w/4)*64 ..
    const unsigned taddr = base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)((warp >> 2) * 64);
    float r[32], q[32];
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x * 100.f + i;
    float acc = 0.f;
    for (int it = 0; it < reps; ++it) {
        tmem_st32(taddr, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tmem_ld32(taddr, q);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) { acc += q[i]; r[i] = q[i] + 1.f; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(base) : "memory");
}
int main()
{
    float *d; cudaMalloc(&d, 148 * 512 * 4);
    k<<<148, 512>>>(d, 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s\n", cudaGetErrorString(e));
    float h[1024]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    // expected acc for thread t, reps=1: sum_i (t*100 + i) = 3200 t + 496
    int bad = 0; for (int t = 0; t < 512; ++t) if (h[t] != 3200.f * t + 496.f) ++bad;
    printf("mismatches: %d (h[1]=%f h[511]=%f)\n", bad, h[1], h[511]);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 2000;
    cudaEventRecord(e0); k<<<148, 512>>>(d, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double bytes = 148.0 * 512 * 128 * reps;   // per direction
    printf("%d st+ld round trips of 128 B/thread: %.3f ms -> %.1f GB/s per direction chip-wide, %.1f B/clk/SM at 1.9 GHz\n", reps, ms, bytes / ms * 1e-6, bytes / 148 / (ms * 1e-3) / 1.9e9);
    return 0;
}
