// probe_p2p_copy.cu -- copy-engine peer bandwidth between two B200s: 1-D vs 2-D copies, one vs both directions,
// idle vs under an HBM-heavy kernel.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe_p2p_copy probe_p2p_copy.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void hog(float4 *a, const float4 *b, size_t n, int reps)
{
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = b[i];
}

int main()
{
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    const size_t bytes = 256ull << 20;
    char *buf[2][2];
    cudaStream_t st[2][4];
    cudaEvent_t e0[2], e1[2];
    float4 *ha[2], *hb[2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        for (int k = 0; k < 2; ++k) CK(cudaMalloc(&buf[d][k], bytes));
        for (int k = 0; k < 4; ++k) CK(cudaStreamCreateWithFlags(&st[d][k], cudaStreamNonBlocking));
        CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d]));
        CK(cudaMalloc(&ha[d], 1ull << 30)); CK(cudaMalloc(&hb[d], 1ull << 30));
    }
    auto run = [&](const char *name, bool both, bool twod, int nstreams, bool hogging) -> int {
        for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
        if (hogging)
            for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); hog<<<148 * 4, 512, 0, st[d][3]>>>(ha[d], hb[d], (1ull << 30) / 16, 6); }
        const int ndir = both ? 2 : 1;
        for (int d = 0; d < ndir; ++d) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], st[d][0])); }
        const int reps = 4;
        for (int r = 0; r < reps; ++r)
            for (int d = 0; d < ndir; ++d) {
                CK(cudaSetDevice(d));
                const size_t piece = bytes / nstreams;
                for (int s = 0; s < nstreams; ++s) {
                    char *dst = buf[1 - d][1] + s * piece, *src = buf[d][0] + s * piece;
                    if (twod) CK(cudaMemcpy2DAsync(dst, piece / 4, src, piece / 4, piece / 4 - 4096, 4, cudaMemcpyDeviceToDevice, st[d][s]));
                    else CK(cudaMemcpyAsync(dst, src, piece, cudaMemcpyDeviceToDevice, st[d][s]));
                }
            }
        float worst = 0.f;
        for (int d = 0; d < ndir; ++d) {
            CK(cudaSetDevice(d));
            for (int s = 1; s < nstreams; ++s) {
                cudaEvent_t j; CK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming));
                CK(cudaEventRecord(j, st[d][s])); CK(cudaStreamWaitEvent(st[d][0], j, 0));
            }
            CK(cudaEventRecord(e1[d], st[d][0]));
        }
        for (int d = 0; d < ndir; ++d) {
            CK(cudaSetDevice(d));
            CK(cudaEventSynchronize(e1[d]));
            float ms; CK(cudaEventElapsedTime(&ms, e0[d], e1[d]));
            if (ms > worst) worst = ms;
        }
        printf("%-44s %8.1f GB/s per direction\n", name, reps * (double)bytes / worst * 1e-6);
        for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
        return 0;
    };
    run("1-D, one direction, 1 stream", false, false, 1, false);
    run("1-D, both directions, 1 stream", true, false, 1, false);
    run("1-D, both directions, 4 streams", true, false, 4, false);
    run("2-D (4 rows), both directions, 1 stream", true, true, 1, false);
    run("2-D (4 rows), both directions, 4 streams", true, true, 4, false);
    run("1-D, both directions, 1 stream, HBM busy", true, false, 1, true);
    run("1-D, both directions, 4 streams, HBM busy", true, false, 4, true);
    run("2-D, both directions, 4 streams, HBM busy", true, true, 4, true);
    return 0;
}
