// probe_f32x2.cu -- does the packed fp32 pipe of sm_100a (FFMA2 / FADD2 / FMUL2) buy issue slots or flops?
// Measures thread-level flop/s of (a) scalar fmaf chains and (b) __ffma2_rn chains with the same number of
// independent accumulators per thread, plus a complex-multiply-add mix as used by the FFT butterflies.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_f32x2 probe_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ACC = 8;      // independent float2 accumulators per thread
constexpr int ITERS = 4096;

__global__ void k_scalar(float2 *out, float2 a, float2 b)
{
    float2 v[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) v[i] = make_float2(threadIdx.x + i, blockIdx.x - i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            v[i].x = fmaf(v[i].x, a.x, b.x);
            v[i].y = fmaf(v[i].y, a.y, b.y);
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < ACC; ++i) { s.x += v[i].x; s.y += v[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_packed(float2 *out, float2 a, float2 b)
{
    float2 v[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) v[i] = make_float2(threadIdx.x + i, blockIdx.x - i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) v[i] = __ffma2_rn(v[i], a, b);
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < ACC; ++i) s = __fadd2_rn(s, v[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// butterfly-like mix: v = v * w (complex) then radix-2 add/sub with the neighbour
__global__ void k_cmix_scalar(float2 *out, float2 w)
{
    float2 v[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) v[i] = make_float2(threadIdx.x + i, blockIdx.x - i);
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            const float2 z = v[i];
            v[i] = make_float2(z.x * w.x - z.y * w.y, z.x * w.y + z.y * w.x);
        }
#pragma unroll
        for (int i = 0; i < ACC; i += 2) {
            const float2 a = v[i], b = v[i + 1];
            v[i] = make_float2(a.x + b.x, a.y + b.y);
            v[i + 1] = make_float2(a.x - b.x, a.y - b.y);
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < ACC; ++i) { s.x += v[i].x; s.y += v[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_cmix_packed(float2 *out, float2 w)
{
    float2 v[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) v[i] = make_float2(threadIdx.x + i, blockIdx.x - i);
    const float2 wxx = make_float2(w.x, w.x), wyy = make_float2(-w.y, w.y);
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            const float2 z = v[i];
            // (z.x wx - z.y wy, z.y wx + z.x wy) = z * (wx,wx) + swap(z) * (-wy, wy)
            const float2 zs = make_float2(z.y, z.x);
            v[i] = __ffma2_rn(zs, wyy, __fmul2_rn(z, wxx));
        }
#pragma unroll
        for (int i = 0; i < ACC; i += 2) {
            const float2 a = v[i], b = v[i + 1];
            v[i] = __fadd2_rn(a, b);
            v[i + 1] = __fadd2_rn(a, make_float2(-b.x, -b.y));
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < ACC; ++i) s = __fadd2_rn(s, v[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main()
{
    const int blocks = 148 * 8, threads = 256;
    float2 *out;
    CK(cudaMalloc(&out, sizeof(float2) * blocks * threads));
    const float2 a = make_float2(0.999f, 1.001f), b = make_float2(1e-3f, -1e-3f), w = make_float2(0.8f, 0.6f);
    const double fl = 2.0 * 2 * ACC * (double)ITERS * blocks * threads;   // flops of the fma kernels
    float ms;
    ms = time_ms([&] { k_scalar<<<blocks, threads>>>(out, a, b); });
    printf("scalar FFMA      : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms * 1e-9);
    ms = time_ms([&] { k_packed<<<blocks, threads>>>(out, a, b); });
    printf("packed FFMA2     : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms * 1e-9);
    ms = time_ms([&] { k_cmix_scalar<<<blocks, threads>>>(out, w); });
    printf("cmul+bfly scalar : %.3f ms\n", ms);
    ms = time_ms([&] { k_cmix_packed<<<blocks, threads>>>(out, w); });
    printf("cmul+bfly packed : %.3f ms\n", ms);
    CK(cudaDeviceSynchronize());
    return 0;
}
