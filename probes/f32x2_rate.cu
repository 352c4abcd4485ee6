// probes/f32x2_rate.cu -- issue / pipe rate of the packed FP32 instructions (FADD2 / FFMA2, PTX add/fma.rn.f32x2) against scalar
// FFMA on sm_100a.  One CTA of 1024 threads per SM, 8 independent chains per thread; reports FP32 lane-operations per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) rate(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a + i + threadIdx.x;
    unsigned long long ab;
    { float2 t = make_float2(a, b); ab = *reinterpret_cast<unsigned long long*>(&t); }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v = ((unsigned long long)__float_as_uint(x[i + 1]) << 32) | __float_as_uint(x[i]);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v) : "l"(ab));
                x[i] = __uint_as_float((uint32_t)v); x[i + 1] = __uint_as_float((uint32_t)(v >> 32));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v = ((unsigned long long)__float_as_uint(x[i + 1]) << 32) | __float_as_uint(x[i]);
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(ab));
                x[i] = __uint_as_float((uint32_t)v); x[i + 1] = __uint_as_float((uint32_t)(v >> 32));
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * 1024 + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 1024 * 4);
    const int iters = 4096;
    const char* names[3] = {"FFMA  (scalar)", "FFMA2 (fma.rn.f32x2)", "FADD2 (add.rn.f32x2)"};
    for (int m = 0; m < 3; ++m) {
        for (int rep = 0; rep < 2; ++rep) {
            if (m == 0) rate<0><<<148, 1024>>>(d, iters, 1.0001f, 0.5f);
            if (m == 1) rate<1><<<148, 1024>>>(d, iters, 1.0001f, 0.5f);
            if (m == 2) rate<2><<<148, 1024>>>(d, iters, 1.0001f, 0.5f);
            cudaDeviceSynchronize();
        }
        float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
        const double lane_ops = 1024.0 * 16 * iters;
        printf("%-24s %10.0f cycles  -> %.1f FP32 lane-ops / clk / SM  (%s)\n", names[m], cyc, lane_ops / cyc, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
