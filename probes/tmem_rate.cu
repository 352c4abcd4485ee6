// tmem_rate.cu -- TMEM <-> register bandwidth probe: 4 warps (one per lane quarter) stream 32x32b.x32 loads / stores.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o probes/tmem_rate probes/tmem_rate.cu
#include <cstdio>
#include <cstdlib>
#include "../proba-v_b200/csrc/tc_common.cuh"
using namespace pv::tc;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// mode 0: ld + wait each; 1: 4 lds then one wait; 2: ld, wait, st (in place), st wait at end of 4; 3: like the fused epilogue (ld,ld,wait,st)
__global__ void __launch_bounds__(128) tmem_kernel(int mode, int iters, long long* cycles, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t lane_base = slot + ((uint32_t)(warp * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (mode == 0) {
            for (int c = 0; c < 4; ++c) { uint32_t v[32]; tmem_ld32(lane_base + c * 32, v); tmem_ld_wait(); acc += __uint_as_float(v[i & 31]); }
        } else if (mode == 1) {
            uint32_t v0[32], v1[32], v2[32], v3[32];
            tmem_ld32(lane_base, v0); tmem_ld32(lane_base + 32, v1); tmem_ld32(lane_base + 64, v2); tmem_ld32(lane_base + 96, v3);
            tmem_ld_wait();
            acc += __uint_as_float(v0[i & 31]) + __uint_as_float(v1[i & 31]) + __uint_as_float(v2[i & 31]) + __uint_as_float(v3[i & 31]);
        } else if (mode == 2) {
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32]; tmem_ld32(lane_base + c * 32, v); tmem_ld_wait();
                for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(fmaxf(__uint_as_float(v[e]) + 1.f, 0.f));
                tmem_st32(lane_base + c * 32, v);
            }
            tmem_st_wait();
        } else {
            for (int c = 0; c < 4; ++c) {
                uint32_t e[32], v[32]; tmem_ld32(lane_base + c * 32, e); tmem_ld32(lane_base + 128 + c * 32, v); tmem_ld_wait();
                for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(e[k]) > 0.f ? v[k] : 0u;
                tmem_st32(lane_base + 128 + c * 32, v);
            }
            tmem_st_wait();
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 128 + threadIdx.x] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
    long long* d; float* sink; CK(cudaMalloc(&d, 148 * 8)); CK(cudaMalloc(&sink, 148 * 128 * 4));
    const int iters = 2048;
    const char* names[] = {"ld32 + wait, x4 (128 cols)", "4 x ld32 then one wait", "ld, relu, st in place (fwd epilogue)", "ld,ld,mask,st (bwd epilogue)"};
    const double kb[] = {64, 64, 64, 128};      // KB read from TMEM per iteration per CTA (128 lanes x cols x 4 B)
    for (int mode = 0; mode < 4; ++mode) {
        tmem_kernel<<<148, 128>>>(mode, iters, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d ERROR %s\n", mode, cudaGetErrorString(e)); return 1; }
        long long h[148]; CK(cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost));
        long long mx = 0; for (int i = 0; i < 148; ++i) if (h[i] > mx) mx = h[i];
        const double cyc = (double)mx / iters;
        printf("%-40s %.0f cycles per %3.0f KB read -> %.1f B/cycle/SM TMEM read\n", names[mode], cyc, kb[mode], kb[mode] * 1024 / cyc);
    }
    return 0;
}
