// reduce.cuh -- fixed-order (deterministic, atomic-free) reduction of per-CTA partial sums, shared by the weight-gradient
// kernels (wgrad_tc.cu, resblock_tc.cu).  A block of 256 threads owns 128 consecutive outputs: warp y sums the partial
// rows y, y+8, y+16, ... with float4 loads (512 contiguous bytes per row and warp, four independent chains so ~20 loads
// are in flight per thread), the eight per-warp sums are then combined through shared memory in a fixed tree.  The
// summation order depends only on `nrows`, never on scheduling, so a step is bit-reproducible.
#pragma once
#include <cuda_runtime.h>

namespace pv {

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// Sum over r < nrows of the float4 at src[rowoff(r) + col4*4 .. +3].  All 256 threads must call; the result is returned to
// the threads of warp 0 (threadIdx.x < 32, col4 = threadIdx.x); `sm` is a [8][32] float4 scratch.
template <typename RowOff>
__device__ __forceinline__ float4 block_rowsum4(const float* __restrict__ src, int nrows, RowOff rowoff, int col4_base, bool col_ok, float4* sm) {
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    if (col_ok) {
        const size_t c = (size_t)(col4_base + x) * 4;
        int r = y;
        for (; r + 24 < nrows; r += 32) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(src + rowoff(r) + c));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(src + rowoff(r + 8) + c));
            const float4 v2 = __ldg(reinterpret_cast<const float4*>(src + rowoff(r + 16) + c));
            const float4 v3 = __ldg(reinterpret_cast<const float4*>(src + rowoff(r + 24) + c));
            a0 = f4add(a0, v0); a1 = f4add(a1, v1); a2 = f4add(a2, v2); a3 = f4add(a3, v3);
        }
        for (; r < nrows; r += 8) a0 = f4add(a0, __ldg(reinterpret_cast<const float4*>(src + rowoff(r) + c)));
    }
    sm[y * 32 + x] = f4add(f4add(a0, a1), f4add(a2, a3));
    __syncthreads();
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y == 0) {
        const float4 s01 = f4add(sm[x], sm[32 + x]), s23 = f4add(sm[64 + x], sm[96 + x]);
        const float4 s45 = f4add(sm[128 + x], sm[160 + x]), s67 = f4add(sm[192 + x], sm[224 + x]);
        s = f4add(f4add(s01, s23), f4add(s45, s67));
    }
    __syncthreads();
    return s;
}

}  // namespace pv
