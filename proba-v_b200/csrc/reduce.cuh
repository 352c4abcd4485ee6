// reduce.cuh -- fixed-order (deterministic, atomic-free) reduction of per-CTA partial sums, shared by the weight-gradient
// kernels (wgrad_tc.cu, resblock_tc.cu).  A block of NW warps owns 128 consecutive outputs: warp y sums the partial
// rows y, y+NW, y+2NW, ... with float4 loads (512 contiguous bytes per row and warp, independent chains), the per-warp
// sums are then combined through shared memory in a fixed order.  The
// summation order depends only on `nrows`, never on scheduling, so a step is bit-reproducible.
#pragma once
#include <cuda_runtime.h>

namespace pv {

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// Sum over r < nrows of the float4 at src[rowoff(r) + col4*4 .. +3].  All NW*32 threads of the block must call; the result
// is returned to the threads of warp 0 (threadIdx.x < 32, col4 = col4_base + threadIdx.x); `sm` is a [NW][32] float4
// scratch.  Warp y owns rows y, y+NW, y+2NW, ...; with NW = 32 a 148-row reduction is at most five independent loads per
// thread, all in flight at once (the 8-warp version spent ~5 dependent L2 round trips per launch, 15-25 us measured).
template <int NW, typename RowOff>
__device__ __forceinline__ float4 block_rowsum4(const float* __restrict__ src, int nrows, RowOff rowoff, int col4_base, bool col_ok, float4* sm) {
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    if (col_ok) {
        const size_t c = (size_t)(col4_base + x) * 4;
        int r = y;
        for (; r + 3 * NW < nrows; r += 4 * NW) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(src + rowoff(r) + c));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(src + rowoff(r + NW) + c));
            const float4 v2 = __ldg(reinterpret_cast<const float4*>(src + rowoff(r + 2 * NW) + c));
            const float4 v3 = __ldg(reinterpret_cast<const float4*>(src + rowoff(r + 3 * NW) + c));
            a0 = f4add(a0, v0); a1 = f4add(a1, v1); a2 = f4add(a2, v2); a3 = f4add(a3, v3);
        }
        float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0, t2 = t0;       // up to three leftover rows, loaded together
        if (r < nrows) t0 = __ldg(reinterpret_cast<const float4*>(src + rowoff(r) + c));
        if (r + NW < nrows) t1 = __ldg(reinterpret_cast<const float4*>(src + rowoff(r + NW) + c));
        if (r + 2 * NW < nrows) t2 = __ldg(reinterpret_cast<const float4*>(src + rowoff(r + 2 * NW) + c));
        a0 = f4add(a0, t0); a1 = f4add(a1, t1); a2 = f4add(a2, t2);
    }
    sm[y * 32 + x] = f4add(f4add(a0, a1), f4add(a2, a3));
    __syncthreads();
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y == 0) {
#pragma unroll
        for (int w = 0; w < NW; w += 4) {
            const float4 p01 = f4add(sm[w * 32 + x], sm[(w + 1) * 32 + x]), p23 = f4add(sm[(w + 2) * 32 + x], sm[(w + 3) * 32 + x]);
            s = f4add(s, f4add(p01, p23));
        }
    }
    __syncthreads();
    return s;
}

}  // namespace pv
