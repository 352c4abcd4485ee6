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
// scratch.  Warp y owns rows y, y+NW, y+2NW, ...  and walks them in chunks of up to 20: all loads of a chunk are issued
// before the first add, so a 148-row reduction by 8 warps is ONE memory round trip with 19 loads in flight per thread
// (a 4-way unrolled loop spent five dependent round trips, 15-25 us per launch).
template <int NW, typename RowOff>
__device__ __forceinline__ float4 block_rowsum4(const float* __restrict__ src, int nrows, RowOff rowoff, int col4_base, bool col_ok, float4* sm) {
    constexpr int CH = 20;
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok) {
        const size_t c = (size_t)(col4_base + x) * 4;
        for (int r0 = y; r0 < nrows; r0 += CH * NW) {
            float4 v[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int r = r0 + i * NW;
                v[i] = r < nrows ? __ldg(reinterpret_cast<const float4*>(src + rowoff(r) + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 a0 = v[0], a1 = v[1], a2 = v[2], a3 = v[3];
#pragma unroll
            for (int i = 4; i < CH; i += 4) { a0 = f4add(a0, v[i]); a1 = f4add(a1, v[i + 1]); a2 = f4add(a2, v[i + 2]); a3 = f4add(a3, v[i + 3]); }
            acc = f4add(acc, f4add(f4add(a0, a1), f4add(a2, a3)));
        }
    }
    sm[y * 32 + x] = acc;
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
