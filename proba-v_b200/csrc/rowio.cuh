// rowio.cuh -- coalesced global-memory I/O for epilogues whose threads own one 128-byte activation row each.
//
// After tcgen05.ld (32x32b) lane L of a warp holds row L of a [32 rows x 32 fp32] block, and the 32 rows are contiguous
// in global memory.  Letting every lane read / write its own row with LDG.128 / STG.128 makes each warp instruction
// touch 32 different 128-byte lines: 32 L1 tag look-ups per instruction, ~2000 LSU cycles per 128-row tile for a
// load + store epilogue (measured with ncu on rowconv3_tc_kernel: 31.6 tag requests per instruction, l1tex at 72 % of
// peak while the tensor pipe sat at 37 %).  These helpers go through a small per-warp shared-memory scratch instead, so
// every global instruction covers four whole rows (512 contiguous bytes, 4 tag look-ups).  The scratch holds 16 rows
// (2 KB per warp, 128-byte aligned) in a 16-byte-chunk XOR swizzle (chunk ^= row & 7) that makes both the row-wise and
// the chunk-wise accesses bank-conflict free; a 32-row block takes two rounds.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace pv {

constexpr int ROWIO_SCRATCH_BYTES = 2048;      // per warp

// Coalesced prefetch of the warp's 32 rows (g0 = address of row 0; rowmask bit r = row r may be read): lane holds the
// 16-byte chunk (lane & 7) of rows i*4 + (lane >> 3), i = 0..7.  Pair with rowio_rows_from_chunks().
__device__ __forceinline__ void rowio_ldg_chunks(const float* __restrict__ g0, uint32_t rowmask, float4 (&t)[8]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3);
        t[i] = ((rowmask >> row) & 1u) ? __ldg(reinterpret_cast<const float4*>(g0 + (size_t)row * 32) + (lane & 7)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// chunk-distributed registers (rowio_ldg_chunks) -> row-distributed: lane L receives its row L as eight float4
__device__ __forceinline__ void rowio_rows_from_chunks(const float4 (&t)[8], float4 (&row)[8], uint8_t* sc) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = i * 4 + (lane >> 3);
            *reinterpret_cast<float4*>(sc + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4)) = t[h * 4 + i];
        }
        __syncwarp();
        if ((lane >> 4) == h) {
            const int rr = lane & 15;
#pragma unroll
            for (int g4 = 0; g4 < 8; ++g4) row[g4] = *reinterpret_cast<const float4*>(sc + rr * 128 + ((g4 ^ (rr & 7)) << 4));
        }
        __syncwarp();
    }
}

// row-distributed registers (lane L holds o[0..31] = its row) -> coalesced global stores of the rows selected by rowmask
__device__ __forceinline__ void rowio_store_rows(float* __restrict__ g0, const float (&o)[32], uint32_t rowmask, uint8_t* sc) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if ((lane >> 4) == h) {
            const int rr = lane & 15;
#pragma unroll
            for (int g4 = 0; g4 < 8; ++g4)
                *reinterpret_cast<float4*>(sc + rr * 128 + ((g4 ^ (rr & 7)) << 4)) = make_float4(o[4 * g4], o[4 * g4 + 1], o[4 * g4 + 2], o[4 * g4 + 3]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = i * 4 + (lane >> 3), row = h * 16 + rr;
            const float4 v = *reinterpret_cast<const float4*>(sc + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
            if ((rowmask >> row) & 1u) reinterpret_cast<float4*>(g0 + (size_t)row * 32)[lane & 7] = v;
        }
        __syncwarp();
    }
}

}  // namespace pv
