// wgrad_reduce.cuh -- the fixed-order reductions of per-CTA weight-gradient partials (reduce.cuh) as device functions,
// so that they can run either right after their producer (selftests, fp32 twins) or all together in ONE deferred launch
// at the end of the backward pass (deferred_reduce.cu): 29 small latency-bound launches per step become one
// bandwidth-bound launch.  Every body is written for a block of 256 threads and a block index local to its job.
#pragma once
#include "reduce.cuh"
#include "rows.h"

namespace pv {

// where a tensor-core weight-gradient partial [group][128][32] lands in dweff (see wgrad_tc.cu for the three modes)
struct WgradScatter {
    float* dw; int dw_cols;
    float* db;
    int dwr0[MAX_TAPS], dwc0[MAX_TAPS];
};

struct ReduceJob {
    int kind;                          // 0 rowwgrad_tc, 1 resfront_bwd_weight, 2 first_conv_pr_wgrad, 3 skip2d
    int nblocks;
    const float* partials; const float* dbp; int ncta;
    int ngroup, mode, nbias;           // kind 0
    WgradScatter sc;                   // kind 0 (kind 2: dw / db only)
    float *dwd, *dwe, *dbe, *dbd;      // kind 1
    float *o0, *o1, *o2, *o3, *o4, *o5; int S, C;   // kind 3: dw1, dw2, dw3, db1, db2, db3
};

constexpr int MAX_REDUCE_JOBS = 40;
struct ReduceJobs { int njobs; int block_start[MAX_REDUCE_JOBS + 1]; ReduceJob jobs[MAX_REDUCE_JOBS]; };

inline int wgrad_reduce_blocks(int ngroup, int nbias) { return ngroup * 32 + (nbias * 8 + 31) / 32; }

// mode 0: conv3 (group = (dt,dh), m = q*32+ci, n = co); mode 1: wide x (group g, m = channel in group -> K index
// g*128+m, n = co); mode 2: wide gz (m = co in group, n = ci).  Blocks [0, total/128) own 128 weight-gradient outputs
// each, the remaining blocks the bias sums.
__device__ __forceinline__ void wgrad_reduce_body(int blk, const float* __restrict__ partials, const float* __restrict__ dbp, int ncta,
                                                  int ngroup, int mode, const WgradScatter& p, int nbias, float4* sm) {
    const int total = ngroup * 4096;
    const int nmain = total / 128;
    if (blk < nmain) {
        // conv3 partials are [group][128 = 4 x 32][32] with only three of the four 32-row quarters holding dw taps: the blocks
        // of the fourth quarter have nothing to sum (and the producer does not store it)
        if (mode == 0 && (blk & 31) >= 24) return;
        const float4 s = block_rowsum4<8>(partials, ncta, [total](int r) { return (size_t)r * total; }, blk * 32, true, sm);
        if (threadIdx.x >= 32) return;
        const float v[4] = {s.x, s.y, s.z, s.w};
        const int idx0 = (blk * 32 + threadIdx.x) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = idx0 + e;
            const int g = idx / 4096, m = (idx / 32) % 128, n = idx % 32;
            if (mode == 0) {
                const int qd = m / 32, ci = m % 32;
                if (qd < 3) { const int tap = g * 3 + qd; p.dw[(size_t)(p.dwr0[tap] + ci) * p.dw_cols + p.dwc0[tap] + n] = v[e]; }
            } else if (mode == 1) {
                const int k = g * 128 + m, tap = k / 32;
                p.dw[(size_t)(p.dwr0[tap] + k % 32) * p.dw_cols + p.dwc0[tap] + n] = v[e];
            } else {
                p.dw[(size_t)(p.dwr0[0] + n) * p.dw_cols + p.dwc0[0] + g * 128 + m] = v[e];
            }
        }
    } else {
        const int bb = blk - nmain, w = nbias * 32;
        const bool ok = (bb * 32 + (int)(threadIdx.x & 31)) * 4 < w;
        const float4 s = block_rowsum4<8>(dbp, ncta * 4, [w](int r) { return (size_t)r * w; }, bb * 32, ok, sm);
        if (threadIdx.x < 32 && ok && p.db) {
            float* o = p.db + (bb * 32 + threadIdx.x) * 4;
            o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = s.w;
        }
    }
}

// epilogue groups (of four warps) of resfront_bwd_weight_kernel and the per-CTA bias-gradient partial they write:
// [group][256 (dbe)] then [group x 4 warps][32 (dbd)]
#ifndef PV_RESBW_GROUPS
#define PV_RESBW_GROUPS 2
#endif
constexpr int RESBW_GROUPS = PV_RESBW_GROUPS;
constexpr int RESBW_DBP = RESBW_GROUPS * 256 + RESBW_GROUPS * 4 * 32;
constexpr int RESFRONT_REDUCE_BLOCKS = 131;
// dWd [256][32] (= dweff of decConv), dWe [32][256] (= dweff of expConv), dbe [256], dbd [32] from the per-CTA partials
// [cta][4][128][32] and [cta][RESBW_DBP].  Blocks [0,128): weight gradients; 128,129: dbe; 130: dbd.
__device__ __forceinline__ void resfront_reduce_body(int b, const float* __restrict__ partials, const float* __restrict__ dbp, int ncta,
                                                     float* __restrict__ dwd, float* __restrict__ dwe, float* __restrict__ dbe,
                                                     float* __restrict__ dbd, float4* sm) {
    const int x = threadIdx.x & 31;
    if (b < 128) {
        const float4 s = block_rowsum4<8>(partials, ncta, [](int r) { return (size_t)r * 16384; }, b * 32, true, sm);
        if (threadIdx.x >= 32) return;
        const float v[4] = {s.x, s.y, s.z, s.w};
        const int idx0 = (b * 32 + x) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = idx0 + e;
            const int g = idx / 4096, m = (idx / 32) % 128, n = idx % 32;
            if (g < 2) dwd[(size_t)(g * 128 + m) * 32 + n] = v[e];                 // [ch][co]
            else dwe[(size_t)n * 256 + (g - 2) * 128 + m] = v[e];                  // [ci][ch]
        }
    } else if (b < 130) {       // dbe: both epilogue groups of every CTA
        const float4 s = block_rowsum4<8>(dbp, RESBW_GROUPS * ncta, [](int r) { return (size_t)(r / RESBW_GROUPS) * RESBW_DBP + (r % RESBW_GROUPS) * 256; }, (b - 128) * 32, true, sm);
        if (threadIdx.x < 32) { float* o = dbe + ((b - 128) * 32 + x) * 4; o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = s.w; }
    } else {                    // dbd: eight epilogue warps of every CTA
        const bool ok = x < 8;
        constexpr int NWD = RESBW_GROUPS * 4;      // epilogue warps per CTA
        const float4 s = block_rowsum4<8>(dbp, NWD * ncta, [](int r) { return (size_t)(r / NWD) * RESBW_DBP + RESBW_GROUPS * 256 + (r % NWD) * 32; }, 0, ok, sm);
        if (threadIdx.x < 32 && ok) { float* o = dbd + x * 4; o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = s.w; }
    }
}

constexpr int FIRST_CONV_REDUCE_BLOCKS = 7;
// mainConv1: [cta][28][32] partials (27 taps + bias row)
__device__ __forceinline__ void first_conv_reduce_body(int blk, const float* __restrict__ partials, int ncta, float* __restrict__ dw,
                                                       float* __restrict__ db, float4* sm) {
    const float4 s = block_rowsum4<8>(partials, ncta, [](int r) { return (size_t)r * (28 * 32); }, blk * 32, true, sm);
    if (threadIdx.x >= 32) return;
    const int i = (blk * 32 + threadIdx.x) * 4;
    float* o = i < 27 * 32 ? dw + i : db + (i - 27 * 32);
    o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = s.w;
}

// skip2d: per-patch partial vector { dW1[9C], dW2[9CC], dW3[9CC], db1[C], db2[C], db3[C] } padded to a multiple of 4
__host__ __device__ inline int skip2d_npart(int C) { return ((9 * C + 18 * C * C + 3 * C + 3) / 4) * 4; }
inline int skip2d_reduce_blocks(int C) { return (skip2d_npart(C) + 127) / 128; }
__device__ __forceinline__ void skip2d_reduce_body(int blk, const float* __restrict__ partials, int B, int C, float* dw1, float* dw2,
                                                   float* dw3, float* db1, float* db2, float* db3, float4* sm) {
    const int np = skip2d_npart(C);
    const bool ok = (blk * 32 + (int)(threadIdx.x & 31)) * 4 < np;
    const float4 s = block_rowsum4<8>(partials, B, [np](int r) { return (size_t)r * np; }, blk * 32, ok, sm);
    if (threadIdx.x >= 32 || !ok) return;
    const float v[4] = {s.x, s.y, s.z, s.w};
    const int n1 = 9 * C, n2 = 9 * C * C;
    for (int e = 0; e < 4; ++e) {
        int i = (blk * 32 + threadIdx.x) * 4 + e;
        if (i < n1) { dw1[i] = v[e]; continue; }
        i -= n1;
        if (i < n2) { dw2[i] = v[e]; continue; }
        i -= n2;
        if (i < n2) { dw3[i] = v[e]; continue; }
        i -= n2;
        if (i < C) db1[i] = v[e];
        else if (i < 2 * C) db2[i - C] = v[e];
        else if (i < 3 * C) db3[i - 2 * C] = v[e];
    }
}

// Host side: a bump allocator over the trainer's partial arena plus the job list of the current backward pass.  A
// launcher that is handed a queue carves its partial region from it and appends a job instead of launching its own
// reduction; when the arena or the job table is full it falls back to the immediate reduction.
struct ReduceQueue {
    float* arena = nullptr; size_t arena_floats = 0, used = 0;
    ReduceJobs jobs;
    void reset(size_t reserve_floats) { used = reserve_floats; jobs.njobs = 0; jobs.block_start[0] = 0; }
    float* take(size_t floats) {
        floats = (floats + 63) / 64 * 64;
        if (!arena || used + floats > arena_floats || jobs.njobs >= MAX_REDUCE_JOBS) return nullptr;
        float* p = arena + used; used += floats; return p;
    }
    void push(const ReduceJob& j) {
        jobs.jobs[jobs.njobs] = j;
        jobs.block_start[jobs.njobs + 1] = jobs.block_start[jobs.njobs] + j.nblocks;
        ++jobs.njobs;
    }
};
int launch_deferred_reduce(ReduceQueue& q, cudaStream_t st);

}  // namespace pv
