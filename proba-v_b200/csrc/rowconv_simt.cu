// rowconv_simt.cu -- fp32 CUDA-core kernels on the row layouts of rows.h.
//
// These are (a) the exact-fp32 implementation of every convolution of the tensor-core engine's graph, used for the
// layers that are not tensor-core shaped (mainConv1, Cin = 1) and as the on-device cross-check of each tcgen05
// kernel (pv_selftest), and (b) the layout glue between the PR trunk and the valid-conv tail.
// Reference semantics: Keras Conv3D inside TFA WeightNormalization (modelsTF.py:191-197), tf.pad REFLECT (:157-158).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include "wgrad_reduce.cuh"
#include "rows.h"

namespace pv {
namespace {

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}


__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// ------------------------------------------------------------------------------------------ forward / data gradient
template <int BN>
__global__ void __launch_bounds__(256) rowconv_vec_kernel(RowConvP p) {
    constexpr int BM = 128, BK = 16, TXN = BN / 4, TYN = 256 / TXN, TM = BM / TYN;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
    const long long M = (long long)p.B * p.og.nrows;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int kq = tid & 3;
    long long ibase[2];
    bool rvalid[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const long long m = m0 + (tid >> 2) + 64 * e;
        rvalid[e] = m < M;
        const long long b = rvalid[e] ? m / p.og.nrows : 0;
        const int r = p.og.row0 + (rvalid[e] ? (int)(m % p.og.nrows) : 0);
        ibase[e] = (p.in_lead + b * p.in_pstride + r) * p.xc + kq * 4;
    }
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < p.ntap; ++tap) {
        const long long toff = (long long)p.off[tap] * p.xc + p.c0[tap];
        for (int c0 = 0; c0 < p.kc; c0 += BK) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rvalid[e]) v = ldg4(p.x + ibase[e] + toff + c0);
                const int row = (tid >> 2) + 64 * e;
                As[kq * 4 + 0][row] = v.x; As[kq * 4 + 1][row] = v.y;
                As[kq * 4 + 2][row] = v.z; As[kq * 4 + 3][row] = v.w;
            }
            if (tid < BK * BN / 4) {
                const int kk = tid / (BN / 4), nn = (tid % (BN / 4)) * 4;
                float4 wv;
                if (p.w_kmajor) {
                    const float* wp = p.w + (long long)(p.wr0[tap] + n0 + nn) * p.w_cols + p.wc0[tap] + c0 + kk;
                    wv = make_float4(__ldg(wp), __ldg(wp + p.w_cols), __ldg(wp + 2 * p.w_cols), __ldg(wp + 3 * p.w_cols));
                } else {
                    wv = ldg4(p.w + (long long)(p.wr0[tap] + c0 + kk) * p.w_cols + p.wc0[tap] + n0 + nn);
                }
                *reinterpret_cast<float4*>(&Bs[kk][nn]) = wv;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float a[TM];
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
                    a[i] = av.x; a[i + 1] = av.y; a[i + 2] = av.z; a[i + 3] = av.w;
                }
                const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    acc[i][0] = fmaf(a[i], bv.x, acc[i][0]); acc[i][1] = fmaf(a[i], bv.y, acc[i][1]);
                    acc[i][2] = fmaf(a[i], bv.z, acc[i][2]); acc[i][3] = fmaf(a[i], bv.w, acc[i][3]);
                }
            }
            __syncthreads();
        }
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bv = ldg4(p.bias + n0 + tx * 4);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long m = m0 + ty * TM + i;
        if (m >= M) continue;
        const long long b = m / p.og.nrows;
        const int r = p.og.row0 + (int)(m % p.og.nrows);
        const long long yo = (p.og.lead + b * p.og.pstride + r) * p.n + n0 + tx * 4;
        float4 o = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
        if (p.residual) { const float4 q = ldg4(p.residual + yo); o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
        if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        if (p.relumask) {
            const float4 q = ldg4(p.relumask + yo);
            o.x = q.x > 0.f ? o.x : 0.f; o.y = q.y > 0.f ? o.y : 0.f; o.z = q.z > 0.f ? o.z : 0.f; o.w = q.w > 0.f ? o.w : 0.f;
        }
        if (!row_valid(p.og, r)) o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.round_tf32) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        *reinterpret_cast<float4*>(p.y + yo) = o;
    }
}

// ------------------------------------------------------------------------------------------ weight gradient
template <int BN>
__global__ void __launch_bounds__(256) rowwgrad_vec_kernel(RowWgradP p, int m_per_cta) {
    constexpr int BKO = 64, BMC = 32, TXN = BN / 4, TYN = 256 / TXN, TK = BKO / TYN;
    __shared__ __align__(16) float At[BMC][BKO + 4];
    __shared__ __align__(16) float Ys[BMC][BN];
    const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
    const long long M = (long long)p.B * p.og.nrows;
    const int k0 = blockIdx.x * BKO, n0 = blockIdx.y * BN;
    const int Ktot = p.ntap * p.kc;
    const long long mlo = (long long)blockIdx.z * m_per_cta;
    long long mhi = mlo + m_per_cta; if (mhi > M) mhi = M;
    const int kq = tid & 15;
    const int kg = k0 + kq * 4;
    const bool kvalid = kg < Ktot;
    const int tap = kvalid ? kg / p.kc : 0, ci = kvalid ? kg % p.kc : 0;
    const long long toff = (long long)p.off[tap] * p.xc + p.c0[tap] + ci;

    float acc[TK][4];
#pragma unroll
    for (int i = 0; i < TK; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    const bool do_bias = (blockIdx.x == 0) && (ty == 0) && p.db != nullptr;

    for (long long mc = mlo; mc < mhi; mc += BMC) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int row = (tid >> 4) + 16 * e;
            const long long m = mc + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < mhi && kvalid) {
                const long long b = m / p.og.nrows;
                const int r = p.og.row0 + (int)(m % p.og.nrows);
                v = ldg4(p.x + (p.in_lead + b * p.in_pstride + r) * p.xc + toff);
            }
            *reinterpret_cast<float4*>(&At[row][kq * 4]) = v;
        }
        for (int f = tid; f < BMC * BN / 4; f += 256) {
            const int row = f / (BN / 4), nn = (f % (BN / 4)) * 4;
            const long long m = mc + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < mhi) {
                const long long b = m / p.og.nrows;
                const int r = p.og.row0 + (int)(m % p.og.nrows);
                v = ldg4(p.gz + (p.og.lead + b * p.og.pstride + r) * p.n + n0 + nn);
            }
            *reinterpret_cast<float4*>(&Ys[row][nn]) = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int mm = 0; mm < BMC; ++mm) {
            float a[TK];
#pragma unroll
            for (int i = 0; i < TK; ++i) a[i] = At[mm][ty * TK + i];
            const float4 yv = *reinterpret_cast<const float4*>(&Ys[mm][tx * 4]);
#pragma unroll
            for (int i = 0; i < TK; ++i) {
                acc[i][0] = fmaf(a[i], yv.x, acc[i][0]); acc[i][1] = fmaf(a[i], yv.y, acc[i][1]);
                acc[i][2] = fmaf(a[i], yv.z, acc[i][2]); acc[i][3] = fmaf(a[i], yv.w, acc[i][3]);
            }
            if (do_bias) { bsum[0] += yv.x; bsum[1] += yv.y; bsum[2] += yv.z; bsum[3] += yv.w; }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TK; ++i) {
        const int k = k0 + ty * TK + i;
        if (k >= Ktot) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            atomicAdd(p.dw + (long long)(p.dwr0[k / p.kc] + k % p.kc) * p.dw_cols + p.dwc0[k / p.kc] + n0 + tx * 4 + j, acc[i][j]);
    }
    if (do_bias)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(p.db + n0 + tx * 4 + j, bsum[j]);
}

// ------------------------------------------------------------------------------------------ mainConv1 into PR
// One CTA per patch: the normalised LR patch (S x S x T, 17 KB at 22 x 22 x 9) is staged zero-padded in shared memory
// ('same' padding without predicates), the 27 x 32 weights and the bias sit next to it.  A thread owns two w-adjacent
// voxels x 8 output channels: per (dt, dh) line it reads 4 inputs and 3 x 2 weight float4s for 48 FMAs, so the kernel
// is FMA-bound instead of load-bound, and a warp stores 8 consecutive rows x 64 B per instruction.
__global__ void __launch_bounds__(256) first_conv_pr_kernel(const float* __restrict__ xn, const float* __restrict__ w,
                                                            const float* __restrict__ bias, int B, int S, int T,
                                                            float* __restrict__ y, RowGeom g, float* __restrict__ y_lo, int round_tf32) {
    pdl_grid_wait();
    extern __shared__ float fsm[];
    // blockIdx.y selects a chunk of temporal planes [tc0, tc0 + tcn) (one chunk per patch measured fastest at B = 128: 47 us vs 58 us for three)
    const int tchunk = (T + gridDim.y - 1) / gridDim.y, tc0 = blockIdx.y * tchunk, tcn = min(tchunk, T - tc0);
    if (tcn <= 0) return;
    const int Sp = S + 2, Tp = tcn + 2;
    float* xs = fsm;                              // [Tp][Sp][Sp + 1] zero padded (t, h, w); +1 column keeps the pair loads in range for odd S
    const int wp = Sp + 1;
    float* ws = xs + Tp * Sp * wp;                // [27][32], then bias [32]
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < Tp * Sp * wp; i += 256) xs[i] = 0.f;
    for (int i = threadIdx.x; i < 27 * 32 + 32; i += 256) ws[i] = i < 27 * 32 ? w[i] : bias[i - 27 * 32];
    __syncthreads();
    for (int i = threadIdx.x; i < S * S * T; i += 256) {          // xn is [h][w][t]
        const int t = i % T, ww = (i / T) % S, hh = i / (T * S);
        const int tl = t - tc0 + 1;                                // plane inside this chunk's padded window
        if (tl >= 0 && tl < Tp) xs[(tl * Sp + hh + 1) * wp + ww + 1] = xn[(size_t)b * S * S * T + i];
    }
    __syncthreads();
    const int pairs_w = (S + 1) / 2;
    const int items = tcn * S * pairs_w * 4;
    for (int it = threadIdx.x; it < items; it += 256) {
        const int cg = it & 3;
        int v = it >> 2;
        const int pw2 = v % pairs_w; v /= pairs_w;
        const int hh = v % S, tt = v / S;
        const int w0 = 2 * pw2;
        float a0[8], a1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a0[j] = a1[j] = ws[27 * 32 + cg * 8 + j];
#pragma unroll
        for (int dt = 0; dt < 3; ++dt)
#pragma unroll
            for (int dh = 0; dh < 3; ++dh) {
                const float* xp = xs + ((tt + dt) * Sp + hh + dh) * wp + w0;       // padded coordinates: tap (dt,dh,dw) of voxel w0 is xp[dw]
                const float x0 = xp[0], x1 = xp[1], x2 = xp[2], x3 = xp[3];
                const float* wr = ws + ((dt * 3 + dh) * 3) * 32 + cg * 8;
#pragma unroll
                for (int dw = 0; dw < 3; ++dw) {
                    const float4 wa = *reinterpret_cast<const float4*>(wr + dw * 32), wb = *reinterpret_cast<const float4*>(wr + dw * 32 + 4);
                    const float xa = dw == 0 ? x0 : (dw == 1 ? x1 : x2), xb = dw == 0 ? x1 : (dw == 1 ? x2 : x3);
                    a0[0] = fmaf(xa, wa.x, a0[0]); a0[1] = fmaf(xa, wa.y, a0[1]); a0[2] = fmaf(xa, wa.z, a0[2]); a0[3] = fmaf(xa, wa.w, a0[3]);
                    a0[4] = fmaf(xa, wb.x, a0[4]); a0[5] = fmaf(xa, wb.y, a0[5]); a0[6] = fmaf(xa, wb.z, a0[6]); a0[7] = fmaf(xa, wb.w, a0[7]);
                    a1[0] = fmaf(xb, wa.x, a1[0]); a1[1] = fmaf(xb, wa.y, a1[1]); a1[2] = fmaf(xb, wa.z, a1[2]); a1[3] = fmaf(xb, wa.w, a1[3]);
                    a1[4] = fmaf(xb, wb.x, a1[4]); a1[5] = fmaf(xb, wb.y, a1[5]); a1[6] = fmaf(xb, wb.z, a1[6]); a1[7] = fmaf(xb, wb.w, a1[7]);
                }
            }
        const long long oidx = (g.lead + (long long)b * g.pstride + (long long)(g.t0 + tc0 + tt) * g.plane + hh * g.pw + w0) * 32 + cg * 8;
        float* o = y + oidx;
        if (y_lo) {     // error-compensated engine: y = tf32(v), y_lo = v - tf32(v)
            float* ol = y_lo + oidx;
            float h0[8], l0[8], h1[8], l1[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float v0 = fmaxf(a0[j], 0.f), v1 = fmaxf(a1[j], 0.f);
                h0[j] = to_tf32(v0); l0[j] = v0 - h0[j]; h1[j] = to_tf32(v1); l1[j] = v1 - h1[j];
            }
            *reinterpret_cast<float4*>(o) = make_float4(h0[0], h0[1], h0[2], h0[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(h0[4], h0[5], h0[6], h0[7]);
            *reinterpret_cast<float4*>(ol) = make_float4(l0[0], l0[1], l0[2], l0[3]);
            *reinterpret_cast<float4*>(ol + 4) = make_float4(l0[4], l0[5], l0[6], l0[7]);
            if (w0 + 1 < S) {
                *reinterpret_cast<float4*>(o + 32) = make_float4(h1[0], h1[1], h1[2], h1[3]);
                *reinterpret_cast<float4*>(o + 36) = make_float4(h1[4], h1[5], h1[6], h1[7]);
                *reinterpret_cast<float4*>(ol + 32) = make_float4(l1[0], l1[1], l1[2], l1[3]);
                *reinterpret_cast<float4*>(ol + 36) = make_float4(l1[4], l1[5], l1[6], l1[7]);
            }
            continue;
        }
        if (round_tf32) {       // the output only feeds kind::tf32 MMAs, which truncate a raw fp32 operand: round to nearest here
#pragma unroll
            for (int j = 0; j < 8; ++j) { a0[j] = to_tf32(fmaxf(a0[j], 0.f)); a1[j] = to_tf32(fmaxf(a1[j], 0.f)); }
        }
        *reinterpret_cast<float4*>(o) = make_float4(fmaxf(a0[0], 0.f), fmaxf(a0[1], 0.f), fmaxf(a0[2], 0.f), fmaxf(a0[3], 0.f));
        *reinterpret_cast<float4*>(o + 4) = make_float4(fmaxf(a0[4], 0.f), fmaxf(a0[5], 0.f), fmaxf(a0[6], 0.f), fmaxf(a0[7], 0.f));
        if (w0 + 1 < S) {
            *reinterpret_cast<float4*>(o + 32) = make_float4(fmaxf(a1[0], 0.f), fmaxf(a1[1], 0.f), fmaxf(a1[2], 0.f), fmaxf(a1[3], 0.f));
            *reinterpret_cast<float4*>(o + 36) = make_float4(fmaxf(a1[4], 0.f), fmaxf(a1[5], 0.f), fmaxf(a1[6], 0.f), fmaxf(a1[7], 0.f));
        }
    }
}

// dw[tap][n] = sum_vox xn[vox + tap] * gz[row(vox)][n]; db[n] = sum gz.  A CTA stages 128 voxels at a time in shared
// memory (their 27 input taps + a constant 1 for the bias, and their 32 gradient channels) and forms the 28 x 32
// outer-product sums with warp w owning taps {w, w+8, w+16, w+24} and lane = output channel.  Per-CTA partials, reduced below.
__global__ void __launch_bounds__(256) first_conv_pr_wgrad_kernel(const float* __restrict__ xn, const float* __restrict__ gz,
                                                                  int B, int S, int T, RowGeom g, float* __restrict__ partials) {
    pdl_grid_wait();
    __shared__ float xs[128][29];
    __shared__ __align__(16) float gs[128][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long nvox = (long long)B * T * S * S;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long v0 = (long long)blockIdx.x * 128; v0 < nvox; v0 += (long long)gridDim.x * 128) {
        if (threadIdx.x < 128) {
            long long v = v0 + threadIdx.x;
            const bool ok = v < nvox;
            if (!ok) v = 0;
            const int ww = (int)(v % S); v /= S;
            const int hh = (int)(v % S); v /= S;
            const int tt = (int)(v % T); const long long b = v / T;
#pragma unroll
            for (int tap = 0; tap < 27; ++tap) {
                const int t2 = tt + tap / 9 - 1, h2 = hh + (tap / 3) % 3 - 1, w2 = ww + tap % 3 - 1;
                float x = 0.f;
                if (ok && t2 >= 0 && t2 < T && h2 >= 0 && h2 < S && w2 >= 0 && w2 < S) x = __ldg(xn + ((b * S + h2) * S + w2) * T + t2);
                xs[threadIdx.x][tap] = x;
            }
            xs[threadIdx.x][27] = ok ? 1.f : 0.f;
        } else {
            for (int f = threadIdx.x - 128; f < 128 * 8; f += 128) {
                const int i = f >> 3, c4 = f & 7;
                long long v = v0 + i;
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (v < nvox) {
                    const int ww = (int)(v % S); v /= S;
                    const int hh = (int)(v % S); v /= S;
                    const int tt = (int)(v % T); const long long b = v / T;
                    q = __ldg(reinterpret_cast<const float4*>(gz + (g.lead + b * g.pstride + (long long)(g.t0 + tt) * g.plane + hh * g.pw + ww) * 32) + c4);
                }
                *reinterpret_cast<float4*>(&gs[i][c4 * 4]) = q;
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < 128; ++i) {
            const float gv = gs[i][lane];
            acc[0] = fmaf(xs[i][warp], gv, acc[0]);
            acc[1] = fmaf(xs[i][warp + 8], gv, acc[1]);
            acc[2] = fmaf(xs[i][warp + 16], gv, acc[2]);
            if (warp < 4) acc[3] = fmaf(xs[i][warp + 24], gv, acc[3]);
        }
        __syncthreads();
    }
    float* out = partials + (size_t)blockIdx.x * (28 * 32);
    out[warp * 32 + lane] = acc[0];
    out[(warp + 8) * 32 + lane] = acc[1];
    out[(warp + 16) * 32 + lane] = acc[2];
    if (warp < 4) out[(warp + 24) * 32 + lane] = acc[3];
}

// Second formulation (the one launched for S = 22 when the partial budget allows it): a work item is (patch, chunk of temporal
// planes).  The item's normalised LR window is staged zero-padded in shared memory exactly as in the forward kernel; lane =
// output channel, a warp owns whole image lines (t, h): it first issues the line's S coalesced 128-byte gradient-row loads
// (S independent loads in flight per lane), then for each (dt, dh) reads the S + 2 inputs of that line once (broadcast LDS)
// and forms 3 S FMAs per lane -- 0.4 loads per FMA instead of 1.25, no scattered global gathers, no integer divisions in the
// inner loop.  Each warp keeps all 28 x 32 sums (27 taps + bias) in registers over every item of its CTA; the 8 warps are
// summed in a fixed order at the end and the CTA writes ONE partial [28][32] (layout of the kernel above, same reduction).
template <int S>
__global__ void __launch_bounds__(256) first_conv_pr_wgrad_items_kernel(const float* __restrict__ xn, const float* __restrict__ gz,
                                                                        int B, int T, int tchunks, RowGeom g,
                                                                        float* __restrict__ partials) {
    pdl_grid_wait();
    extern __shared__ float fsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tchunk = (T + tchunks - 1) / tchunks;
    constexpr int Sp = S + 2, wp = Sp + 1;
    float acc[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0.f;
    for (int item = blockIdx.x; item < B * tchunks; item += gridDim.x) {
        const int b = item / tchunks, tc0 = (item % tchunks) * tchunk, tcn = min(tchunk, T - tc0);
        if (tcn <= 0) continue;
        const int Tp = tcn + 2;
        __syncthreads();                                            // the previous item's window is no longer read
        for (int i = threadIdx.x; i < Tp * Sp * wp; i += 256) fsm[i] = 0.f;
        __syncthreads();
        for (int i = threadIdx.x; i < S * S * T; i += 256) {        // xn is [h][w][t]
            const int t = i % T, ww = (i / T) % S, hh = i / (T * S);
            const int tl = t - tc0 + 1;
            if (tl >= 0 && tl < Tp) fsm[(tl * Sp + hh + 1) * wp + ww + 1] = __ldg(xn + (size_t)b * S * S * T + i);
        }
        __syncthreads();
        const float* gzb = gz + (g.lead + (long long)b * g.pstride + (long long)(g.t0 + tc0) * g.plane) * 32 + lane;
#pragma unroll 1
        for (int line = warp; line < tcn * S; line += 8) {
            const int tt = line / S, hh = line % S;
            const float* gp = gzb + ((long long)tt * g.plane + hh * g.pw) * 32;
            float gv[S];
#pragma unroll
            for (int w = 0; w < S; ++w) gv[w] = __ldg(gp + w * 32);
            float bs = 0.f;
#pragma unroll
            for (int w = 0; w < S; ++w) bs += gv[w];
            acc[27] += bs;
#pragma unroll
            for (int dt = 0; dt < 3; ++dt)
#pragma unroll
                for (int dh = 0; dh < 3; ++dh) {
                    const float* xp = fsm + ((tt + dt) * Sp + hh + dh) * wp;           // padded line: tap dw of voxel w is xp[w + dw]
                    float xv[S + 2];
#pragma unroll
                    for (int w = 0; w < S + 2; ++w) xv[w] = xp[w];
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
                    for (int w = 0; w < S; ++w) {
                        a0 = fmaf(xv[w], gv[w], a0);
                        a1 = fmaf(xv[w + 1], gv[w], a1);
                        a2 = fmaf(xv[w + 2], gv[w], a2);
                    }
                    float* a3 = acc + (dt * 3 + dh) * 3;
                    a3[0] += a0; a3[1] += a1; a3[2] += a2;
                }
        }
    }
    __syncthreads();
    float* red = fsm;                                               // [8 warps][28][32]
#pragma unroll
    for (int k = 0; k < 28; ++k) red[(warp * 28 + k) * 32 + lane] = acc[k];
    __syncthreads();
    float* out = partials + (size_t)blockIdx.x * (28 * 32);
    for (int i = threadIdx.x; i < 28 * 32; i += 256) {
        float s0 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s0 += red[w * 28 * 32 + i];
        out[i] = s0;
    }
}

__global__ void __launch_bounds__(256)
first_conv_pr_wgrad_reduce_kernel(const float* __restrict__ partials, int ncta, float* __restrict__ dw, float* __restrict__ db) {
    __shared__ float4 sm[256];     // 7 blocks x 32 float4 columns = the 28 x 32 outputs (27 taps + bias); fixed order (wgrad_reduce.cuh)
    first_conv_reduce_body(blockIdx.x, partials, ncta, dw, db, sm);
}

// ------------------------------------------------------------------------------------------ PR <-> G (reflect pad)
__device__ __forceinline__ int refl(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// `pad` = reflect padding per side on H and W (tf.pad REFLECT, modelsTF.py:125-135,157): 1, or 0 for a plain re-layout
// (ConvReduceAndUpscalev2 pads nothing, modelsTF.py:166-175).  The source may be a PR or a G buffer.
// a_lo / g_pack (error-compensated engine, C4 == 8): the source is a (hi, lo) pair of row arrays; g0 receives the hi rows (what the
// backward pass reads) and g_pack the packed fp16 pair rows of (hi, lo) (what the compensated 3x3x3 convolution reads, rows.h) --
// one launch instead of two copies and a packing pass.
__global__ void pr_to_g_reflect_kernel(const float* __restrict__ a, RowGeom pr, float* __restrict__ g0, RowGeom gg,
                                       long long n, int C4, int pad, const float* __restrict__ a_lo, float* __restrict__ g_pack) {
    pdl_grid_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % C4); long long r = i / C4;
    const int w = (int)(r % gg.nw); r /= gg.nw;
    const int h = (int)(r % gg.nh); r /= gg.nh;
    const int t = (int)(r % gg.nt); const long long b = r / gg.nt;
    const long long src = pr.lead + b * pr.pstride + (long long)(pr.t0 + t) * pr.plane + refl(h - pad, pr.nh) * pr.pw + refl(w - pad, pr.nw);
    const long long dst = gg.lead + b * gg.pstride + (long long)(gg.t0 + t) * gg.plane + h * gg.pw + w;
    const float4 vh = __ldg(reinterpret_cast<const float4*>(a) + src * C4 + c);
    reinterpret_cast<float4*>(g0)[dst * C4 + c] = vh;
    if (g_pack) {
        const float4 l = __ldg(reinterpret_cast<const float4*>(a_lo) + src * C4 + c);
        const __half2 h0 = __floats2half2_rn(vh.x, vh.y), h1 = __floats2half2_rn(vh.z, vh.w);
        // lo' = lo + (hi - fp16(hi)): the packed pair always sums to hi + lo (rows.h)
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn((l.x + (vh.x - f0.x)) * PACK_SCALE, (l.y + (vh.y - f0.y)) * PACK_SCALE);
        const __half2 l1 = __floats2half2_rn((l.z + (vh.z - f1.x)) * PACK_SCALE, (l.w + (vh.w - f1.y)) * PACK_SCALE);
        uint2* pk = reinterpret_cast<uint2*>(g_pack + dst * 32);
        pk[c] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        pk[8 + c] = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    }
}

__device__ __forceinline__ int preimages(int i, int n, int p, int (&o)[3]) {
    int c = 0;
    o[c++] = i + p;
    if (i >= 1 && i <= p) o[c++] = p - i;
    if (i <= n - 2 && i >= n - 1 - p) o[c++] = p + 2 * (n - 1) - i;
    return c;
}

// channels 4q .. 4q+3 of 32-channel row `row` as a bf16 pair [bf16(v) x 32 | bf16(v - bf16(v)) x 32] (128 bytes per row)
__device__ __forceinline__ void store_bf16_pair4(float* pack, long long row, int q, float4 v) {
    const __nv_bfloat162 a0 = __floats2bfloat162_rn(v.x, v.y), a1 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f0 = __bfloat1622float2(a0), f1 = __bfloat1622float2(a1);
    const __nv_bfloat162 b0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), b1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
    uint2* dst = reinterpret_cast<uint2*>(pack + row * 32);
    dst[q] = make_uint2(*reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&a1));
    dst[8 + q] = make_uint2(*reinterpret_cast<const uint32_t*>(&b0), *reinterpret_cast<const uint32_t*>(&b1));
}

// adjoint of the kernel above; `relumask` (nullable, same rows as ga): the result is multiplied by (relumask > 0), i.e. the
// gradient flows into the ReLU output the padded tensor was made from (convReducePad_2/3 of the T = 13 graph)
__global__ void pr_to_g_reflect_bwd_kernel(const float* __restrict__ gg0, RowGeom gg, float* __restrict__ ga, RowGeom pr,
                                           long long n, int C4, int pad, const float* __restrict__ relumask, int round_tf32,
                                           float* __restrict__ ga_pack) {
    pdl_grid_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % C4); long long r = i / C4;
    const int w = (int)(r % pr.nw); r /= pr.nw;
    const int h = (int)(r % pr.nh); r /= pr.nh;
    const int t = (int)(r % pr.nt); const long long b = r / pr.nt;
    int hs[3], ws[3];
    const int nh = preimages(h, pr.nh, pad, hs), nw = preimages(w, pr.nw, pad, ws);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int x = 0; x < nh; ++x)
        for (int y = 0; y < nw; ++y) {
            const long long src = gg.lead + b * gg.pstride + (long long)(gg.t0 + t) * gg.plane + hs[x] * gg.pw + ws[y];
            const float4 q = __ldg(reinterpret_cast<const float4*>(gg0) + src * C4 + c);
            s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
        }
    const long long dst = pr.lead + b * pr.pstride + (long long)(pr.t0 + t) * pr.plane + h * pr.pw + w;
    if (relumask) {
        const float4 r4 = __ldg(reinterpret_cast<const float4*>(relumask) + dst * C4 + c);
        s.x = r4.x > 0.f ? s.x : 0.f; s.y = r4.y > 0.f ? s.y : 0.f; s.z = r4.z > 0.f ? s.z : 0.f; s.w = r4.w > 0.f ? s.w : 0.f;
    }
    // the result only feeds kind::tf32 MMAs, which TRUNCATE a raw fp32 operand (a -2.4e-4 relative bias on the whole upstream
    // gradient chain, profiles/r02_tf32_numerics_study.md): store it rounded to nearest instead
    if (ga_pack) store_bf16_pair4(ga_pack, dst, c, s);       // C4 == 8: the un-rounded result as a bf16 pair row (conv3_tc.cu MODE 2)
    if (round_tf32) { s.x = to_tf32(s.x); s.y = to_tf32(s.y); s.z = to_tf32(s.z); s.w = to_tf32(s.w); }
    reinterpret_cast<float4*>(ga)[dst * C4 + c] = s;
}

// sr[b, s*h+i, s*w+j] = (U[row(b,0,h,w)][i*s+j] + resid[b,h,w,i*s+j]) * std + mean
__global__ void tail_rows_kernel(const float* __restrict__ u, RowGeom g, int uc, const float* __restrict__ resid, long long n,
                                 int P, int s, float mean, float stdv, int clip_round, float* __restrict__ sr) {
    pdl_grid_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int PS = P * s;
    const int X = (int)(i % PS); long long r = i / PS;
    const int Y = (int)(r % PS); const long long b = r / PS;
    const int c = (Y % s) * s + (X % s);
    const long long urow = g.lead + b * g.pstride + (long long)g.t0 * g.plane + (Y / s) * g.pw + X / s;
    float v = (__ldg(u + urow * uc + c) + __ldg(resid + ((b * P + Y / s) * P + X / s) * (s * s) + c)) * stdv + mean;
    if (clip_round) v = rintf(fminf(fmaxf(v, 0.f), 65536.f));
    sr[i] = v;
}

__global__ void tail_bwd_rows_kernel(const float* __restrict__ dsr, long long n, int P, int s, float stdv,
                                     float* __restrict__ gu, RowGeom g, int uc, float* __restrict__ dtail, int round_tf32,
                                     float* __restrict__ gu_pack) {
    pdl_grid_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // index into [B,P,P,s*s]
    if (i >= n) return;
    const int c = (int)(i % (s * s)); long long r = i / (s * s);
    const int w = (int)(r % P); r /= P;
    const int h = (int)(r % P); const long long b = r / P;
    const int PS = P * s;
    const float v = __ldg(dsr + (b * PS + h * s + c / s) * PS + w * s + c % s) * stdv;
    dtail[i] = v;
    const long long row = g.lead + b * g.pstride + (long long)g.t0 * g.plane + h * g.pw + w;
    gu[row * uc + c] = round_tf32 ? to_tf32(v) : v;   // MMA operand: see pr_to_g_reflect_bwd_kernel
    if (gu_pack) {                                    // uc == 32: bf16 pair row, channels >= s * s stay zero
        __nv_bfloat16* pk = reinterpret_cast<__nv_bfloat16*>(gu_pack + row * 32);
        const __nv_bfloat16 a = __float2bfloat16_rn(v);
        pk[c] = a;
        pk[32 + c] = __float2bfloat16_rn(v - __bfloat162float(a));
    }
}

}  // namespace

int launch_tail_rows(const float* u, RowGeom g, int uc, const float* resid, int B, int P, int scale, float mean, float stdv,
                     int clip_round, float* sr, cudaStream_t st) {
    const long long n = (long long)B * P * scale * P * scale;
    PV_TIMED("tail", st, 0.0, (double)n * 12.0);
    PV_CUDA(launch_pdl_simple(tail_rows_kernel, cdiv(n, 256), 256, 0, st, u, g, uc, resid, n, P, scale, mean, stdv, clip_round, sr));
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_tail_bwd_rows(const float* dsr, int B, int P, int scale, float stdv, float* gu, RowGeom g, int uc, float* dtail, cudaStream_t st, int round_tf32,
                         float* gu_pack) {
    if (gu_pack && uc != 32) return set_error(PV_ERR_BAD_ARG, "tail_bwd: bf16 pair rows need 32-channel rows");
    const long long n = (long long)B * P * P * scale * scale;
    PV_TIMED("tail_bwd", st, 0.0, (double)n * 12.0);
    PV_CUDA(launch_pdl_simple(tail_bwd_rows_kernel, cdiv(n, 256), 256, 0, st, dsr, n, P, scale, stdv, gu, g, uc, dtail, round_tf32, gu_pack));
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_rowconv_simt(const RowConvP& p, cudaStream_t st) {
    const long long M = (long long)p.B * p.og.nrows;
    if (M <= 0 || p.kc % 16 || p.n % 32 || p.xc % 4) return set_error(PV_ERR_BAD_ARG, "rowconv_simt: bad shape");
    PV_TIMED(p.tag ? p.tag : "rowconv_simt", st, p.flops, 0.0);
    if (p.n % 64 == 0) { dim3 grid(cdiv(M, 128), p.n / 64); rowconv_vec_kernel<64><<<grid, 256, 0, st>>>(p); }
    else { dim3 grid(cdiv(M, 128), p.n / 32); rowconv_vec_kernel<32><<<grid, 256, 0, st>>>(p); }
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_rowwgrad_simt(const RowWgradP& p, cudaStream_t st) {
    const long long M = (long long)p.B * p.og.nrows;
    const int Ktot = p.ntap * p.kc;
    if (M <= 0 || p.kc % 16 || p.n % 32 || p.xc % 4) return set_error(PV_ERR_BAD_ARG, "rowwgrad_simt: bad shape");
    PV_TIMED(p.tag ? p.tag : "rowwgrad_simt", st, p.flops, 0.0);
    const int BN = (p.n % 64 == 0) ? 64 : 32;
    const int kt = cdiv(Ktot, 64), nt = p.n / BN;
    int msplit = 148 * 4 / (kt * nt); if (msplit < 1) msplit = 1;
    long long per = (M + msplit - 1) / msplit;
    per = ((per + 31) / 32) * 32;
    msplit = cdiv(M, per);
    dim3 grid(kt, nt, msplit);
    if (BN == 64) rowwgrad_vec_kernel<64><<<grid, 256, 0, st>>>(p, (int)per);
    else rowwgrad_vec_kernel<32><<<grid, 256, 0, st>>>(p, (int)per);
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_first_conv_pr(const float* xn, const float* w, const float* bias, int B, int S, int T, float* y, RowGeom g, cudaStream_t st, float* y_lo, int round_tf32) {
    const size_t smem = sizeof(float) * ((size_t)(T + 2) * (S + 2) * (S + 3) + 27 * 32 + 32);
    if (smem > 200 * 1024) return set_error(PV_ERR_BAD_ARG, "first_conv_pr: patch of %dx%dx%d does not fit in shared memory", S, S, T);
    static size_t attr[16] = {};
    PV_CUDA(ensure_dyn_smem(first_conv_pr_kernel, smem, attr));
    PV_TIMED("first_conv_pr", st, 2.0 * B * T * S * S * 27 * 32, 0.0);
    PV_CUDA(launch_pdl_simple(first_conv_pr_kernel, dim3(B, 1), 256, smem, st, xn, w, bias, B, S, T, y, g, y_lo, round_tf32));
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_first_conv_pr_wgrad(const float* xn, const float* gz, int B, int S, int T, RowGeom g, float* dw, float* db,
                               float* partials, size_t partial_floats, cudaStream_t st, ReduceQueue* rq) {
    const int max_grid = 148 * 4;                                   // partial budget of the trainer's arena (engine.cu)
    // items = (patch, chunk of planes): three chunks per patch while that fits the partial budget (B <= 197), else fewer
    int tchunks = 3;
    while (tchunks > 1 && (long long)B * tchunks > max_grid) --tchunks;
    if (tchunks > T) tchunks = T;
    const int tchunk = (T + tchunks - 1) / tchunks;
    const size_t smem_win = sizeof(float) * (size_t)(tchunk + 2) * (S + 2) * (S + 3), smem_red = sizeof(float) * 8 * 28 * 32;
    const size_t smem = smem_win > smem_red ? smem_win : smem_red;
    const bool items = S == 22 && getenv("PV_FIRST_WGRAD_V1") == nullptr && smem <= 200 * 1024;
    const long long nitems = (long long)B * tchunks;
    const int grid = items ? (int)(nitems < max_grid ? nitems : max_grid) : max_grid;
    float* deferred = rq ? rq->take((size_t)grid * 28 * 32) : nullptr;
    if (deferred) { partials = deferred; partial_floats = (size_t)grid * 28 * 32; }
    if (!partials || partial_floats < (size_t)grid * 28 * 32) return set_error(PV_ERR_BAD_ARG, "first_conv_pr_wgrad: partial buffer too small");
    PV_TIMED("first_conv_pr_wgrad", st, 2.0 * B * T * S * S * 27 * 32, 0.0);
    if (items) {
        static size_t attr[16] = {};
        PV_CUDA(ensure_dyn_smem(first_conv_pr_wgrad_items_kernel<22>, smem, attr));
        PV_CUDA(launch_pdl_simple(first_conv_pr_wgrad_items_kernel<22>, grid, 256, smem, st, xn, gz, B, T, tchunks, g, partials));
    } else {
        PV_CUDA(launch_pdl_simple(first_conv_pr_wgrad_kernel, grid, 256, 0, st, xn, gz, B, S, T, g, partials));
    }
    PV_LAUNCH_CHECK();
    if (deferred) {
        ReduceJob j;
        memset(&j, 0, sizeof j);
        j.kind = 2; j.nblocks = FIRST_CONV_REDUCE_BLOCKS; j.partials = partials; j.ncta = grid; j.sc.dw = dw; j.sc.db = db;
        rq->push(j);
    } else {
        first_conv_pr_wgrad_reduce_kernel<<<FIRST_CONV_REDUCE_BLOCKS, 256, 0, st>>>(partials, grid, dw, db);
        PV_LAUNCH_CHECK();
    }
    return 0;
}

int launch_pr_to_g_reflect(const float* a, RowGeom pr, float* g0, RowGeom gg, int B, int C, cudaStream_t st, int pad, const float* a_lo, float* g_pack) {
    if ((a_lo != nullptr) != (g_pack != nullptr) || (g_pack && C != 32)) return set_error(PV_ERR_BAD_ARG, "pr_to_g_reflect: pair rows need the lo source and 32-channel rows");
    if (pad < 0 || pad > 1 || gg.nh != pr.nh + 2 * pad || gg.nw != pr.nw + 2 * pad || gg.nt != pr.nt)
        return set_error(PV_ERR_BAD_ARG, "pr_to_g_reflect: %dx%dx%d + pad %d does not give %dx%dx%d", pr.nh, pr.nw, pr.nt, pad, gg.nh, gg.nw, gg.nt);
    const long long n = (long long)B * gg.nt * gg.nh * gg.nw * (C / 4);
    PV_TIMED("pr_to_g_reflect", st);
    PV_CUDA(launch_pdl_simple(pr_to_g_reflect_kernel, cdiv(n, 256), 256, 0, st, a, pr, g0, gg, n, C / 4, pad, a_lo, g_pack));
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_pr_to_g_reflect_bwd(const float* gg0, RowGeom gg, float* ga, RowGeom pr, int B, int C, cudaStream_t st, int pad,
                               const float* relumask, int round_tf32, float* ga_pack) {
    if (ga_pack && C != 32) return set_error(PV_ERR_BAD_ARG, "pr_to_g_reflect_bwd: bf16 pair rows need 32-channel rows");
    if (pad < 0 || pad > 1 || gg.nh != pr.nh + 2 * pad || gg.nw != pr.nw + 2 * pad || gg.nt != pr.nt)
        return set_error(PV_ERR_BAD_ARG, "pr_to_g_reflect_bwd: geometry mismatch");
    const long long n = (long long)B * pr.nt * pr.nh * pr.nw * (C / 4);
    PV_TIMED("pr_to_g_reflect_bwd", st);
    PV_CUDA(launch_pdl_simple(pr_to_g_reflect_bwd_kernel, cdiv(n, 256), 256, 0, st, gg0, gg, ga, pr, n, C / 4, pad, relumask, round_tf32, ga_pack));
    PV_LAUNCH_CHECK();
    return 0;
}

}  // namespace pv
