// conv_simt.cu -- fp32 CUDA-core implicit-GEMM convolution: forward / data-gradient / weight-gradient.
//
// This is the exact-fp32 path (pv_cfg.precision = 0) for every weight-normalised Conv3D/Conv2D of the
// WDSR graph (reference models/modelsTF.py:58,179-186,159-163,47-50 -> Keras Conv3D/Conv2D, cuDNN in TF),
// and the permanent path for the layers that are not tensor-core shaped (Cin=1 mainConv1, the 9-channel
// 2-D skip convs, the 32->9 upscale conv).  In TF the backward of these is Conv3DBackpropInputV2 /
// Conv3DBackpropFilterV2 / ReluGrad / BiasAddGrad (tape.gradient, trainClass.py:131).
//
//   conv_vec_kernel  : 128 voxels x BN couts per CTA, K chunks of 16 channels inside one tap, float4 gathers
//                      (needs cin % 16 == 0, cout % 32 == 0).  Zero padding = predicated gather.
//   conv_direct_kernel: one thread per (voxel, cout) -- the tiny / odd-shaped layers.
//   wgrad_vec_kernel : 64 k-rows x BN couts per CTA, reduction over a slice of the voxels, fp32 atomics out.
//   wgrad_direct_kernel: one thread per (k, cout), CTA-uniform walk over a voxel slice.
#include "kernels.h"

namespace pv {
namespace {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------------------------------ conv_vec
template <int BN>
__global__ void __launch_bounds__(256) conv_vec_kernel(ConvP p) {
    constexpr int BM = 128, BK = 16, TXN = BN / 4, TYN = 256 / TXN, TM = BM / TYN;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
    const long long M = (long long)p.B * p.Ho * p.Wo * p.To;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // gather bookkeeping for the two rows this thread loads
    int hi0[2], wi0[2], ti0[2];
    long long vbase[2];
    bool rvalid[2];
    const int kq = tid & 3;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int row = (tid >> 2) + 64 * e;
        long long m = m0 + row;
        rvalid[e] = m < M;
        if (!rvalid[e]) m = 0;
        const int to = (int)(m % p.To); m /= p.To;
        const int wo = (int)(m % p.Wo); m /= p.Wo;
        const int ho = (int)(m % p.Ho); m /= p.Ho;
        hi0[e] = ho - p.ph; wi0[e] = wo - p.pw; ti0[e] = to - p.pt;
        vbase[e] = m * p.Hi * p.Wi * p.Ti;
    }
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int taps = p.kh * p.kw * p.kt;
    for (int tap = 0; tap < taps; ++tap) {
        const int dt = tap % p.kt, dw = (tap / p.kt) % p.kw, dh = tap / (p.kt * p.kw);
        long long off[2];
        bool inb[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int hi = hi0[e] + dh, wi = wi0[e] + dw, ti = ti0[e] + dt;
            inb[e] = rvalid[e] && hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi && ti >= 0 && ti < p.Ti;
            off[e] = (vbase[e] + ((long long)hi * p.Wi + wi) * p.Ti + ti) * p.cin + kq * 4;
        }
        for (int c0 = 0; c0 < p.cin; c0 += BK) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (inb[e]) {
                    v = ldg4(p.x + off[e] + c0);
                    if (p.xmask) {
                        const float4 r = ldg4(p.xmask + off[e] + c0);
                        v.x = r.x > 0.f ? v.x : 0.f; v.y = r.y > 0.f ? v.y : 0.f;
                        v.z = r.z > 0.f ? v.z : 0.f; v.w = r.w > 0.f ? v.w : 0.f;
                    }
                }
                const int row = (tid >> 2) + 64 * e;
                As[kq * 4 + 0][row] = v.x; As[kq * 4 + 1][row] = v.y;
                As[kq * 4 + 2][row] = v.z; As[kq * 4 + 3][row] = v.w;
            }
            if (tid < BK * BN / 4) {
                const int kk = tid / (BN / 4), nn = (tid % (BN / 4)) * 4;
                const float4 wv = ldg4(p.w + ((long long)tap * p.cin + c0 + kk) * p.cout + n0 + nn);
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
    // epilogue: bias, residual, ReLU, float4 store (y is [M][cout])
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bv = ldg4(p.bias + n0 + tx * 4);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long m = m0 + ty * TM + i;
        if (m >= M) continue;
        float4 o = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
        const long long yo = m * p.cout + n0 + tx * 4;
        if (p.residual) {
            const float4 r = ldg4(p.residual + yo);
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(p.y + yo) = o;
    }
}

// ------------------------------------------------------------------------------------------ conv_direct
__global__ void __launch_bounds__(256) conv_direct_kernel(ConvP p) {
    const long long M = (long long)p.B * p.Ho * p.Wo * p.To;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= M * p.cout) return;
    const int co = (int)(idx % p.cout);
    long long m = idx / p.cout;
    const int to = (int)(m % p.To); m /= p.To;
    const int wo = (int)(m % p.Wo); m /= p.Wo;
    const int ho = (int)(m % p.Ho); m /= p.Ho;
    const long long vbase = m * p.Hi * p.Wi * p.Ti;
    float acc = p.bias ? __ldg(p.bias + co) : 0.f;
    for (int a = 0; a < p.kh; ++a) {
        const int hi = ho - p.ph + a;
        if (hi < 0 || hi >= p.Hi) continue;
        for (int b = 0; b < p.kw; ++b) {
            const int wi = wo - p.pw + b;
            if (wi < 0 || wi >= p.Wi) continue;
            for (int c = 0; c < p.kt; ++c) {
                const int ti = to - p.pt + c;
                if (ti < 0 || ti >= p.Ti) continue;
                const long long xo = (vbase + ((long long)hi * p.Wi + wi) * p.Ti + ti) * p.cin;
                const float* wrow = p.w + (long long)(((a * p.kw + b) * p.kt + c)) * p.cin * p.cout + co;
                for (int ci = 0; ci < p.cin; ++ci) {
                    float xv = __ldg(p.x + xo + ci);
                    if (p.xmask && !(__ldg(p.xmask + xo + ci) > 0.f)) xv = 0.f;
                    acc = fmaf(xv, __ldg(wrow + (long long)ci * p.cout), acc);
                }
            }
        }
    }
    if (p.residual) acc += __ldg(p.residual + idx);
    if (p.relu) acc = fmaxf(acc, 0.f);
    p.y[idx] = acc;
}

// ------------------------------------------------------------------------------------------ wgrad_vec
template <int BN>
__global__ void __launch_bounds__(256) wgrad_vec_kernel(WgradP p, int m_per_cta, float* __restrict__ part) {
    constexpr int BKO = 64, BMC = 32, TXN = BN / 4, TYN = 256 / TXN, TK = BKO / TYN;
    __shared__ __align__(16) float At[BMC][BKO + 4];
    __shared__ __align__(16) float Ys[BMC][BN];
    const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
    const long long M = (long long)p.B * p.Ho * p.Wo * p.To;
    const int k0 = blockIdx.x * BKO, n0 = blockIdx.y * BN;
    const int Ktot = p.kh * p.kw * p.kt * p.cin;
    long long mlo = (long long)blockIdx.z * m_per_cta;
    long long mhi = mlo + m_per_cta; if (mhi > M) mhi = M;

    // the (tap, ci) of the float4 column this thread gathers never changes
    const int kq = tid & 15;
    const int kg = k0 + kq * 4;
    const bool kvalid = kg < Ktot;
    const int tap = kvalid ? kg / p.cin : 0, ci = kvalid ? kg % p.cin : 0;
    const int dt = tap % p.kt, dw = (tap / p.kt) % p.kw, dh = tap / (p.kt * p.kw);

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
            long long m = mc + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < mhi && kvalid) {
                const int to = (int)(m % p.To); m /= p.To;
                const int wo = (int)(m % p.Wo); m /= p.Wo;
                const int ho = (int)(m % p.Ho); m /= p.Ho;
                const int hi = ho - p.ph + dh, wi = wo - p.pw + dw, ti = to - p.pt + dt;
                if (hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi && ti >= 0 && ti < p.Ti)
                    v = ldg4(p.x + (((m * p.Hi + hi) * p.Wi + wi) * p.Ti + ti) * p.cin + ci);
            }
            *reinterpret_cast<float4*>(&At[row][kq * 4]) = v;
        }
        for (int f = tid; f < BMC * BN / 4; f += 256) {
            const int row = f / (BN / 4), nn = (f % (BN / 4)) * 4;
            const long long m = mc + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < mhi) {
                const long long yo = m * p.cout + n0 + nn;
                v = ldg4(p.dy + yo);
                if (p.ymask) {
                    const float4 r = ldg4(p.ymask + yo);
                    v.x = r.x > 0.f ? v.x : 0.f; v.y = r.y > 0.f ? v.y : 0.f;
                    v.z = r.z > 0.f ? v.z : 0.f; v.w = r.w > 0.f ? v.w : 0.f;
                }
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
    if (part) {     // deterministic: this split's partial sums, one writer per element; summed in split order by wgrad_sum_kernel
        float* mine = part + (size_t)blockIdx.z * (size_t)(Ktot + 1) * p.cout;
#pragma unroll
        for (int i = 0; i < TK; ++i) {
            const int k = k0 + ty * TK + i;
            if (k >= Ktot) continue;
            *reinterpret_cast<float4*>(mine + (size_t)k * p.cout + n0 + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        if (do_bias) *reinterpret_cast<float4*>(mine + (size_t)Ktot * p.cout + n0 + tx * 4) = make_float4(bsum[0], bsum[1], bsum[2], bsum[3]);
        return;
    }
#pragma unroll
    for (int i = 0; i < TK; ++i) {
        const int k = k0 + ty * TK + i;
        if (k >= Ktot) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(p.dw + (long long)k * p.cout + n0 + tx * 4 + j, acc[i][j]);
    }
    if (do_bias)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(p.db + n0 + tx * 4 + j, bsum[j]);
}

// dw[k][n] (+ db[n] from row Ktot) = sum over the splits, in split order
__global__ void __launch_bounds__(256) wgrad_sum_kernel(const float* __restrict__ part, int nsplit, int Ktot, int cout, float* __restrict__ dw,
                                                        float* __restrict__ db) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int n_all = (Ktot + 1) * cout;
    if (idx >= n_all) return;
    float s = 0.f;
    for (int z = 0; z < nsplit; ++z) s += part[(size_t)z * n_all + idx];
    if (idx < Ktot * cout) dw[idx] += s;
    else if (db) db[idx - Ktot * cout] += s;
}

// ------------------------------------------------------------------------------------------ wgrad_direct
__global__ void __launch_bounds__(256) wgrad_direct_kernel(WgradP p, int m_per_cta, float* __restrict__ part) {
    const long long M = (long long)p.B * p.Ho * p.Wo * p.To;
    const int Ktot = p.kh * p.kw * p.kt * p.cin;
    const int idx = blockIdx.x * 256 + threadIdx.x;          // (k, n) pair, n fastest; k == Ktot rows -> bias
    const bool isw = idx < Ktot * p.cout;
    const bool isb = !isw && idx < (Ktot + 1) * p.cout && p.db != nullptr;
    const int n = idx % p.cout;
    const int k = isw ? idx / p.cout : 0;
    const int ci = k % p.cin, tap = k / p.cin;
    const int dt = tap % p.kt, dw = (tap / p.kt) % p.kw, dh = tap / (p.kt * p.kw);
    long long mlo = (long long)blockIdx.y * m_per_cta;
    long long mhi = mlo + m_per_cta; if (mhi > M) mhi = M;
    if (mlo >= mhi) {
        if (part && (isw || isb)) part[(size_t)blockIdx.y * (size_t)(Ktot + 1) * p.cout + idx] = 0.f;
        return;
    }
    // CTA-uniform incremental decode of the voxel index
    long long t = mlo;
    int to = (int)(t % p.To); t /= p.To;
    int wo = (int)(t % p.Wo); t /= p.Wo;
    int ho = (int)(t % p.Ho); t /= p.Ho;
    long long b = t;
    float acc = 0.f;
    for (long long m = mlo; m < mhi; ++m) {
        if (isw || isb) {
            float g = __ldg(p.dy + m * p.cout + n);
            if (p.ymask && !(__ldg(p.ymask + m * p.cout + n) > 0.f)) g = 0.f;
            if (isb) acc += g;
            else {
                const int hi = ho - p.ph + dh, wi = wo - p.pw + dw, ti = to - p.pt + dt;
                if (hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi && ti >= 0 && ti < p.Ti)
                    acc = fmaf(__ldg(p.x + (((b * p.Hi + hi) * p.Wi + wi) * p.Ti + ti) * p.cin + ci), g, acc);
            }
        }
        if (++to == p.To) { to = 0; if (++wo == p.Wo) { wo = 0; if (++ho == p.Ho) { ho = 0; ++b; } } }
    }
    if (part) { if (isw || isb) part[(size_t)blockIdx.y * (size_t)(Ktot + 1) * p.cout + idx] = acc; return; }
    if (isw) atomicAdd(p.dw + idx, acc);
    else if (isb) atomicAdd(p.db + n, acc);
}

}  // namespace

int launch_conv(const ConvP& p, cudaStream_t st) {
    const long long M = (long long)p.B * p.Ho * p.Wo * p.To;
    if (M <= 0) return set_error(PV_ERR_BAD_ARG, "conv: empty output");
    const bool vec = (p.cin % 16 == 0) && (p.cout % 32 == 0);
    PV_TIMED(p.tag ? p.tag : (vec ? "conv_vec" : "conv_direct"), st,
             2.0 * (double)M * p.kh * p.kw * p.kt * p.cin_r * p.cout_r,
             4.0 * ((double)p.B * p.Hi * p.Wi * p.Ti * p.cin_r + (double)M * p.cout_r));
    if (vec) {
        if (p.cout % 64 == 0) {
            dim3 grid(cdiv(M, 128), p.cout / 64);
            conv_vec_kernel<64><<<grid, 256, 0, st>>>(p);
        } else {
            dim3 grid(cdiv(M, 128), p.cout / 32);
            conv_vec_kernel<32><<<grid, 256, 0, st>>>(p);
        }
    } else {
        conv_direct_kernel<<<cdiv(M * p.cout, 256), 256, 0, st>>>(p);
    }
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_wgrad(const WgradP& p, cudaStream_t st) {
    const long long M = (long long)p.B * p.Ho * p.Wo * p.To;
    const int Ktot = p.kh * p.kw * p.kt * p.cin;
    const bool vec = (p.cin % 16 == 0) && (p.cout % 32 == 0);
    PV_TIMED(p.tag ? p.tag : (vec ? "wgrad_vec" : "wgrad_direct"), st,
             2.0 * (double)M * Ktot / p.cin * p.cin_r * p.cout_r,
             4.0 * ((double)p.B * p.Hi * p.Wi * p.Ti * p.cin_r + (double)M * p.cout_r));
    if (vec) {
        const int BN = (p.cout % 64 == 0) ? 64 : 32;
        const int kt = cdiv(Ktot, 64), nt = p.cout / BN;
        int msplit = 148 * 4 / (kt * nt); if (msplit < 1) msplit = 1;
        long long per = (M + msplit - 1) / msplit;
        per = ((per + 31) / 32) * 32;
        msplit = cdiv(M, per);
        dim3 grid(kt, nt, msplit);
        const size_t need = (size_t)msplit * (Ktot + 1) * p.cout;
        float* part = (p.partials && p.partial_floats >= need) ? p.partials : nullptr;
        if (part && !p.db) PV_CUDA(cudaMemsetAsync(part, 0, need * sizeof(float), st));       // bias rows are only written when db is set
        if (BN == 64) wgrad_vec_kernel<64><<<grid, 256, 0, st>>>(p, (int)per, part);
        else wgrad_vec_kernel<32><<<grid, 256, 0, st>>>(p, (int)per, part);
        PV_LAUNCH_CHECK();
        if (part) wgrad_sum_kernel<<<cdiv((long long)(Ktot + 1) * p.cout, 256), 256, 0, st>>>(part, msplit, Ktot, p.cout, p.dw, p.db);
    } else {
        const int nb = cdiv((long long)(Ktot + 1) * p.cout, 256);
        int msplit = 148 * 8 / nb; if (msplit < 1) msplit = 1;
        if (msplit > M) msplit = (int)M;
        const long long per = (M + msplit - 1) / msplit;
        const int nsplit = cdiv(M, per);
        dim3 grid(nb, nsplit);
        const size_t need = (size_t)nsplit * (Ktot + 1) * p.cout;
        float* part = (p.partials && p.partial_floats >= need) ? p.partials : nullptr;
        if (part && !p.db) PV_CUDA(cudaMemsetAsync(part, 0, need * sizeof(float), st));
        wgrad_direct_kernel<<<grid, 256, 0, st>>>(p, (int)per, part);
        PV_LAUNCH_CHECK();
        if (part) wgrad_sum_kernel<<<cdiv((long long)(Ktot + 1) * p.cout, 256), 256, 0, st>>>(part, nsplit, Ktot, p.cout, p.dw, p.db);
    }
    PV_LAUNCH_CHECK();
    return 0;
}

}  // namespace pv
