// glue.cu -- the bandwidth-bound pieces around the convolutions:
//   prep            : (x-mean)/std + normalised temporal mean            (modelsTF.py:23-27,199-200)
//   reflect pad     : tf.pad(mode='reflect') and its adjoint              (modelsTF.py:157-158, 78-138)
//   tail            : depth_to_space x2 + add + denormalise [+ clip/round] (modelsTF.py:38-41,52,73; test.py:118-119)
//   weight norm     : TFA WeightNormalization kernel = g * v/||v||, all layers in one launch, + backward
//   optimizers      : Keras Nadam / Adam / SGD over the flat parameter arena (train.py:77-83)
//   scene geometry  : reflect-pad-3 + unfold(22, stride 16) and the n x n stitch (dataGenerator.py:108-121; test.py:149-160)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.h"
#include "rows.h"

namespace pv {
namespace {

// ------------------------------------------------------------------------------------------ prep
__global__ void prep_kernel(const float* __restrict__ lr, long long nvox, int T, float mean, float inv_std,
                            float* __restrict__ xn, float* __restrict__ mn) {
    pdl_grid_wait();
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, h, w)
    if (v >= nvox) return;
    const float* src = lr + v * T;
    float s = 0.f;
    for (int t = 0; t < T; ++t) {
        const float x = __ldg(src + t);
        s += x;
        xn[v * T + t] = (x - mean) * inv_std;
    }
    mn[v] = (s / (float)T - mean) * inv_std;
}

// ------------------------------------------------------------------------------------------ reflect pad
__device__ __forceinline__ int refl(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void reflect_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H, int W, int T,
                                   int C4, int ph, int pw, int pt) {
    const int Ho = H + 2 * ph, Wo = W + 2 * pw, To = T + 2 * pt;
    const long long n = (long long)B * Ho * Wo * To * C4;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % C4); long long r = i / C4;
    const int t = (int)(r % To); r /= To;
    const int w = (int)(r % Wo); r /= Wo;
    const int h = (int)(r % Ho); r /= Ho;
    const long long src = (((r * H + refl(h - ph, H)) * W + refl(w - pw, W)) * T + refl(t - pt, T)) * C4 + c;
    reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(in) + src);
}

// pre-images of position i of the un-padded axis inside the padded axis: i+p, and its mirror images
__device__ __forceinline__ int preimages(int i, int n, int p, int (&o)[3]) {
    int c = 0;
    o[c++] = i + p;
    if (i >= 1 && i <= p) o[c++] = p - i;
    if (i <= n - 2 && i >= n - 1 - p) o[c++] = p + 2 * (n - 1) - i;
    return c;
}

__global__ void reflect_pad_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int B, int H, int W,
                                       int T, int C4, int ph, int pw, int pt) {
    const int Ho = H + 2 * ph, Wo = W + 2 * pw, To = T + 2 * pt;
    const long long n = (long long)B * H * W * T * C4;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % C4); long long r = i / C4;
    const int t = (int)(r % T); r /= T;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    int hs[3], ws[3], ts[3];
    const int nh = preimages(h, H, ph, hs), nw = preimages(w, W, pw, ws), nt = preimages(t, T, pt, ts);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int x = 0; x < nh; ++x)
        for (int y = 0; y < nw; ++y)
            for (int z = 0; z < nt; ++z) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gout) +
                                       (((r * Ho + hs[x]) * Wo + ws[y]) * To + ts[z]) * C4 + c);
                a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
            }
    reinterpret_cast<float4*>(gin)[i] = a;
}

// ------------------------------------------------------------------------------------------ tail
// sr[b, s*h+i, s*w+j] = (up[b,h,w,i*s+j] + resid[b,h,w,i*s+j]) * std + mean     (depth_to_space, Cout = 1)
__global__ void tail_kernel(const float* __restrict__ up, const float* __restrict__ resid, long long n, int P, int s,
                            float mean, float stdv, int clip_round, float* __restrict__ sr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int PS = P * s;
    const int X = (int)(i % PS); long long r = i / PS;
    const int Y = (int)(r % PS); const long long b = r / PS;
    const long long src = ((b * P + Y / s) * P + X / s) * (s * s) + (Y % s) * s + (X % s);
    float v = (__ldg(up + src) + __ldg(resid + src)) * stdv + mean;
    if (clip_round) v = rintf(fminf(fmaxf(v, 0.f), 65536.f));      // tf.clip_by_value(0, 2**16); tf.round = half-to-even
    sr[i] = v;
}

// d(up) = d(resid) = space_to_depth(dsr) * std
__global__ void tail_bwd_kernel(const float* __restrict__ dsr, long long n, int P, int s, float stdv, float* __restrict__ dtail) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // index into [B,P,P,s*s]
    if (i >= n) return;
    const int c = (int)(i % (s * s)); long long r = i / (s * s);
    const int w = (int)(r % P); r /= P;
    const int h = (int)(r % P); const long long b = r / P;
    const int PS = P * s;
    dtail[i] = __ldg(dsr + (b * PS + h * s + c / s) * PS + w * s + c % s) * stdv;
}

// ------------------------------------------------------------------------------------------ weight norm
__device__ __forceinline__ const WnLayer& find_layer(const WnLayer* tab, int nlayers, int block, int& co) {
    int lo = 0, hi = nlayers - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab[mid].first_block <= block) lo = mid; else hi = mid - 1;
    }
    co = block - tab[lo].first_block;
    return tab[lo];
}

// row engine: TF stores taps as (dh, dw, dt) (Keras kernel [kh,kw,kt,...] on (H,W,T)); the row layouts order them (dt, dh, dw)
__device__ __forceinline__ int row_tap(const WnLayer& L, int tap) {
    if (L.mode == 0 || L.taps != 27) return tap;
    const int dh = tap / 9, dw = (tap / 3) % 3, dt = tap % 3;
    return (dt * 3 + dh) * 3 + dw;
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__device__ __forceinline__ float block_sum_128(float v, float* red) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    return red[0] + red[1] + red[2] + red[3];
}

// one CTA (128 threads) per (layer, cout):  w = v * g * rsqrt(max(sum v^2, 1e-12))   (tf.nn.l2_normalize * g)
// writes the forward layout weff[(tap*cin_s+ci)*cout_s+co] and the data-gradient layout
// weffT[((taps-1-tap)*cout_s+co)*cin_s+ci] (flipped taps, transposed channels), plus the padded bias.
__global__ void __launch_bounds__(128) wn_prep_kernel(const WnLayer* __restrict__ tab, int nlayers,
                                                      const float* __restrict__ params, float* __restrict__ weff,
                                                      float* __restrict__ weffT, float* __restrict__ bias_s,
                                                      float* __restrict__ scale, float* __restrict__ weff_lo,
                                                      float* __restrict__ weffT_lo, float* __restrict__ weffT_pack,
                                                      float* __restrict__ weff_pack) {
    pdl_grid_wait();
    __shared__ float red[4];
    int co;
    const WnLayer L = find_layer(tab, nlayers, blockIdx.x, co);
    const int K = L.taps * L.cin;
    const float* v = params + L.v_off;
    float ss = 0.f;
    for (int k = threadIdx.x; k < K; k += 128) { const float x = v[(long long)k * L.cout + co]; ss = fmaf(x, x, ss); }
    ss = block_sum_128(ss, red);
    const float rn = rsqrtf(fmaxf(ss, 1e-12f));
    const float sc = params[L.g_off + co] * rn;
    if (threadIdx.x == 0) {
        scale[L.scale_off + co] = sc;
        scale[L.scale_off + L.cout + co] = rn;
        bias_s[L.bias_s_off + co] = params[L.b_off + co];
    }
    for (int k = threadIdx.x; k < K; k += 128) {
        const int tap = k / L.cin, ci = k % L.cin;
        const float wf = v[(long long)k * L.cout + co] * sc;
        float w = wf;
        if (L.round_tf32) w = to_tf32(w);
        const int rt = row_tap(L, tap);
        weff[L.weff_off + ((long long)rt * L.cin_s + ci) * L.cout_s + co] = w;
        if (L.mode == 0) weffT[L.weffT_off + ((long long)(L.taps - 1 - tap) * L.cout_s + co) * L.cin_s + ci] = w;
        else weffT[L.weffT_off + (long long)co * (L.taps * L.cin_s) + (long long)rt * L.cin_s + ci] = w;
        if (weff_lo && L.mode == 1) {       // error-compensated engine: the remainder of the tf32 rounding, same two layouts
            weff_lo[L.weff_off + ((long long)rt * L.cin_s + ci) * L.cout_s + co] = wf - w;
            weffT_lo[L.weffT_off + (long long)co * (L.taps * L.cin_s) + (long long)rt * L.cin_s + ci] = wf - w;
        }
        {
            if (weffT_pack && L.mode == 1 && L.cin_s % 32 == 0) {     // (also the single-pass engine: its inference convs read the hi halves)
                // packed fp16 pair row of (co, tap, 32-channel chunk of ci): [ fp16(PACK_SCALE * w_lo) x 32 | fp16(w_hi) x 32 ] (rows.h) in the
                // 128 bytes the 32 fp32 K-values of weffT occupy (3x3x3 layers: one chunk per tap; decay layer: 8 chunks)
                __half* row = reinterpret_cast<__half*>(weffT_pack + L.weffT_off + (long long)co * (L.taps * L.cin_s) + (long long)rt * L.cin_s + (ci / 32) * 32);
                const int cw = ci % 32;
                const __half w16 = __float2half_rn(w);                 // == w unless |w| is below fp16's normal range; the lo half absorbs that
                const __half l16 = __float2half_rn((wf - __half2float(w16)) * PACK_SCALE);
                row[cw] = l16;
                row[32 + cw] = w16;
                if (weff_pack && L.taps == 27 && L.cin_s == 32 && L.cout_s == 32) {     // the data gradient's bf16 pair, row (tap, ci), K = co (the layout of weff): [w_a | w - w_a]
                    __nv_bfloat16* rowd = reinterpret_cast<__nv_bfloat16*>(weff_pack + L.weff_off + ((long long)rt * L.cin_s + ci) * L.cout_s);
                    const __nv_bfloat16 wa = __float2bfloat16_rn(wf);
                    rowd[co] = wa;
                    rowd[32 + co] = __float2bfloat16_rn(wf - __bfloat162float(wa));
                }
            }
        }
    }
}

// dL/dg = rn * sum(dW*v);  dL/dv = g*rn * (dW - v * rn^2 * sum(dW*v));  dL/dbias copied from the padded scratch
__global__ void __launch_bounds__(128) wn_bwd_kernel(const WnLayer* __restrict__ tab, int nlayers,
                                                     const float* __restrict__ params, const float* __restrict__ scale,
                                                     const float* __restrict__ dweff, const float* __restrict__ dbias_s,
                                                     float* __restrict__ grads, int block0) {
    pdl_grid_wait();
    __shared__ float red[4];
    int co;
    const WnLayer L = find_layer(tab, nlayers, blockIdx.x + block0, co);
    const int K = L.taps * L.cin;
    const float* v = params + L.v_off;
    const float* dw = dweff + L.weff_off;
    float dot = 0.f;
    for (int k = threadIdx.x; k < K; k += 128) {
        const int tap = row_tap(L, k / L.cin), ci = k % L.cin;
        dot = fmaf(dw[((long long)tap * L.cin_s + ci) * L.cout_s + co], v[(long long)k * L.cout + co], dot);
    }
    dot = block_sum_128(dot, red);
    const float sc = scale[L.scale_off + co], rn = scale[L.scale_off + L.cout + co];
    if (threadIdx.x == 0) {
        grads[L.g_off + co] = rn * dot;
        grads[L.b_off + co] = dbias_s[L.bias_s_off + co];
    }
    const float proj = dot * rn * rn;
    for (int k = threadIdx.x; k < K; k += 128) {
        const int tap = row_tap(L, k / L.cin), ci = k % L.cin;
        const float d = dw[((long long)tap * L.cin_s + ci) * L.cout_s + co];
        grads[L.v_off + (long long)k * L.cout + co] = sc * (d - v[(long long)k * L.cout + co] * proj);
    }
}

// TFA WeightNormalization first-call initialisation: g <- ||v||
__global__ void __launch_bounds__(128) g_from_v_kernel(const WnLayer* __restrict__ tab, int nlayers, float* __restrict__ params) {
    __shared__ float red[4];
    int co;
    const WnLayer L = find_layer(tab, nlayers, blockIdx.x, co);
    const int K = L.taps * L.cin;
    const float* v = params + L.v_off;
    float ss = 0.f;
    for (int k = threadIdx.x; k < K; k += 128) { const float x = v[(long long)k * L.cout + co]; ss = fmaf(x, x, ss); }
    ss = block_sum_128(ss, red);
    if (threadIdx.x == 0) params[L.g_off + co] = sqrtf(ss);
}

// ------------------------------------------------------------------------------------------ optimizers
__global__ void nadam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, NadamScalars s) {
    pdl_grid_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float g_hat = gi / s.one_minus_Pt;
    const float mi = s.b1 * m[i] + (1.f - s.b1) * gi;
    const float m_hat = mi / s.one_minus_Pt1;
    const float vi = s.b2 * v[i] + (1.f - s.b2) * gi * gi;
    const float v_hat = vi / s.one_minus_b2t;
    const float m_bar = (1.f - s.mu_t) * g_hat + s.mu_t1 * m_hat;
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - s.lr * m_bar / (sqrtf(v_hat) + s.eps);
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr_t, float b1, float b2, float eps) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, long long n, float lr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = p[i] - lr * g[i];
}

// ------------------------------------------------------------------------------------------ scene geometry
// patches[(s*n*n + i*n + j), y, x, t] = scene[s, t, refl(i*patch + y - pad), refl(j*patch + x - pad)]
__global__ void scene_to_patches_kernel(const float* __restrict__ scenes, long long ntot, int T, int H, int W,
                                        int patch, int pad, int S, int n, float* __restrict__ patches) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ntot) return;
    const int t = (int)(idx % T); long long r = idx / T;
    const int x = (int)(r % S); r /= S;
    const int y = (int)(r % S); r /= S;
    const int j = (int)(r % n); r /= n;
    const int i = (int)(r % n); const long long s = r / n;
    const int gy = refl(i * patch + y - pad, H), gx = refl(j * patch + x - pad, W);
    patches[idx] = __ldg(scenes + ((s * T + t) * H + gy) * W + gx);
}

// scenes[s, i*P + y, j*P + x] = sr[(s*n*n + i*n + j), y, x]
__global__ void stitch_kernel(const float* __restrict__ sr, long long ntot, int n, int P, float* __restrict__ scenes) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ntot) return;
    const int NP = n * P;
    const int X = (int)(idx % NP); long long r = idx / NP;
    const int Y = (int)(r % NP); const long long s = r / NP;
    scenes[idx] = __ldg(sr + (((s * n + Y / P) * n + X / P) * P + Y % P) * P + X % P);
}

}  // namespace

int launch_prep(const float* lr, int B, int HW, int T, float mean, float stdv, float* xn, float* mn, cudaStream_t st) {
    PV_TIMED("prep", st);
    const long long nvox = (long long)B * HW;
    PV_CUDA(launch_pdl_simple(prep_kernel, cdiv(nvox, 256), 256, 0, st, lr, nvox, T, mean, 1.0f / stdv, xn, mn));
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_reflect_pad(const float* in, float* out, int B, int H, int W, int T, int C, int ph, int pw, int pt, cudaStream_t st) {
    PV_TIMED("reflect_pad", st);
    if (C % 4) return set_error(PV_ERR_BAD_ARG, "reflect_pad: C %% 4 != 0");
    const long long n = (long long)B * (H + 2 * ph) * (W + 2 * pw) * (T + 2 * pt) * (C / 4);
    reflect_pad_kernel<<<cdiv(n, 256), 256, 0, st>>>(in, out, B, H, W, T, C / 4, ph, pw, pt);
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_reflect_pad_bwd(const float* gout, float* gin, int B, int H, int W, int T, int C, int ph, int pw, int pt, cudaStream_t st) {
    PV_TIMED("reflect_pad_bwd", st);
    if (C % 4) return set_error(PV_ERR_BAD_ARG, "reflect_pad_bwd: C %% 4 != 0");
    const long long n = (long long)B * H * W * T * (C / 4);
    reflect_pad_bwd_kernel<<<cdiv(n, 256), 256, 0, st>>>(gout, gin, B, H, W, T, C / 4, ph, pw, pt);
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_tail(const float* up, const float* resid, int B, int P, int scale, float mean, float stdv, int clip_round,
                float* sr, cudaStream_t st) {
    PV_TIMED("tail", st);
    const long long n = (long long)B * P * scale * P * scale;
    tail_kernel<<<cdiv(n, 256), 256, 0, st>>>(up, resid, n, P, scale, mean, stdv, clip_round, sr);
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_tail_bwd(const float* dsr, int B, int P, int scale, float stdv, float* dtail, cudaStream_t st) {
    PV_TIMED("tail_bwd", st);
    const long long n = (long long)B * P * P * scale * scale;
    tail_bwd_kernel<<<cdiv(n, 256), 256, 0, st>>>(dsr, n, P, scale, stdv, dtail);
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_wn_prep(const WnLayer* tab, int nlayers, int nblocks, const float* params, float* weff, float* weffT,
                   float* bias_s, float* scale, cudaStream_t st, float* weff_lo, float* weffT_lo, float* weffT_pack, float* weff_pack) {
    PV_TIMED("wn_prep", st);
    PV_CUDA(launch_pdl_simple(wn_prep_kernel, nblocks, 128, 0, st, tab, nlayers, params, weff, weffT, bias_s, scale, weff_lo, weffT_lo, weffT_pack, weff_pack));
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_wn_bwd(const WnLayer* tab, int nlayers, int nblocks, const float* params, const float* scale,
                  const float* dweff, const float* dbias_s, float* grads, cudaStream_t st, int block0) {
    if (nblocks <= 0) return 0;
    PV_TIMED("wn_bwd", st);
    PV_CUDA(launch_pdl_simple(wn_bwd_kernel, nblocks, 128, 0, st, tab, nlayers, params, scale, dweff, dbias_s, grads, block0));
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_g_from_v(const WnLayer* tab, int nlayers, int nblocks, float* params, cudaStream_t st) {
    PV_TIMED("g_from_v", st);
    g_from_v_kernel<<<nblocks, 128, 0, st>>>(tab, nlayers, params);
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_nadam(float* p, const float* g, float* m, float* v, long long n, NadamScalars s, cudaStream_t st) {
    PV_TIMED("nadam", st);
    PV_CUDA(launch_pdl_simple(nadam_kernel, cdiv(n, 256), 256, 0, st, p, g, m, v, n, s));
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2, float eps, cudaStream_t st) {
    PV_TIMED("adam", st);
    adam_kernel<<<cdiv(n, 256), 256, 0, st>>>(p, g, m, v, n, lr_t, b1, b2, eps);
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_sgd(float* p, const float* g, long long n, float lr, cudaStream_t st) {
    PV_TIMED("sgd", st);
    sgd_kernel<<<cdiv(n, 256), 256, 0, st>>>(p, g, n, lr);
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_scene_to_patches(const float* scenes, int ns, int T, int H, int W, int patch, int max_shift, float* patches, cudaStream_t st) {
    PV_TIMED("scene_to_patches", st);
    const int n = H / patch, S = patch + max_shift;
    const long long ntot = (long long)ns * n * n * S * S * T;
    scene_to_patches_kernel<<<cdiv(ntot, 256), 256, 0, st>>>(scenes, ntot, T, H, W, patch, max_shift / 2, S, n, patches);
    PV_LAUNCH_CHECK();
    return 0;
}

int launch_stitch(const float* sr, int ns, int n, int P, float* scenes, cudaStream_t st) {
    PV_TIMED("stitch", st);
    const long long ntot = (long long)ns * n * P * n * P;
    stitch_kernel<<<cdiv(ntot, 256), 256, 0, st>>>(sr, ntot, n, P, scenes);
    PV_LAUNCH_CHECK();
    return 0;
}

}  // namespace pv
