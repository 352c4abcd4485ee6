// skip2d.cu -- the learned low-frequency skip path (reference models/modelsTF.py:45-53 WDSRNetLRResidualPath: three
// weight-normalised 3x3 'valid' Conv2D, 1 -> C (ReLU) -> C -> C with C = scale^2, on the normalised temporal mean) as
// ONE kernel per direction.  The path is 0.02 % of the graph's FLOPs but ran as 11 generic launches (3 forward, 2 data
// gradient, 3 weight gradient with atomics, ...) costing 0.39 ms of a 7.6 ms step; here one CTA owns one patch, keeps
// the 22x22 input, all three activations and the weights in shared memory, and the weight gradients leave the CTA as
// one partial vector per patch that a fixed-order reduction (reduce.cuh) sums -- deterministic, no atomics.
//
// Layouts (dense engine conventions, kernels.h): activations [B, H, W, C] fp32; effective weights w[(tap*cin + ci)*C + co]
// with tap = a*3 + b (TF order); partial vector = { dW1[9*C], dW2[9*C*C], dW3[9*C*C], db1[C], db2[C], db3[C] }.
#include "kernels.h"
#include "rows.h"
#include "wgrad_reduce.cuh"

namespace pv {
namespace {

constexpr int SK_THREADS = 256;

// optional fused tail of the forward kernel (sr == nullptr: skip path only)
struct SkipTail {
    const float* u; RowGeom g; int uc;       // upscale conv output rows (G layout, one plane per patch), channels per row
    int scale; float mean, stdv; int clip_round;
    float* sr;                               // [B, scale*P, scale*P]
};

struct Skip2dShape {
    int S, C;                        // input side, channels (scale^2)
    __host__ __device__ int s1() const { return S - 2; }
    __host__ __device__ int s2() const { return S - 4; }
    __host__ __device__ int s3() const { return S - 6; }
    __host__ __device__ int nw1() const { return 9 * C; }
    __host__ __device__ int nw2() const { return 9 * C * C; }
    __host__ __device__ int npart() const { return skip2d_npart(C); }
};

// out[p][co] = act(b[co] + sum_{tap,ci} in[p + tap][ci] * w[tap][ci][co]) over a So x So output from an (So+2)^2 input
template <int CT, int CINT>
__device__ __forceinline__ void conv3x3_valid(const float* __restrict__ in, int Si, int cin_rt, const float* __restrict__ w,
                                              const float* __restrict__ b, int C_rt, bool relu, float* __restrict__ out) {
    const int C = CT ? CT : C_rt, cin = CINT ? CINT : cin_rt;      // compile-time counts let the ci loop unroll: independent LDS, no latency chain
    const int So = Si - 2;
    for (int idx = threadIdx.x; idx < So * So * C; idx += SK_THREADS) {
        const int co = idx % C, p = idx / C, h = p / So, x = p % So;
        float acc = b[co];
        float acc1 = 0.f, acc2 = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float* ip = in + ((h + a) * Si + x) * cin;
            const float* wp = w + (a * 3) * cin * C + co;
#pragma unroll
            for (int ci = 0; ci < cin; ++ci) {
                acc = fmaf(ip[ci], wp[ci * C], acc);
                acc1 = fmaf(ip[cin + ci], wp[(cin + ci) * C], acc1);
                acc2 = fmaf(ip[2 * cin + ci], wp[(2 * cin + ci) * C], acc2);
            }
        }
        acc += acc1 + acc2;
        out[idx] = relu ? fmaxf(acc, 0.f) : acc;
    }
}

__global__ void __launch_bounds__(SK_THREADS)
skip2d_fwd_kernel(const float* __restrict__ mn, const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                  const float* __restrict__ b2, const float* __restrict__ w3, const float* __restrict__ b3, Skip2dShape sh,
                  float* __restrict__ q1, float* __restrict__ q2, float* __restrict__ q3, SkipTail tl) {
    pdl_grid_wait();
    extern __shared__ float sm[];
    const int S = sh.S, C = sh.C, n0 = S * S, n1 = sh.s1() * sh.s1() * C, n2 = sh.s2() * sh.s2() * C, n3 = sh.s3() * sh.s3() * C;
    float* x = sm; float* a1 = x + n0; float* a2 = a1 + n1; float* a3 = a2 + n2;
    float* sw1 = a3 + n3; float* sw2 = sw1 + sh.nw1(); float* sw3 = sw2 + sh.nw2(); float* sb = sw3 + sh.nw2();
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < n0; i += SK_THREADS) x[i] = mn[(size_t)b * n0 + i];
    for (int i = threadIdx.x; i < sh.nw1(); i += SK_THREADS) sw1[i] = w1[i];
    for (int i = threadIdx.x; i < sh.nw2(); i += SK_THREADS) { sw2[i] = w2[i]; sw3[i] = w3[i]; }
    if (threadIdx.x < C) { sb[threadIdx.x] = b1[threadIdx.x]; sb[C + threadIdx.x] = b2[threadIdx.x]; sb[2 * C + threadIdx.x] = b3[threadIdx.x]; }
    __syncthreads();
    if (C == 9) {
        conv3x3_valid<9, 1>(x, S, 1, sw1, sb, C, true, a1);
        __syncthreads();
        conv3x3_valid<9, 9>(a1, sh.s1(), C, sw2, sb + C, C, false, a2);
        __syncthreads();
        conv3x3_valid<9, 9>(a2, sh.s2(), C, sw3, sb + 2 * C, C, false, a3);
    } else {
        conv3x3_valid<0, 0>(x, S, 1, sw1, sb, C, true, a1);
        __syncthreads();
        conv3x3_valid<0, 0>(a1, sh.s1(), C, sw2, sb + C, C, false, a2);
        __syncthreads();
        conv3x3_valid<0, 0>(a2, sh.s2(), C, sw3, sb + 2 * C, C, false, a3);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n1; i += SK_THREADS) q1[(size_t)b * n1 + i] = a1[i];
    for (int i = threadIdx.x; i < n2; i += SK_THREADS) q2[(size_t)b * n2 + i] = a2[i];
    for (int i = threadIdx.x; i < n3; i += SK_THREADS) q3[(size_t)b * n3 + i] = a3[i];
    if (tl.sr) {
        // Fused tail (modelsTF.py:38-41,52,73; test.py:118-119): depth_to_space of the upscale conv's output U and of this path's a3,
        // add, de-normalise [clip, round half-even] -- the skip path's result never leaves shared memory.  Threads sweep the
        // (scale * P)^2 output row-major, so the SR stores are coalesced; the 9 of 32 channels of a U row are read once each.
        const int P = sh.s3(), s = tl.scale, PS = P * s;
        for (int i = threadIdx.x; i < PS * PS; i += SK_THREADS) {
            const int X = i % PS, Y = i / PS, c = (Y % s) * s + (X % s);
            const long long urow = tl.g.lead + (long long)b * tl.g.pstride + (long long)tl.g.t0 * tl.g.plane + (Y / s) * tl.g.pw + X / s;
            float v = (__ldg(tl.u + urow * tl.uc + c) + a3[((Y / s) * P + X / s) * C + c]) * tl.stdv + tl.mean;
            if (tl.clip_round) v = rintf(fminf(fmaxf(v, 0.f), 65536.f));
            tl.sr[(size_t)b * PS * PS + i] = v;
        }
    }
}

// gin[p][ci] = sum_{tap,co} g[p - tap][co] * w[tap][ci][co]   (full correlation: So = Sg + 2), optionally * (ref > 0)
template <int CT>
__device__ __forceinline__ void dgrad3x3(const float* __restrict__ g, int Sg, const float* __restrict__ w, int C_rt,
                                         const float* __restrict__ relu_ref, float* __restrict__ gin) {
    const int C = CT ? CT : C_rt;
    const int So = Sg + 2;
    for (int idx = threadIdx.x; idx < So * So * C; idx += SK_THREADS) {
        const int ci = idx % C, p = idx / C, h = p / So, x = p % So;
        float acc = 0.f, acc1 = 0.f, acc2 = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int hh = h - a;
            if (hh < 0 || hh >= Sg) continue;
            const float* gp = g + (hh * Sg + x) * C;
            const float* wp = w + ((a * 3) * C + ci) * C;
            const bool ok0 = x < Sg, ok1 = x >= 1 && x - 1 < Sg, ok2 = x >= 2;
#pragma unroll
            for (int co = 0; co < C; ++co) {
                if (ok0) acc = fmaf(gp[co], wp[co], acc);
                if (ok1) acc1 = fmaf(gp[co - C], wp[C * C + co], acc1);
                if (ok2) acc2 = fmaf(gp[co - 2 * C], wp[2 * C * C + co], acc2);
            }
        }
        acc += acc1 + acc2;
        if (relu_ref && !(relu_ref[idx] > 0.f)) acc = 0.f;
        gin[idx] = acc;
    }
}

// dw[tap][ci][co] = sum_p in[p + tap][ci] * g[p][co],  db[co] = sum_p g[p][co]   (g is Sg x Sg, in is (Sg+2)^2)
template <int CT>
__device__ __forceinline__ void wgrad3x3(const float* __restrict__ in, int cin, const float* __restrict__ g, int Sg, int C_rt,
                                         float* __restrict__ dw, float* __restrict__ db) {
    const int C = CT ? CT : C_rt;
    const int Si = Sg + 2, nw = 9 * cin * C;
    for (int idx = threadIdx.x; idx < nw + C; idx += SK_THREADS) {
        float acc = 0.f;
        if (idx < nw) {
            const int co = idx % C, ci = (idx / C) % cin, tap = idx / (C * cin), a = tap / 3, bb = tap % 3;
            float acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;      // four independent chains over the pixels of a line
            for (int h = 0; h < Sg; ++h) {
                const float* ip = in + ((h + a) * Si + bb) * cin + ci;
                const float* gp = g + (h * Sg) * C + co;
                int x = 0;
                for (; x + 3 < Sg; x += 4) {
                    acc = fmaf(ip[x * cin], gp[x * C], acc);
                    acc1 = fmaf(ip[(x + 1) * cin], gp[(x + 1) * C], acc1);
                    acc2 = fmaf(ip[(x + 2) * cin], gp[(x + 2) * C], acc2);
                    acc3 = fmaf(ip[(x + 3) * cin], gp[(x + 3) * C], acc3);
                }
                for (; x < Sg; ++x) acc = fmaf(ip[x * cin], gp[x * C], acc);
            }
            dw[idx] = (acc + acc1) + (acc2 + acc3);
        } else {
            const int co = idx - nw;
            for (int p = 0; p < Sg * Sg; ++p) acc += g[p * C + co];
            db[co] = acc;
        }
    }
}

__global__ void __launch_bounds__(SK_THREADS)
skip2d_bwd_kernel(const float* __restrict__ mn, const float* __restrict__ q1, const float* __restrict__ q2, const float* __restrict__ g3,
                  const float* __restrict__ w2, const float* __restrict__ w3, Skip2dShape sh, float* __restrict__ partials) {
    pdl_grid_wait();
    extern __shared__ float sm[];
    const int S = sh.S, C = sh.C, n0 = S * S, n1 = sh.s1() * sh.s1() * C, n2 = sh.s2() * sh.s2() * C, n3 = sh.s3() * sh.s3() * C;
    float* x = sm; float* a1 = x + n0; float* a2 = a1 + n1; float* gg3 = a2 + n2;
    float* gg2 = gg3 + n3; float* gg1 = gg2 + n2;
    float* sw2 = gg1 + n1; float* sw3 = sw2 + sh.nw2();
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < n0; i += SK_THREADS) x[i] = mn[(size_t)b * n0 + i];
    for (int i = threadIdx.x; i < n1; i += SK_THREADS) a1[i] = q1[(size_t)b * n1 + i];
    for (int i = threadIdx.x; i < n2; i += SK_THREADS) a2[i] = q2[(size_t)b * n2 + i];
    for (int i = threadIdx.x; i < n3; i += SK_THREADS) gg3[i] = g3[(size_t)b * n3 + i];
    for (int i = threadIdx.x; i < sh.nw2(); i += SK_THREADS) { sw2[i] = w2[i]; sw3[i] = w3[i]; }
    __syncthreads();
    float* out = partials + (size_t)b * sh.npart();
    float* dw1 = out; float* dw2 = dw1 + sh.nw1(); float* dw3 = dw2 + sh.nw2(); float* db = dw3 + sh.nw2();
    if (C == 9) {
        dgrad3x3<9>(gg3, sh.s3(), sw3, C, nullptr, gg2);
        wgrad3x3<9>(a2, C, gg3, sh.s3(), C, dw3, db + 2 * C);
        __syncthreads();
        dgrad3x3<9>(gg2, sh.s2(), sw2, C, a1, gg1);           // flows into residConv1's ReLU
        wgrad3x3<9>(a1, C, gg2, sh.s2(), C, dw2, db + C);
        __syncthreads();
        wgrad3x3<9>(x, 1, gg1, sh.s1(), C, dw1, db);
    } else {
        dgrad3x3<0>(gg3, sh.s3(), sw3, C, nullptr, gg2);
        wgrad3x3<0>(a2, C, gg3, sh.s3(), C, dw3, db + 2 * C);
        __syncthreads();
        dgrad3x3<0>(gg2, sh.s2(), sw2, C, a1, gg1);
        wgrad3x3<0>(a1, C, gg2, sh.s2(), C, dw2, db + C);
        __syncthreads();
        wgrad3x3<0>(x, 1, gg1, sh.s1(), C, dw1, db);
    }
    for (int i = sh.nw1() + 2 * sh.nw2() + 3 * C + threadIdx.x; i < sh.npart(); i += SK_THREADS) out[i] = 0.f;
}

__global__ void __launch_bounds__(256)
skip2d_reduce_kernel(const float* __restrict__ partials, int B, int C, float* __restrict__ dw1, float* __restrict__ dw2,
                     float* __restrict__ dw3, float* __restrict__ db1, float* __restrict__ db2, float* __restrict__ db3) {
    __shared__ float4 smr[256];
    skip2d_reduce_body(blockIdx.x, partials, B, C, dw1, dw2, dw3, db1, db2, db3, smr);
}

}  // namespace

bool skip2d_supported(int S, int C) { return S >= 8 && S <= 40 && C >= 1 && C <= 16; }

int launch_skip2d_fwd(const float* mn, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                      const float* b3, int B, int S, int C, float* q1, float* q2, float* q3, cudaStream_t st) {
    SkipTail tl;
    memset(&tl, 0, sizeof tl);
    return launch_skip2d_fwd_tail(mn, w1, b1, w2, b2, w3, b3, B, S, C, q1, q2, q3, nullptr, tl.g, 0, 0, 0.f, 1.f, 0, nullptr, st);
}

// the same with the graph's tail fused in: sr = (depth_to_space(U[:, :9]) + depth_to_space(skip path)) * std + mean [clip, round]
int launch_skip2d_fwd_tail(const float* mn, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                           const float* b3, int B, int S, int C, float* q1, float* q2, float* q3, const float* u, RowGeom ug, int uc,
                           int scale, float mean, float stdv, int clip_round, float* sr, cudaStream_t st) {
    Skip2dShape sh{S, C};
    if (sr && (scale * scale != C || !u)) return set_error(PV_ERR_BAD_ARG, "skip2d_fwd_tail: scale^2 must equal the path's channel count");
    SkipTail tl;
    memset(&tl, 0, sizeof tl);
    tl.u = u; tl.g = ug; tl.uc = uc; tl.scale = scale; tl.mean = mean; tl.stdv = stdv; tl.clip_round = clip_round; tl.sr = sr;
    const size_t smem = sizeof(float) * ((size_t)S * S + (size_t)(sh.s1() * sh.s1() + sh.s2() * sh.s2() + sh.s3() * sh.s3()) * C +
                                         sh.nw1() + 2 * sh.nw2() + 3 * C);
    static size_t attr[16] = {};
    PV_CUDA(ensure_dyn_smem(skip2d_fwd_kernel, smem, attr));
    PV_TIMED(sr ? "skip2d_fwd_tail" : "skip2d_fwd", st, 2.0 * B * C * (9.0 * sh.s1() * sh.s1() + 9.0 * C * (sh.s2() * sh.s2() + sh.s3() * sh.s3())),
             sr ? (double)B * sh.s3() * sh.s3() * C * 12.0 : 0.0);
    PV_CUDA(launch_pdl_simple(skip2d_fwd_kernel, B, SK_THREADS, smem, st, mn, w1, b1, w2, b2, w3, b3, sh, q1, q2, q3, tl));
    PV_LAUNCH_CHECK();
    return 0;
}

size_t skip2d_partial_floats(int B, int S, int C) { return (size_t)B * Skip2dShape{S, C}.npart(); }

int launch_skip2d_bwd(const float* mn, const float* q1, const float* q2, const float* g3, const float* w2, const float* w3,
                      int B, int S, int C, float* partials, size_t partial_floats, float* dw1, float* dw2, float* dw3,
                      float* db1, float* db2, float* db3, cudaStream_t st, ReduceQueue* rq) {
    Skip2dShape sh{S, C};
    float* deferred = rq ? rq->take(skip2d_partial_floats(B, S, C)) : nullptr;
    if (deferred) { partials = deferred; partial_floats = skip2d_partial_floats(B, S, C); }
    if (!partials || partial_floats < skip2d_partial_floats(B, S, C)) return set_error(PV_ERR_BAD_ARG, "skip2d_bwd: partial buffer too small");
    const size_t smem = sizeof(float) * ((size_t)S * S + (size_t)(2 * sh.s1() * sh.s1() + 2 * sh.s2() * sh.s2() + sh.s3() * sh.s3()) * C + 2 * sh.nw2());
    static size_t attr[16] = {};
    PV_CUDA(ensure_dyn_smem(skip2d_bwd_kernel, smem, attr));
    {
        PV_TIMED("skip2d_bwd", st, 4.0 * B * C * (9.0 * sh.s1() * sh.s1() + 9.0 * C * (sh.s2() * sh.s2() + sh.s3() * sh.s3())), 0.0);
        PV_CUDA(launch_pdl_simple(skip2d_bwd_kernel, B, SK_THREADS, smem, st, mn, q1, q2, g3, w2, w3, sh, partials));
        PV_LAUNCH_CHECK();
    }
    if (deferred) {
        ReduceJob j;
        memset(&j, 0, sizeof j);
        j.kind = 3; j.nblocks = skip2d_reduce_blocks(C); j.partials = partials; j.ncta = B; j.C = C; j.S = S;
        j.o0 = dw1; j.o1 = dw2; j.o2 = dw3; j.o3 = db1; j.o4 = db2; j.o5 = db3;
        rq->push(j);
    } else {
        PV_TIMED("wgrad_reduce", st);
        skip2d_reduce_kernel<<<skip2d_reduce_blocks(C), 256, 0, st>>>(partials, B, C, dw1, dw2, dw3, db1, db2, db3);
        PV_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace pv
