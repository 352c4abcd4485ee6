// shift_loss.cu -- shift-search, clearance-masked, bias-corrected L1 / L2 / cPSNR, forward + fused backward.
//
// Replaces reference models/loss.py: shiftCompensatedL1Loss :73-84, shiftCompensatedL2Loss :55-71,
// shiftCompensatedcPSNR :37-53 and their helpers stackL1Loss :140-152, stackL2Loss :154-166,
// stackcPSNR :168-180, computeBiasBrightness :182-187, computeL1Loss :226-228, computeL2Loss :230-232,
// computecPSNR :234-238, cropImage utils/utils.py:42-44.  The reference unrolls 49 shifts into ~800
// tiny TF kernels per call (and again for the metric); here ONE CTA per 48x48 sample stages HR, mask and
// SR in shared memory once and evaluates all 49 shifts, L1 and L2 together, then emits dL/dSR of the
// winning shift in the same launch (closed form).
//
// Work decomposition (per 42x42 crop tile): thread t<252 owns the 7-pixel strip (row t/6, cols 7*(t%6)..+6)
// of the fixed SR crop in registers.  For a row shift i it reads 13 (h,m) pairs of HR row (r+i) once and
// reuses them for the 7 column shifts j -> 0.27 shared loads per (pixel,shift).
//   pass 1: N_ij = sum m, sum h, sum p*m         -> bias b_ij                        (loss.py:143-146,184)
//   pass 2: sum |h-(p+b)m|, sum (h-(p+b)m)^2     -> L1_ij, L2_ij, cPSNR_ij           (loss.py:148-151)
//   pass 3: gradient at the first arg-min shift                                        (Appendix C.3)
// Precision: inputs are centred by c = SR[b,3,3] (h' = h - c*m, p' = p - c), an exact re-parametrisation
// of the reference formulas (b and r are unchanged) that keeps fp32 partial sums small.
// Larger targets (evaluate.py:76-87 scores 384x384 scenes) are tiled 42x42 over a (tiles, samples) grid with a
// two-phase reduction (pass1 partials -> bias -> pass2 partials -> finalize).
#include "common.cuh"

namespace pv {
namespace sl {

constexpr int S = 7;          // shifts per axis = 2*cropBorder+1 (loss.py:18,48-49)
constexpr int NS = S * S;     // 49
constexpr int BORDER = 3;     // Losses.cropBorder default, never overridden (loss.py:13, train.py:87)
constexpr int CT = 42;        // crop tile side
constexpr int WT = CT + 2 * BORDER;   // 48: HR window side of one tile
constexpr int PX = 7;         // pixels per thread (one strip of a crop row)
constexpr int TPR = CT / PX;  // 6 threads per crop row
constexpr int NT = 256;       // threads per CTA, 252 of them own a strip
constexpr int NW = NT / 32;
constexpr int RS = 58;        // smem row stride (float2) of the window: conflict-free for the strip pattern
constexpr int HMW = PX + S - 1;       // 13 window pixels feed one strip over 7 column shifts
constexpr int RT = CT;                // row stride of the residual tile: 42 = 10 (mod 32) makes the strip pattern (thread = row t / 6,
                                      // columns 7 (t % 6) + k) hit 32 different banks per warp instruction (43 gave 2-way conflicts: 36 % of the
                                      // L1Edge kernel's shared-memory wavefronts, profiles/r02_ncu_shift_loss_l1edge_b65536.md)

struct __align__(16) Smem {
    float2 hm[WT * RS];       // (h', m) of the HR window; reused as the dSR tile in pass 3
    float part[NW][NS][3];    // per-warp partial sums
    float bias[NS];
    float cnt[NS];
    float l1[NS];
    float l2[NS];
    float red[NW];
    float center;
    int best;
    float edge[NS];           // sobel term of the L1Edge loss per shift
    float rt[2][CT * RT];     // residual tile r = h - (p + b) m of one shift (double buffered); rt[1] doubles as the adjoint tile
};

// Reduce N (power of two <= 32) per-lane values across the warp with N-1 + log2(32/N) shuffles:
// afterwards lane l holds in v[0] the warp total of value index l / (32/N).
template <int N>
__device__ __forceinline__ void warp_reduce_multi(float (&v)[N], int lane) {
    int o = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1, o >>= 1) {
        const int half = n >> 1;
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const float send = upper ? v[k] : v[k + half];
            const float keep = upper ? v[k + half] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    for (; o >= 1; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
}

// ---------------------------------------------------------------------------------------------- staging
// Fill the (h', m) window of tile (ty, tx) of sample b and this thread's SR strip p'.
// vx[x] = 1 where crop pixel x of the strip exists (partial tiles of big targets), else 0.
template <bool FULL>
__device__ __forceinline__ void stage(Smem& s, const float* __restrict__ hr, const uint8_t* __restrict__ mask,
                                      const float* __restrict__ sr, int H, int W, int ty, int tx,
                                      float (&p)[PX], float (&vx)[PX], bool& active, int& r, int& c0) {
    const int t = threadIdx.x;
    if (t == 0) s.center = sr[BORDER * W + BORDER];
    __syncthreads();
    const float c = s.center;
    if (FULL) {
        // the whole 48x48 sample is the window: contiguous, 16-byte aligned rows
        const float4* h4 = reinterpret_cast<const float4*>(hr);
        const uchar4* m4 = reinterpret_cast<const uchar4*>(mask);
        for (int q = t; q < WT * WT / 4; q += NT) {
            const float4 hv = __ldg(h4 + q);
            const uchar4 mv = __ldg(m4 + q);
            const int row = (q * 4) / WT, col = (q * 4) % WT;
            float2* dst = &s.hm[row * RS + col];
            const float m0 = mv.x ? 1.f : 0.f, m1 = mv.y ? 1.f : 0.f, m2 = mv.z ? 1.f : 0.f, m3 = mv.w ? 1.f : 0.f;
            dst[0] = make_float2(hv.x - c * m0, m0);
            dst[1] = make_float2(hv.y - c * m1, m1);
            dst[2] = make_float2(hv.z - c * m2, m2);
            dst[3] = make_float2(hv.w - c * m3, m3);
        }
    } else {
        const int y0 = ty * CT, x0 = tx * CT;
        for (int q = t; q < WT * WT; q += NT) {
            const int row = q / WT, col = q % WT;
            const int gy = y0 + row, gx = x0 + col;
            float h = 0.f, m = 0.f;
            if (gy < H && gx < W) {
                m = mask[(size_t)gy * W + gx] ? 1.f : 0.f;
                h = hr[(size_t)gy * W + gx] - c * m;
            }
            s.hm[row * RS + col] = make_float2(h, m);
        }
    }
    active = t < CT * TPR;
    r = active ? t / TPR : 0;
    c0 = active ? (t % TPR) * PX : 0;
    const int cropH = H - 2 * BORDER, cropW = W - 2 * BORDER;
    const int gy = ty * CT + r;
#pragma unroll
    for (int x = 0; x < PX; ++x) {
        const int gx = tx * CT + c0 + x;
        const bool ok = active && gy < cropH && gx < cropW;
        vx[x] = ok ? 1.f : 0.f;
        p[x] = ok ? sr[(size_t)(gy + BORDER) * W + (gx + BORDER)] - c : 0.f;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------- pass 1
// per-warp partials of (N, sum h', sum p'*m) for all 49 shifts -> s.part
template <bool FULL>
__device__ __forceinline__ void pass1(Smem& s, const float (&p)[PX], const float (&vx)[PX], bool active, int r, int c0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int i = 0; i < S; ++i) {
        float n[8], sh[8], spm[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) n[j] = sh[j] = spm[j] = 0.f;
        if (active) {
            float2 w[HMW];
            const float2* row = &s.hm[(r + i) * RS + c0];
#pragma unroll
            for (int k = 0; k < HMW; ++k) w[k] = row[k];
            if (FULL) {
#pragma unroll
                for (int x = 0; x < PX; ++x) { n[0] += w[x].y; sh[0] += w[x].x; }
#pragma unroll
                for (int j = 1; j < S; ++j) {   // sliding 7-wide box
                    n[j] = n[j - 1] - w[j - 1].y + w[j + PX - 1].y;
                    sh[j] = sh[j - 1] - w[j - 1].x + w[j + PX - 1].x;
                }
#pragma unroll
                for (int j = 0; j < S; ++j)
#pragma unroll
                    for (int x = 0; x < PX; ++x) spm[j] = fmaf(p[x], w[x + j].y, spm[j]);
            } else {
#pragma unroll
                for (int j = 0; j < S; ++j)
#pragma unroll
                    for (int x = 0; x < PX; ++x) {
                        const float m = w[x + j].y * vx[x];
                        n[j] += m;
                        sh[j] = fmaf(w[x + j].x, vx[x], sh[j]);
                        spm[j] = fmaf(p[x], m, spm[j]);
                    }
            }
        }
        warp_reduce_multi<8>(n, lane);
        warp_reduce_multi<8>(sh, lane);
        warp_reduce_multi<8>(spm, lane);
        if ((lane & 3) == 0 && (lane >> 2) < S) {
            float* dst = s.part[warp][i * S + (lane >> 2)];
            dst[0] = n[0]; dst[1] = sh[0]; dst[2] = spm[0];
        }
    }
}

// ---------------------------------------------------------------------------------------------- pass 2
// per-warp partials of (sum |r|, sum r^2), r = h' - (p'+b)*m, for all 49 shifts -> s.part[..][0..1]
template <bool FULL>
__device__ __forceinline__ void pass2(Smem& s, const float (&p)[PX], const float (&vx)[PX], bool active, int r, int c0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int i = 0; i < S; ++i) {
        float a1[8], a2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
        if (active) {
            float2 w[HMW];
            const float2* row = &s.hm[(r + i) * RS + c0];
#pragma unroll
            for (int k = 0; k < HMW; ++k) w[k] = row[k];
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const float b = s.bias[i * S + j];
#pragma unroll
                for (int x = 0; x < PX; ++x) {
                    float t = fmaf(-(p[x] + b), w[x + j].y, w[x + j].x);
                    if (!FULL) t *= vx[x];
                    a1[j] += fabsf(t);
                    a2[j] = fmaf(t, t, a2[j]);
                }
            }
        }
        warp_reduce_multi<8>(a1, lane);
        warp_reduce_multi<8>(a2, lane);
        if ((lane & 3) == 0 && (lane >> 2) < S) {
            float* dst = s.part[warp][i * S + (lane >> 2)];
            dst[0] = a1[0]; dst[1] = a2[0];
        }
    }
}

__device__ __forceinline__ int reflect42(int i) { return i < 0 ? -i : (i >= CT ? 2 * CT - 2 - i : i); }

// ---------------------------------------------------------------------------------------------- pass 2 + sobel (L1Edge)
// Same sums as pass2<true> plus, per shift, sum |sobel_y(r)| + |sobel_x(r)| of the residual tile r = h' - (p' + b) m
// (tf.image.sobel_edges pads with REFLECT; loss.py:219-224 is linear in r).  The residual strip is already in registers for
// the L1 / L2 sums, so the only extra shared-memory traffic per shift is 7 stores + the 3 x 9 neighbourhood loads (round 1
// re-read the (h', m) window for a separate 49-shift sweep: 2.6x the wavefronts).  Results: s.part[warp][shift][0..2].
__device__ __forceinline__ void pass2_edge(Smem& s, const float (&p)[PX], bool active, int r, int c0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ym = reflect42(r - 1) * RT, y0 = r * RT, yp = reflect42(r + 1) * RT;
    const int xl = reflect42(c0 - 1), xr = reflect42(c0 + PX);
#pragma unroll 1
    for (int i = 0; i < S; ++i) {
        float a1[8], a2[8], ed[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a1[j] = a2[j] = ed[j] = 0.f;
        float2 w[HMW];
        if (active) {
            const float2* row = &s.hm[(r + i) * RS + c0];
#pragma unroll
            for (int k = 0; k < HMW; ++k) w[k] = row[k];
        }
#pragma unroll
        for (int j = 0; j < S; ++j) {
            float* R = s.rt[(i * S + j) & 1];       // strict alternation over the 49 shifts (S is odd)
            if (active) {
                const float b = s.bias[i * S + j];
#pragma unroll
                for (int x = 0; x < PX; ++x) {
                    const float t = fmaf(-(p[x] + b), w[x + j].y, w[x + j].x);
                    a1[j] += fabsf(t);
                    a2[j] = fmaf(t, t, a2[j]);
                    R[y0 + c0 + x] = t;
                }
            }
            __syncthreads();      // tile (i, j) complete; the other buffer was last read one shift ago, before this barrier's predecessor
            if (active) {
                float v[PX + 2], d[PX + 2];
#pragma unroll
                for (int k = 0; k < PX + 2; ++k) {
                    const int xc = k == 0 ? xl : (k == PX + 1 ? xr : c0 + k - 1);
                    const float a = R[ym + xc], b2 = R[y0 + xc], c2 = R[yp + xc];
                    v[k] = a + 2.f * b2 + c2;
                    d[k] = c2 - a;
                }
#pragma unroll
                for (int x = 0; x < PX; ++x) ed[j] += fabsf(d[x] + 2.f * d[x + 1] + d[x + 2]) + fabsf(v[x + 2] - v[x]);
            }
        }
        warp_reduce_multi<8>(a1, lane);
        warp_reduce_multi<8>(a2, lane);
        warp_reduce_multi<8>(ed, lane);
        if ((lane & 3) == 0 && (lane >> 2) < S) {
            float* dst = s.part[warp][i * S + (lane >> 2)];
            dst[0] = a1[0]; dst[1] = a2[0]; dst[2] = ed[0];
        }
    }
}

__device__ __forceinline__ float cpsnr_from_l2(float l2) {
    // loss.py:234-238: 10*log(65535^2/L2)/log(10)
    return 10.0f * (logf(65535.0f * 65535.0f / l2) / logf(10.0f));
}

// first arg-min over 49 values held in smem; executed by warp 0; result broadcast through s.best
__device__ __forceinline__ int argmin49(const float* v, int lane) {
    float best = v[lane];
    int bi = lane;
    if (lane + 32 < NS) {
        const float o = v[lane + 32];
        if (o < best || (best != best && o == o)) { best = o; bi = lane + 32; }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        const bool take = (ob < best) || (ob == best && oi < bi) || (best != best && ob == ob);
        if (take) { best = ob; bi = oi; }
    }
    return bi;
}

// ---------------------------------------------------------------------------------------------- L1Edge (sobel) pass
// tf.image.sobel_edges: REFLECT pad by 1, Ky = [[-1,-2,-1],[0,0,0],[1,2,1]], Kx = Ky^T.
// loss.py:219-224 takes sobel(h) - sobel((p+b) m) on the 42x42 crops = sobel(r) (linear), summed |.| over both directions.

__device__ __forceinline__ void sobel_at(const float* __restrict__ R, int y, int x, float& gy, float& gx) {
    const int ym = reflect42(y - 1) * RT, y0 = y * RT, yp = reflect42(y + 1) * RT;
    const int xm = reflect42(x - 1), xp = reflect42(x + 1);
    const float a = R[ym + xm], b = R[ym + x], c = R[ym + xp];
    const float d = R[y0 + xm], f = R[y0 + xp];
    const float g = R[yp + xm], h = R[yp + x], k = R[yp + xp];
    gy = (g + 2.f * h + k) - (a + 2.f * b + c);
    gx = (c + 2.f * f + k) - (a + 2.f * d + g);
}

// ---------------------------------------------------------------------------------------------- fused patch kernel
// one CTA per 48x48 sample (targetShape (48,48,1), the p16 configs: train.py:86-87)
// The loss kind is a template parameter: the sobel sweep of the L1Edge loss needs more registers, which must not cost the L1 / L2
// kernels their occupancy (with a run-time `kind` the L1 kernel went from 2.36 to 3.02 ms at 65 536 samples).
template <int kind>
__global__ void __launch_bounds__(NT, kind == PV_LOSS_L1EDGE ? 3 : 5)      // 42 KB of shared memory per CTA allow five per SM
shift_loss_patch_kernel(const float* __restrict__ hr, const uint8_t* __restrict__ mask,
                        const float* __restrict__ sr, float grad_scale, float* __restrict__ loss_ps,
                        int32_t* __restrict__ best_shift, int32_t* __restrict__ clear_count,
                        float* __restrict__ cpsnr_ps, float* __restrict__ dsr, float* __restrict__ stack_out) {
    __shared__ Smem s;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const size_t so = (size_t)b * WT * WT;
    float p[PX], vx[PX];
    bool active; int r, c0;
    stage<true>(s, hr + so, mask + so, sr + so, WT, WT, 0, 0, p, vx, active, r, c0);

    pass1<true>(s, p, vx, active, r, c0);
    __syncthreads();
    if (t < NS) {
        float N = 0.f, sh = 0.f, spm = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) { N += s.part[w][t][0]; sh += s.part[w][t][1]; spm += s.part[w][t][2]; }
        s.cnt[t] = N;
        s.bias[t] = (1.0f / N) * (sh - spm);          // loss.py:184 (centred form, identical value)
    }
    __syncthreads();
    if (kind == PV_LOSS_L1EDGE) pass2_edge(s, p, active, r, c0);       // L1, L2 and the sobel term of every shift in one sweep
    else pass2<true>(s, p, vx, active, r, c0);
    __syncthreads();
    if (t < NS) {
        float a1 = 0.f, a2 = 0.f, e = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) { a1 += s.part[w][t][0]; a2 += s.part[w][t][1]; e += s.part[w][t][2]; }
        const float inv = 1.0f / s.cnt[t];
        s.l1[t] = inv * a1;                            // loss.py:227
        s.l2[t] = inv * a2;                            // loss.py:231
        s.edge[t] = 0.7f * (inv * a1) + (1.0f - 0.7f) * (inv * e);   // loss.py:224, pi = 0.7 (loss.py:21)
        if (stack_out) {
            float* o = stack_out + ((size_t)b * NS + t) * 4;
            o[0] = s.l1[t]; o[1] = s.l2[t]; o[2] = s.cnt[t]; o[3] = s.bias[t];
        }
    }
    __syncthreads();
    if (warp == 0) {
        const float* sel = (kind == PV_LOSS_L2) ? s.l2 : (kind == PV_LOSS_L1EDGE ? s.edge : s.l1);
        const int bi = argmin49(sel, lane);            // reduce_min over the stack (loss.py:83 / :70 / :96)
        const int b2 = argmin49(s.l2, lane);           // max cPSNR <=> min L2 (loss.py:51)
        if (lane == 0) {
            s.best = bi;
            loss_ps[b] = sel[bi];
            best_shift[b] = bi;
            clear_count[b] = (int)s.cnt[bi];
            if (cpsnr_ps) cpsnr_ps[b] = cpsnr_from_l2(s.l2[b2]);
        }
    }
    if (dsr == nullptr) return;
    __syncthreads();

    // ---- pass 3: closed-form gradient at the winning shift
    const int bi = s.best, i = bi / S, j = bi % S;
    const float bias = s.bias[bi], N = s.cnt[bi];
    float q[PX], mm[PX], qe[PX];
    float sqm = 0.f;
#pragma unroll
    for (int x = 0; x < PX; ++x) qe[x] = 0.f;
    if (kind == PV_LOSS_L1EDGE) {
        // Adjoint of the sobel term, Q = Ky^T sign(Ky r) + Kx^T sign(Kx r) with the REFLECT padding folded in, as a GATHER
        // (round 1 scattered 8 shared-memory atomics per pixel: 6 of the kernel's 11 ms at 65 536 samples).  The two sign
        // tiles are stored as bytes on a domain padded by 2 (zeros); on the un-reflected domain Y, X in [-1, 42]
        //     Qp[Y][X] = (sy[Y-1][X-1] + 2 sy[Y-1][X] + sy[Y-1][X+1]) - (sy[Y+1][X-1] + 2 sy[Y+1][X] + sy[Y+1][X+1])
        //              + (sx[Y-1][X-1] + 2 sx[Y][X-1] + sx[Y+1][X-1]) - (sx[Y-1][X+1] + 2 sx[Y][X+1] + sx[Y+1][X+1])
        // and the reflection maps the virtual lines -1 -> 1 and 42 -> 40:  Q[u][v] = sum over the pre-images of (u, v).
        float* R = s.rt[0];
        constexpr int SGW = 48;                                  // row stride (bytes) of a sign tile, 46 rows
        signed char* SY = reinterpret_cast<signed char*>(s.rt[1]);
        signed char* SX = SY + 46 * SGW;
        for (int k = t; k < 2 * 46 * SGW / 4; k += NT) reinterpret_cast<int*>(s.rt[1])[k] = 0;
        if (active) {
            const float2* row = &s.hm[(r + i) * RS + c0 + j];
#pragma unroll
            for (int x = 0; x < PX; ++x) R[r * RT + c0 + x] = fmaf(-(p[x] + bias), row[x].y, row[x].x);
        }
        __syncthreads();
        if (active) {
#pragma unroll 1
            for (int x = 0; x < PX; ++x) {
                float gy, gx;
                sobel_at(R, r, c0 + x, gy, gx);
                SY[(r + 2) * SGW + c0 + x + 2] = gy > 0.f ? 1 : (gy < 0.f ? -1 : 0);
                SX[(r + 2) * SGW + c0 + x + 2] = gx > 0.f ? 1 : (gx < 0.f ? -1 : 0);
            }
        }
        __syncthreads();
        if (active) {
            auto qp = [&](int Y, int X) -> float {
                const signed char* a = SY + (Y + 1) * SGW + X + 1;    // element (Y - 1, X - 1) of the padded tile
                const signed char* c = SX + (Y + 1) * SGW + X + 1;
                const int vy = (a[0] + 2 * a[1] + a[2]) - (a[2 * SGW] + 2 * a[2 * SGW + 1] + a[2 * SGW + 2]);
                const int vxx = (c[0] + 2 * c[SGW] + c[2 * SGW]) - (c[2] + 2 * c[SGW + 2] + c[2 * SGW + 2]);
                return (float)(vy + vxx);
            };
#pragma unroll 1
            for (int x = 0; x < PX; ++x) {
                const int v = c0 + x;
                float acc = qp(r, v);
                if (r == 1) acc += qp(-1, v);
                if (r == CT - 2) acc += qp(CT, v);
                if (v == 1 || v == CT - 2) {
                    const int X2 = v == 1 ? -1 : CT;
                    acc += qp(r, X2);
                    if (r == 1) acc += qp(-1, X2);
                    if (r == CT - 2) acc += qp(CT, X2);
                }
                qe[x] = acc;
            }
        }
    }
    if (active) {
        const float2* row = &s.hm[(r + i) * RS + c0 + j];
#pragma unroll
        for (int x = 0; x < PX; ++x) {
            const float2 w = row[x];
            const float tt = fmaf(-(p[x] + bias), w.y, w.x);
            // L1: d|r|/dr = sign(r), sign(0) = 0 (tf.abs gradient);  L2: d r^2/dr = 2r;  L1Edge: 0.7 sign(r) + 0.3 sobel adjoint
            const float sg = tt > 0.f ? 1.f : (tt < 0.f ? -1.f : 0.f);
            q[x] = (kind == PV_LOSS_L2) ? 2.0f * tt : (kind == PV_LOSS_L1EDGE ? 0.7f * sg + (1.0f - 0.7f) * qe[x] : sg);
            mm[x] = w.y;
            sqm = fmaf(q[x], w.y, sqm);
        }
    } else {
#pragma unroll
        for (int x = 0; x < PX; ++x) q[x] = mm[x] = 0.f;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) sqm += __shfl_xor_sync(0xffffffffu, sqm, off);
    if (lane == 0) s.red[warp] = sqm;
    __syncthreads();      // also: every thread is done reading s.hm
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) tot += s.red[w];
    float* gt = reinterpret_cast<float*>(s.hm);        // [48][48] gradient tile, zero border
    for (int k = t; k < WT * WT; k += NT) gt[k] = 0.f;
    __syncthreads();
    if (active) {
        const float invN = 1.0f / N;
        const float corr = tot * invN;
#pragma unroll
        for (int x = 0; x < PX; ++x)
            gt[(r + BORDER) * WT + (c0 + x + BORDER)] = grad_scale * (mm[x] * invN) * (corr - q[x]);
    }
    __syncthreads();
    float4* d4 = reinterpret_cast<float4*>(dsr + so);
    const float4* g4 = reinterpret_cast<const float4*>(gt);
    for (int k = t; k < WT * WT / 4; k += NT) d4[k] = g4[k];
}

// ---------------------------------------------------------------------------------------------- tiled path
// pass-1 partials of one 42x42 crop tile: part1[b][tile][49][3]
__global__ void __launch_bounds__(NT)
tile_pass1_kernel(const float* __restrict__ hr, const uint8_t* __restrict__ mask, const float* __restrict__ sr,
                  int H, int W, int ntx, float* __restrict__ part1) {
    __shared__ Smem s;
    const int tile = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const size_t so = (size_t)b * H * W;
    float p[PX], vx[PX];
    bool active; int r, c0;
    stage<false>(s, hr + so, mask + so, sr + so, H, W, tile / ntx, tile % ntx, p, vx, active, r, c0);
    pass1<false>(s, p, vx, active, r, c0);
    __syncthreads();
    if (t < NS) {
        float N = 0.f, sh = 0.f, spm = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) { N += s.part[w][t][0]; sh += s.part[w][t][1]; spm += s.part[w][t][2]; }
        float* o = part1 + (((size_t)b * gridDim.x + tile) * NS + t) * 3;
        o[0] = N; o[1] = sh; o[2] = spm;
    }
}

// per sample: fixed-order sum of tile partials -> (N, bias) per shift: nb[b][49][2]
__global__ void tile_bias_kernel(const float* __restrict__ part1, int ntiles, float* __restrict__ nb) {
    const int b = blockIdx.x, t = threadIdx.x;
    if (t >= NS) return;
    float N = 0.f, sh = 0.f, spm = 0.f;
    for (int k = 0; k < ntiles; ++k) {
        const float* o = part1 + (((size_t)b * ntiles + k) * NS + t) * 3;
        N += o[0]; sh += o[1]; spm += o[2];
    }
    nb[((size_t)b * NS + t) * 2 + 0] = N;
    nb[((size_t)b * NS + t) * 2 + 1] = (1.0f / N) * (sh - spm);
}

__global__ void __launch_bounds__(NT)
tile_pass2_kernel(const float* __restrict__ hr, const uint8_t* __restrict__ mask, const float* __restrict__ sr,
                  int H, int W, int ntx, const float* __restrict__ nb, float* __restrict__ part2) {
    __shared__ Smem s;
    const int tile = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const size_t so = (size_t)b * H * W;
    float p[PX], vx[PX];
    bool active; int r, c0;
    if (t < NS) s.bias[t] = nb[((size_t)b * NS + t) * 2 + 1];
    stage<false>(s, hr + so, mask + so, sr + so, H, W, tile / ntx, tile % ntx, p, vx, active, r, c0);
    pass2<false>(s, p, vx, active, r, c0);
    __syncthreads();
    if (t < NS) {
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) { a1 += s.part[w][t][0]; a2 += s.part[w][t][1]; }
        float* o = part2 + (((size_t)b * gridDim.x + tile) * NS + t) * 2;
        o[0] = a1; o[1] = a2;
    }
}

__global__ void tile_final_kernel(int kind, const float* __restrict__ part2, int ntiles, const float* __restrict__ nb,
                                  float* __restrict__ loss_ps, int32_t* __restrict__ best_shift,
                                  int32_t* __restrict__ clear_count, float* __restrict__ cpsnr_ps,
                                  float* __restrict__ stack_out) {
    __shared__ float l1[NS], l2[NS];
    const int b = blockIdx.x, t = threadIdx.x;
    if (t < NS) {
        float a1 = 0.f, a2 = 0.f;
        for (int k = 0; k < ntiles; ++k) {
            const float* o = part2 + (((size_t)b * ntiles + k) * NS + t) * 2;
            a1 += o[0]; a2 += o[1];
        }
        const float N = nb[((size_t)b * NS + t) * 2];
        l1[t] = (1.0f / N) * a1;
        l2[t] = (1.0f / N) * a2;
        if (stack_out) {
            float* o = stack_out + ((size_t)b * NS + t) * 4;
            o[0] = l1[t]; o[1] = l2[t]; o[2] = N; o[3] = nb[((size_t)b * NS + t) * 2 + 1];
        }
    }
    __syncthreads();
    if (t < 32) {
        const float* sel = (kind == PV_LOSS_L2) ? l2 : l1;
        const int bi = argmin49(sel, t);
        const int b2 = argmin49(l2, t);
        if (t == 0) {
            loss_ps[b] = sel[bi];
            best_shift[b] = bi;
            clear_count[b] = (int)nb[((size_t)b * NS + bi) * 2];
            if (cpsnr_ps) cpsnr_ps[b] = cpsnr_from_l2(l2[b2]);
        }
    }
}

// fixed-order mean of n floats (tf.reduce_mean over the batch, loss.py:84; Keras Mean of the metric)
__global__ void mean_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
    __shared__ float red[256];
    float a = 0.f;
    for (int k = threadIdx.x; k < n; k += 256) a += v[k];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o >= 1; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0] / (float)n;
}

}  // namespace sl

int launch_mean(const float* v, int n, float* out, cudaStream_t st) {
    PV_TIMED("mean", st);
    sl::mean_kernel<<<1, 256, 0, st>>>(v, n, out);
    PV_LAUNCH_CHECK();
    return 0;
}

// Device-pointer implementation behind pv_shift_loss (include/probav_b200.h).
int shift_loss_device(int kind, const float* hr, const uint8_t* mask, const float* sr, int B, int H, int W,
                      int border, float grad_scale, float* loss_ps, int32_t* best_shift, int32_t* clear_count,
                      float* cpsnr_ps, float* mean_loss, float* dsr, float* stack_out, cudaStream_t st) {
    using namespace sl;
    if (!hr || !mask || !sr || !loss_ps || !best_shift || !clear_count)
        return set_error(PV_ERR_BAD_ARG, "pv_shift_loss: hr, mask, sr, loss_per_sample, best_shift, clear_count are required");
    if (border != BORDER)
        return set_error(PV_ERR_BAD_ARG, "pv_shift_loss: cropBorder=%d unsupported (reference default 3, loss.py:13)", border);
    if (kind != PV_LOSS_L1 && kind != PV_LOSS_L2 && kind != PV_LOSS_L1EDGE)
        return set_error(PV_ERR_BAD_ARG, "pv_shift_loss: unknown loss kind %d", kind);
    if (kind == PV_LOSS_L1EDGE && !(H == WT && W == WT))
        return set_error(PV_ERR_BAD_ARG, "pv_shift_loss: the sobel/L1 mix is built for 48x48 training targets only (the reference "
                                         "scores whole scenes with cPSNR, evaluate.py:76-87)");
    if (B <= 0 || H <= 2 * border || W <= 2 * border)
        return set_error(PV_ERR_BAD_ARG, "pv_shift_loss: bad shape B=%d H=%d W=%d", B, H, W);
    // algorithmic bytes per sample: HR f32 + SR f32 + bool mask read, dSR f32 written when fused
    PV_TIMED(H == WT && W == WT ? "shift_loss_patch" : "shift_loss_tiled", st, 0.0,
             (double)B * H * W * (9.0 + (dsr ? 4.0 : 0.0)));
    if (H == WT && W == WT &&
        ((reinterpret_cast<uintptr_t>(hr) | reinterpret_cast<uintptr_t>(sr) | reinterpret_cast<uintptr_t>(dsr)) & 15 ||
         reinterpret_cast<uintptr_t>(mask) & 3))
        return set_error(PV_ERR_BAD_ARG, "pv_shift_loss: hr/sr/dsr must be 16-byte aligned and mask 4-byte aligned");
    if (H == WT && W == WT) {
        if (kind == PV_LOSS_L1) shift_loss_patch_kernel<PV_LOSS_L1><<<B, NT, 0, st>>>(hr, mask, sr, grad_scale, loss_ps, best_shift, clear_count, cpsnr_ps, dsr, stack_out);
        else if (kind == PV_LOSS_L2) shift_loss_patch_kernel<PV_LOSS_L2><<<B, NT, 0, st>>>(hr, mask, sr, grad_scale, loss_ps, best_shift, clear_count, cpsnr_ps, dsr, stack_out);
        else shift_loss_patch_kernel<PV_LOSS_L1EDGE><<<B, NT, 0, st>>>(hr, mask, sr, grad_scale, loss_ps, best_shift, clear_count, cpsnr_ps, dsr, stack_out);
        PV_LAUNCH_CHECK();
    } else {
        if (dsr) return set_error(PV_ERR_BAD_ARG, "pv_shift_loss: fused backward is only built for 48x48 targets");
        const int nty = cdiv(H - 2 * BORDER, CT), ntx = cdiv(W - 2 * BORDER, CT), ntiles = nty * ntx;
        float *part, *nb;
        PV_CUDA(cudaMallocAsync(&part, (size_t)B * ntiles * NS * 3 * sizeof(float), st));
        PV_CUDA(cudaMallocAsync(&nb, (size_t)B * NS * 2 * sizeof(float), st));
        dim3 grid(ntiles, B);
        tile_pass1_kernel<<<grid, NT, 0, st>>>(hr, mask, sr, H, W, ntx, part);
        PV_LAUNCH_CHECK();
        tile_bias_kernel<<<B, 64, 0, st>>>(part, ntiles, nb);
        PV_LAUNCH_CHECK();
        tile_pass2_kernel<<<grid, NT, 0, st>>>(hr, mask, sr, H, W, ntx, nb, part);
        PV_LAUNCH_CHECK();
        tile_final_kernel<<<B, 64, 0, st>>>(kind, part, ntiles, nb, loss_ps, best_shift, clear_count, cpsnr_ps, stack_out);
        PV_LAUNCH_CHECK();
        PV_CUDA(cudaFreeAsync(part, st));
        PV_CUDA(cudaFreeAsync(nb, st));
    }
    if (mean_loss) PV_TRY(launch_mean(loss_ps, B, mean_loss, st));
    return 0;
}

}  // namespace pv
