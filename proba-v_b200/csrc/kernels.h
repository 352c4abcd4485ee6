// kernels.h -- launcher prototypes shared by the engine (host side of csrc/*.cu).
#pragma once
#include "common.cuh"

namespace pv {

// One stride-1 convolution over channels-last activations [B, H, W, T, C] (2-D layers use T = kt = 1).
// The same descriptor drives the forward conv, and -- with flipped/transposed weights and
// pad' = k-1-pad -- the data gradient.  `xmask`, when set, is the post-ReLU output of the layer whose
// gradient `x` is: x is taken as x * (xmask > 0)  (tf ReluGrad fused into the consumer).
struct ConvP {
    const float* x;         // [B,Hi,Wi,Ti,cin]
    const float* xmask;     // same shape as x, or nullptr
    const float* w;         // [kh*kw*kt*cin][cout]   (k index = tap*cin + ci, tap = (a*kw+b)*kt+c)
    const float* bias;      // [cout] or nullptr
    const float* residual;  // same shape as y, or nullptr (added before the optional ReLU... never both here)
    float* y;               // [B,Ho,Wo,To,cout]
    int B, Hi, Wi, Ti, Ho, Wo, To;
    int cin, cout, kh, kw, kt, ph, pw, pt;   // zero padding ph/pw/pt on each side
    int relu;
    int cin_r, cout_r;      // un-padded channel counts (algorithmic flop accounting only)
    const char* tag;        // kernel-class label for pv_timing_report
};

// Weight gradient of the same convolution: dw[k][n] += sum_m im2col(x)[m][k] * dy[m][n]; db[n] += sum_m dy[m][n].
// `ymask` is the layer's own post-ReLU output when it has a ReLU.  dw/db must be zeroed by the caller.
struct WgradP {
    const float* x;         // layer input  [B,Hi,Wi,Ti,cin]
    const float* dy;        // grad of layer output [B,Ho,Wo,To,cout]
    const float* ymask;     // layer output (post-ReLU) or nullptr
    float* dw;              // [kh*kw*kt*cin][cout]
    float* db;              // [cout]
    int B, Hi, Wi, Ti, Ho, Wo, To;
    int cin, cout, kh, kw, kt, ph, pw, pt;
    int cin_r, cout_r;
    const char* tag;
    // Deterministic mode: the row range is split over CTAs; each split stores its partial sums into
    // partials[split][(k | bias row)][n] and a second kernel adds the splits in a FIXED order, so the exact (fp32) engine is
    // bit-reproducible (round 1 accumulated the splits with atomicAdd).  nullptr / too small: the atomic path.
    float* partials = nullptr;
    size_t partial_floats = 0;
};
constexpr size_t WGRAD_PARTIAL_FLOATS = (size_t)148 * 4 * 64 * 64 + (size_t)148 * 8 * 256 + 4096;   // covers every split geometry of launch_wgrad

int launch_conv(const ConvP& p, cudaStream_t st);
int launch_wgrad(const WgradP& p, cudaStream_t st);

// (x-mean)/std and the normalised temporal mean (modelsTF.py:23-27)
int launch_prep(const float* lr, int B, int HW, int T, float mean, float stdv, float* xn, float* mn, cudaStream_t st);
// tf.pad(mode='reflect') on H, W, T (modelsTF.py:157-158 and the T=13/19 variants) and its adjoint
int launch_reflect_pad(const float* in, float* out, int B, int H, int W, int T, int C, int ph, int pw, int pt, cudaStream_t st);
int launch_reflect_pad_bwd(const float* gout, float* gin, int B, int H, int W, int T, int C, int ph, int pw, int pt, cudaStream_t st);
// depth_to_space(main) + depth_to_space(resid) -> add -> denormalise [-> clip(0,65536) -> round-half-even]
// (modelsTF.py:38-41,52,73; test.py:118-119)
int launch_tail(const float* up, const float* resid, int B, int P, int scale, float mean, float stdv, int clip_round,
                float* sr, cudaStream_t st);
int launch_tail_bwd(const float* dsr, int B, int P, int scale, float stdv, float* dtail, cudaStream_t st);

// weight normalisation (TFA WeightNormalization.call) for every layer in one launch, and its backward
struct WnLayer {
    long long v_off, g_off, b_off;            // offsets into the param / grad arenas
    long long weff_off, weffT_off, bias_s_off, scale_off;   // offsets into the derived arenas
    int taps, cin, cout, cin_s, cout_s;
    int first_block;                          // prefix sum of cout over layers
    int mode;                                 // 0: dense engine (TF tap order, weffT = flipped taps, [tap][co][ci])
                                              // 1: row engine (taps in (dt,dh,dw) order, weffT = plain transpose [co][Kflat])
    int round_tf32;                           // round the effective weights to tf32 (they only feed tensor-core MMAs)
};
// weff_lo / weffT_lo (nullable, row engine): w - tf32(w) in the layouts of weff / weffT (error-compensated forward, precision 4)
int launch_wn_prep(const WnLayer* table_dev, int nlayers, int nblocks, const float* params, float* weff, float* weffT,
                   float* bias_s, float* scale, cudaStream_t st, float* weff_lo = nullptr, float* weffT_lo = nullptr,
                   float* weffT_pack = nullptr,    // packed fp16 pair rows of the 3x3x3 layers' weights (rows.h PACK_SCALE), forward layout
                   float* weff_pack = nullptr);    // bf16 pair rows [w_a | w - w_a] in the data gradient's layout (rows (tap, ci), K = co)
int launch_wn_bwd(const WnLayer* table_dev, int nlayers, int nblocks, const float* params, const float* scale,
                  const float* dweff, const float* dbias_s, float* grads, cudaStream_t st, int block0 = 0);   // blocks [block0, block0 + nblocks): a layer sub-range
int launch_g_from_v(const WnLayer* table_dev, int nlayers, int nblocks, float* params, cudaStream_t st);

// optimizers over the flat arena (train.py:77-83)
struct NadamScalars { float lr, b1, b2, eps, mu_t, mu_t1, one_minus_Pt, one_minus_Pt1, one_minus_b2t; };
int launch_nadam(float* p, const float* g, float* m, float* v, long long n, NadamScalars s, cudaStream_t st);
int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2, float eps, cudaStream_t st);
int launch_sgd(float* p, const float* g, long long n, float lr, cudaStream_t st);

// scene <-> patch geometry on device (dataGenerator.py:108-121; test.py:149-160)
int launch_scene_to_patches(const float* scenes, int ns, int T, int H, int W, int patch, int max_shift, float* patches, cudaStream_t st);
int launch_stitch(const float* sr_patches, int ns, int n, int P, float* scenes, cudaStream_t st);

struct ReduceQueue;   // wgrad_reduce.cuh
// learned low-frequency skip path (modelsTF.py:45-53) fused: three 3x3 valid Conv2D 1 -> C (ReLU) -> C -> C, one CTA per patch (skip2d.cu)
bool skip2d_supported(int S, int C);
size_t skip2d_partial_floats(int B, int S, int C);
int launch_skip2d_fwd(const float* mn, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                      const float* b3, int B, int S, int C, float* q1, float* q2, float* q3, cudaStream_t st);
int launch_skip2d_bwd(const float* mn, const float* q1, const float* q2, const float* g3, const float* w2, const float* w3,
                      int B, int S, int C, float* partials, size_t partial_floats, float* dw1, float* dw2, float* dw3,
                      float* db1, float* db2, float* db3, cudaStream_t st, ReduceQueue* rq = nullptr);

int launch_mean(const float* v, int n, float* out, cudaStream_t st);
int shift_loss_device(int kind, const float* hr, const uint8_t* mask, const float* sr, int B, int H, int W,
                      int border, float grad_scale, float* loss_ps, int32_t* best_shift, int32_t* clear_count,
                      float* cpsnr_ps, float* mean_loss, float* dsr, float* stack_out, cudaStream_t st);

}  // namespace pv
