// rows.h -- the "row" activation layouts of the tensor-core engine (pv_cfg.precision = 1) and the descriptors of
// the convolutions expressed on them.
//
// An activation is a 2-D array [rows][C] fp32 (C = 32 or 256 channels, 128 B or 1 KB per row).  A row is one voxel.
// Rows are ordered so that every tap of a 3x3x3 convolution is a CONSTANT row offset, which lets the tensor-core
// kernels present each tap's A operand as a row-shifted view of one shared-memory tile (implicit GEMM without
// im2col), and lets zero padding be real zero rows in memory:
//
//   PR layout  ('same' trunk: mainConv1 output, the 12 residual blocks; reference modelsTF.py:58-60,177-189)
//       row(b,t,h,w) = lead + b*pstride + t*529 + h*23 + w        t in 0..8, h,w in 0..21
//       column w = 22 and row h = 22 of every plane are zero, plane t = 9 of every patch is zero (pstride = 10*529):
//       one shared zero row/column/plane serves as BOTH the left and right 'same' padding of its neighbours.
//       tap (dt,dh,dw) in {-1,0,1}^3  ->  offset dt*529 + dh*23 + dw.
//   G layout   (valid-conv tail: reducers + upscale conv on the reflect-padded block output; modelsTF.py:152-164)
//       row(b,t,h,w) = lead + b*pstride + (2+t)*576 + h*24 + w    planes of 24x24, two leading zero planes per patch
//       tap (dt,dh,dw) in {0,1,2}^3 -> offset dt*576 + dh*24 + dw (data gradient: the negated offsets, which is why
//       the leading planes and everything outside the valid extent are kept zero).
//
// Invariant: rows outside the valid extent (RowGeom) are zero in every buffer a convolution reads.
#pragma once
#include "common.cuh"

namespace pv {

struct RowGeom {
    long long lead;      // rows before patch 0
    long long pstride;   // rows per patch
    int plane, pw;       // rows per plane, rows per image line
    int t0;              // first data plane inside a patch
    int nt, nh, nw;      // valid extent
    int row0;            // first row (inside a patch) worth computing; multiple of 128
    int nrows;           // rows to compute per patch starting at row0 (tensor-core kernels round up to 128)
};

__host__ __device__ inline bool row_valid(const RowGeom& g, int r /* row inside the patch */) {
    const int t = r / g.plane - g.t0;
    const int q = r % g.plane;
    const int h = q / g.pw, w = q % g.pw;
    return t >= 0 && t < g.nt && h < g.nh && w < g.nw;
}

// Packed fp16 pair rows (error-compensated engine: what a compensated 3x3x3 convolution reads): a 128-byte row holds 64 halves,
//   activations  [ fp16(x_hi[0..31]) | fp16(PACK_SCALE * x_lo[0..31]) ]        weights (per output channel and tap)  [ fp16(PACK_SCALE * w_lo) | fp16(w_hi) ]
// so that one K = 64 dot product of the two rows is PACK_SCALE * (x_hi w_lo + x_lo w_hi): both correction products of
// x w ~= x_hi w_hi + x_lo w_hi + x_hi w_lo in one kind::f16 MMA chain with the descriptors of the fp32 rows (same 128-byte rows, same
// 32-byte K steps), and the hi halves alone (K steps 0, 1 of an activation row against K steps 2, 3 of a weight row) give the main product.
// hi = tf32(v) has 11 significant bits and is exact in fp16 for |v| in fp16's normal range (values are O(1) after normalisation); where
// fp16(hi) != hi the packed lo half absorbs the difference (lo' = v - fp16(hi)), so hi16 + lo' = v always.  |lo'| <= 2^-12 |v| is scaled
// by 2^12 into fp16's normal range (Ootomo & Yokota's scaling), so the corrections keep 11 bits: 2^-23 of the product.
// The backward pass uses bf16 pair rows instead, [ bf16(g) x 32 | bf16(g - bf16(g)) x 32 ], unscaled (gradients are far below fp16's range),
// for the gradient rows and for the weights of the split-weight 3x3x3 data gradients (conv3_tc.cu MODE 2).
constexpr float PACK_SCALE = 4096.0f;

constexpr int ROW_TAIL = 2048;   // rows allocated (and zero) after the last patch of every row buffer
constexpr int MAX_TAPS = 27;
constexpr int MAX_SLABS = 8;

// out[orow][n] = mask( act( sum_tap sum_k x[irow + off[tap]][c0[tap] + k] * w[tap][k][n] + bias[n] ) (+ res) ) (* relumask>0)
//   orow = og.lead + b*og.pstride + r,  irow = in_lead + b*in_pstride + r,   r in [og.row0, og.row0 + og.nrows)
struct RowConvP {
    const float* x; int xc;            // input rows, channels per input row (row stride)
    // weights: one 2-D fp32 matrix [w_rows][w_cols]; tap `t` uses the 32 x n block whose element (k, n) sits at
    //   w[(wr0[t] + (w_kmajor ? n : k)) * w_cols + wc0[t] + (w_kmajor ? k : n)]
    // w_kmajor = 1: rows are output channels, K contiguous (the tcgen05 B operand, loaded by TMA as a {32, n} box);
    // w_kmajor = 0: rows are K, output channels contiguous (what the CUDA-core kernel prefers).
    const float* w; int w_rows, w_cols, w_kmajor;
    int wr0[MAX_TAPS], wc0[MAX_TAPS];
    const float* bias;                 // [n] or nullptr
    const float* residual;             // [rows][n] same rows as y, or nullptr
    const float* residual2;            // conv3_tc, f16_pack = 1 only: a second addend of the same shape (the lo half of the skip connection)
    float* y_lo;                       // conv3_tc, f16_pack = 1 only: when set, y receives hi = tf32(v) and y_lo the remainder v - hi
    float* y_pack;                     // conv3_tc only: when set, also the result as a pair row: f16_pack = 1: fp16 pair of (hi, lo)
                                       //   (see PACK_SCALE); f16_pack = 2: bf16 pair of the un-rounded result
    int f16_pack;                      // conv3_tc only: 1 = x and w are packed fp16 pair rows; the kernel computes the whole compensated product
                                       //   x_hi w_hi + x_lo w_hi + x_hi w_lo from them with kind::f16 MMAs (main and correction accumulators);
                                       //   2 = x and w are bf16 pair rows [a | v - a], unscaled (a gradient and the weights): x w to 16 bits each;
                                       //   3 = fp16 pair rows, hi halves only: the single-pass product with K = 16 MMAs (inference)
    const float* relumask;             // [rows][n]: output multiplied by (relumask > 0), or nullptr
    float* y; int n;                   // output rows, channels per output row
    int B;
    long long in_lead, in_pstride;
    RowGeom og;
    int ntap; int kc;                  // taps, input channels consumed per tap (32)
    int off[MAX_TAPS];                 // row offset of each tap
    int c0[MAX_TAPS];                  // first input channel of each tap (K-chunks of a wide row; 0 for real taps)
    int relu;
    int round_tf32;                    // round the stored output to tf32 (round-to-nearest) -- it only feeds MMAs
    double flops;                      // algorithmic flops (timing report)
    const char* tag;
};

// dw[tap][k][n] += sum_rows x[irow + off[tap]][c0[tap] + k] * gz[orow][n];   db[n] += sum_rows gz[orow][n]
struct RowWgradP {
    const float* x; int xc;
    const float* gz; int n;
    float* dw; int dw_cols;            // element (tap, k, n) at dw[(dwr0[tap] + k) * dw_cols + dwc0[tap] + n]
    int dwr0[MAX_TAPS], dwc0[MAX_TAPS];
    float* db;                         // [n] or nullptr
    int B;
    long long in_lead, in_pstride;
    RowGeom og;
    int ntap; int kc;
    int off[MAX_TAPS];
    int c0[MAX_TAPS];
    double flops;
    const char* tag;
};

int launch_rowconv_simt(const RowConvP& p, cudaStream_t st);
int launch_rowwgrad_simt(const RowWgradP& p, cudaStream_t st);
int launch_rowconv_tc(const RowConvP& p, cudaStream_t st);          // tcgen05 + TMA implicit GEMM (conv_tc.cu)
struct ReduceQueue;   // wgrad_reduce.cuh: when given, the launcher defers its partial reduction to launch_deferred_reduce()
int launch_rowwgrad_tc(const RowWgradP& p, cudaStream_t st, float* partials, size_t partial_floats, ReduceQueue* rq = nullptr);
// 3x3x3 lattice convolutions on 32-channel rows with the dw taps folded into N = 96 (conv3_tc.cu); launch_rowconv_tc
// dispatches to it unless the environment variable PV_CONV3_N32 is set (A/B timing against the N = 32 formulation)
bool rowconv3_tc_supported(const RowConvP& p);
int launch_rowconv3_tc(const RowConvP& p, cudaStream_t st);

// fused expConv + ReLU + decConv of one residual block (resblock_tc.cu); the expanded tensor stays in TMEM.
// relu_bits: [rows][8] uint32 ReLU bit mask written by the forward (nullable) and consumed by the backward-data kernel.
int launch_resfront_fwd_tc(const float* x, const float* weT_exp, const float* weT_dec, const float* bias_e, const float* bias_d,
                           float* d, uint32_t* relu_bits, const RowGeom& g, int B, int round_tf32, double flops, cudaStream_t st,
                           int out_f16 = 0);      // inference: d receives fp16 pair rows [fp16(tf32(D)) | 0] (what conv3_tc MODE 3 reads)
int launch_resfront_bwd_data_tc(const float* gd, const float* w_dec, const float* w_exp, const uint32_t* relu_bits,
                                const float* residual, const float* relumask, float* ga, const RowGeom& g,
                                int B, int round_tf32, double flops, cudaStream_t st,
                                const float* w_dec_lo = nullptr, const float* w_exp_lo = nullptr,    // lo halves: two MMAs per product
                                float* ga_pack = nullptr);     // also the un-rounded result as bf16 pair rows (conv3_tc MODE 2's input)
// error-compensated forward (resblock_x3_tc.cu): operands and result as (hi, lo) row arrays, three MMAs per product
int launch_resfront_fwd_x3_tc(const float* x_hi, const float* x_lo, const float* weT_exp_hi, const float* weT_exp_lo,
                              const float* weT_dec_pack /* fp16 pair rows of the decay weights */, const float* bias_e, const float* bias_d,
                              float* d_hi, float* d_lo, uint32_t* relu_bits, uint32_t* relu_bits_t, const RowGeom& g, int B, double flops,
                              cudaStream_t st, int pack_out = 0);
int launch_resfront_bwd_weight_tc(const float* x, const float* gd, const float* weT_exp, const float* w_dec, const float* bias_e,
                                  float* dw_dec, float* dw_exp, float* db_exp, float* db_dec, const RowGeom& g, int B,
                                  float* partials, size_t partial_floats, double flops, cudaStream_t st, ReduceQueue* rq = nullptr,
                                  const uint32_t* relu_bits_t = nullptr);   // forward's transposed ReLU mask (resblock_x3_tc.cu) instead of sign(E)

// mainConv1 (Cin = 1, 27 taps, ReLU) from the normalised dense LR [B,S,S,T] (h,w,t order) into the PR layout
// y_lo (nullable): y receives tf32(v) and y_lo the remainder v - tf32(v) (error-compensated engine)
int launch_first_conv_pr(const float* xn, const float* w /*[27][32] taps in (dt,dh,dw) order*/, const float* bias, int B, int S, int T,
                         float* y, RowGeom g, cudaStream_t st, float* y_lo = nullptr, int round_tf32 = 0);
int launch_first_conv_pr_wgrad(const float* xn, const float* gz, int B, int S, int T, RowGeom g, float* dw, float* db,
                               float* partials, size_t partial_floats, cudaStream_t st, ReduceQueue* rq = nullptr);
// PR (or G) tensor -> G layout with the reducer's reflect padding (tf.pad REFLECT by `pad` = 0 or 1 on H and W), and its
// adjoint (optionally multiplied by (relumask > 0): the padded tensor was a ReLU output)
int launch_pr_to_g_reflect(const float* a, RowGeom pr, float* g0, RowGeom gg, int B, int C, cudaStream_t st, int pad = 1,
                           const float* a_lo = nullptr, float* g_pack = nullptr);   // (hi, lo) source -> hi rows + packed fp16 pair rows
// round_tf32: store the result rounded to nearest tf32 (it only feeds tensor-core MMAs, which would truncate it)
int launch_pr_to_g_reflect_bwd(const float* gg0, RowGeom gg, float* ga, RowGeom pr, int B, int C, cudaStream_t st, int pad = 1,
                               const float* relumask = nullptr, int round_tf32 = 0, float* ga_pack = nullptr);   // ga_pack: also as bf16 pair rows

// low-frequency skip path (three 3x3 Conv2D) with the graph's tail fused in (skip2d.cu); u / ug / uc describe the upscale conv's rows
int launch_skip2d_fwd_tail(const float* mn, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                           const float* b3, int B, int S, int C, float* q1, float* q2, float* q3, const float* u, RowGeom ug, int uc,
                           int scale, float mean, float stdv, int clip_round, float* sr, cudaStream_t st);
// tail on the row layouts: sr = (depth_to_space(U[:, :, :9]) + depth_to_space(resid)) * std + mean [clip, round]; and its adjoint
int launch_tail_rows(const float* u, RowGeom g, int uc, const float* resid, int B, int P, int scale, float mean, float stdv,
                     int clip_round, float* sr, cudaStream_t st);
int launch_tail_bwd_rows(const float* dsr, int B, int P, int scale, float stdv, float* gu, RowGeom g, int uc, float* dtail, cudaStream_t st,
                         int round_tf32 = 0, float* gu_pack = nullptr);

}  // namespace pv
