// engine_tc.cu -- forward / backward sequencing of the WDSR graph on the row layouts (rows.h): the tensor-core
// engine (pv_cfg.precision = 1, tf32 tcgen05 kernels) and its fp32 CUDA-core twin on the same layouts (precision = 3).
//
// Same graph as engine.cu (reference models/modelsTF.py:15-43, 55-74, 152-164, 177-189; backward = tape.gradient,
// models/trainClass.py:131), different data layout:
//   prep (dense) -> mainConv1 -> A0 [PR] -> 12 x { expConv+ReLU -> E ; decConv -> D ; normConv + A_i -> A_i+1 } [PR]
//   -> reflect pad -> G0 [G] -> convReducer_1..3 (+ReLU) -> G1..G3 -> upscaleConv1 -> U [G] -> tail (+ 2-D skip path)
// Backward: every data-gradient kernel masks its output with the ReLU of the layer it flows into, so the stored
// tensors are dL/d(pre-activation) and the weight-gradient kernels need no mask.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cmath>
#include <cstring>

#include "engine.h"

namespace pv {
namespace {

struct Taps { int n; int off[MAX_TAPS]; int c0[MAX_TAPS]; int chunk[MAX_TAPS]; };

// PR layout of a 22 x 22 x T trunk activation (T = num_low_res_imgs: 7, 9 or 13 on this engine)
RowGeom pr_geom(int T = 9) {
    RowGeom g;
    g.lead = 640; g.pstride = (long long)(T + 1) * 529; g.plane = 529; g.pw = 23; g.t0 = 0; g.nt = T; g.nh = 22; g.nw = 22;
    g.row0 = 0; g.nrows = T * 529;
    return g;
}
// G layout of a tensor whose valid extent is nt planes of nh x nw (inside planes of 24 x 24, two leading zero planes)
RowGeom g_dims(int nt, int nh, int nw) {
    RowGeom g;
    g.lead = 1280; g.pstride = (long long)(nt + 2) * 576; g.plane = 576; g.pw = 24; g.t0 = 2;
    g.nt = nt; g.nh = nh; g.nw = nw; g.row0 = 2 * 576; g.nrows = nt * 576;
    return g;
}
// ... after k valid 3x3x3 convolutions of the reflect-padded 24x24x9 block output (the T = 9 tail; used by the self-test)
RowGeom g_geom(int k) { return g_dims(9 - 2 * k, 24 - 2 * k, 24 - 2 * k); }

// The reducer tail on G buffers (modelsTF.py:62-69: ConvReduceAndUpscale for T = 9, v2 for T = 7, v3 for T = 13): reducer i
// reads "Gi<i>" -- the reflect-padded (by pad_i on H and W) copy of its predecessor's output, or that output itself when
// pad_i = 0 -- and writes "Go<i>"; the upscale conv reads the last "Go" and writes "U".
struct TailStep { std::string in, out; RowGeom ig, og; int pad; bool copy; };
std::vector<TailStep> tail_plan(const pv_model* m) {
    std::vector<TailStep> v;
    int nt = m->T, nh = m->S, nw = m->S;
    for (int i = 0; i < m->nred; ++i) {
        TailStep s;
        s.pad = m->red_pad[3 * i];
        s.copy = i == 0 || s.pad > 0;                       // the first reducer always needs the PR -> G re-layout
        nh += 2 * s.pad; nw += 2 * s.pad;
        s.in = s.copy ? "Gi" + std::to_string(i + 1) : v.back().out;
        s.ig = g_dims(nt, nh, nw);
        nt -= 2; nh -= 2; nw -= 2;
        s.out = "Go" + std::to_string(i + 1);
        s.og = g_dims(nt, nh, nw);
        v.push_back(s);
    }
    return v;
}
bool tail_plan_supported(const pv_model* m) {
    if (m->nred < 1) return false;
    int nt = m->T, nh = m->S;
    for (int i = 0; i < m->nred; ++i) {
        const int ph = m->red_pad[3 * i], pw = m->red_pad[3 * i + 1], pt = m->red_pad[3 * i + 2];
        if (ph != pw || ph > 1 || pt != 0) return false;   // T = 19 pads T and uses a 5x5x5 reducer: dense engine only
        nh += 2 * ph;
        if (nh > 24) return false;
        nt -= 2; nh -= 2;
    }
    return nt == 3 && nh == m->P + 2;
}
size_t rows_per(const RowGeom& g, int C) { return (size_t)g.pstride * C; }
size_t rows_extra(const RowGeom& g, int C) { return (size_t)(g.lead + ROW_TAIL) * C; }

// taps of a 3x3x3 convolution as row offsets; sign = +1 forward, -1 data gradient (listed in ascending offset order)
Taps conv3_taps(int plane, int pw, bool centred, int sign) {
    Taps t; t.n = 27;
    for (int i = 0; i < 27; ++i) {
        const int tau = sign > 0 ? i : 26 - i;
        const int dt = tau / 9 - (centred ? 1 : 0), dh = (tau / 3) % 3 - (centred ? 1 : 0), dw = tau % 3 - (centred ? 1 : 0);
        t.off[i] = sign * (dt * plane + dh * pw + dw);
        t.c0[i] = 0;
        t.chunk[i] = tau;
    }
    return t;
}
// a pointwise layer whose input rows are `kin` channels wide: K-chunks of 32 as "taps"
Taps chunk_taps(int kin) {
    Taps t; t.n = kin / 32;
    for (int i = 0; i < t.n; ++i) { t.off[i] = 0; t.c0[i] = 32 * i; t.chunk[i] = i; }
    return t;
}

// forward convolution of layer L:  y = act(conv(x) + bias) (+ residual), masked to the valid extent of `og`
// x_f16: x holds fp16 pair rows (hi halves used): the 3x3x3 lattice runs as conv3_tc.cu MODE 3 against the fp16 weight copies
int conv_rows(pv_model* m, const Layer& L, const Taps& tp, const float* x, int xc, const RowGeom& ig, float* y, const RowGeom& og,
              const float* residual, int B, const char* tag, cudaStream_t st, bool round_out = true, bool x_f16 = false) {
    RowConvP p;
    memset(&p, 0, sizeof p);
    p.x = x; p.xc = xc; p.y = y; p.n = L.cout_s; p.B = B;
    p.in_lead = ig.lead; p.in_pstride = ig.pstride; p.og = og;
    p.ntap = tp.n; p.kc = 32;
    const int Kflat = L.taps() * L.cin_s;
    for (int i = 0; i < tp.n; ++i) {
        p.off[i] = tp.off[i]; p.c0[i] = tp.c0[i];
        if (m->use_tc) { p.wr0[i] = 0; p.wc0[i] = 32 * tp.chunk[i]; }       // weffT [cout_s][Kflat]
        else { p.wr0[i] = 32 * tp.chunk[i]; p.wc0[i] = 0; }                  // weff  [Kflat][cout_s]
    }
    if (m->use_tc) { p.w = m->weffT + L.weff_off; p.w_rows = L.cout_s; p.w_cols = Kflat; p.w_kmajor = 1; }
    else { p.w = m->weff + L.weff_off; p.w_rows = Kflat; p.w_cols = L.cout_s; p.w_kmajor = 0; }
    p.bias = m->bias_s + L.bias_s_off; p.residual = residual; p.relumask = nullptr; p.relu = L.relu;
    p.round_tf32 = (m->use_tc && round_out) ? 1 : 0;      // outputs that only feed MMAs are stored round-to-nearest tf32
    p.flops = 2.0 * B * L.Ho * L.Wo * L.To * L.taps() * L.cin * L.cout;
    p.tag = tag;
    if (x_f16) {
        if (!m->use_tc || !m->weffT_pack || !rowconv3_tc_supported(p)) return set_error(PV_ERR_BAD_ARG, "conv_rows: fp16 rows need the 3x3x3 tensor-core kernel");
        p.w = m->weffT_pack + L.weff_off; p.f16_pack = 3;
    }
    return m->use_tc ? launch_rowconv_tc(p, st) : launch_rowconv_simt(p, st);
}

// data gradient of layer L: gx = conv^T(gz) (+ residual), multiplied by (relumask > 0), masked to the valid extent of `xg`
// gz_pack (precision 4, nullable): gz as bf16 pair rows -> the split-weight product in one launch (conv3_tc.cu MODE 2)
int dgrad_rows(pv_model* m, const Layer& L, const Taps& tp /* negated offsets */, int nchunk_out /* cout_s / 32 */,
               const float* gz, const RowGeom& zg, float* gx, const RowGeom& xg, const float* residual, const float* relumask,
               int B, const char* tag, cudaStream_t st, const float* gz_pack = nullptr, float* gx_pack = nullptr) {
    RowConvP p;
    memset(&p, 0, sizeof p);
    p.x = gz; p.xc = L.cout_s; p.y = gx; p.n = L.cin_s; p.B = B;
    p.in_lead = zg.lead; p.in_pstride = zg.pstride; p.og = xg;
    p.kc = 32;
    const int Kflat = L.taps() * L.cin_s;
    int n = 0;
    for (int i = 0; i < tp.n; ++i)
        for (int j = 0; j < nchunk_out; ++j, ++n) {
            if (n >= MAX_TAPS) return set_error(PV_ERR_BAD_ARG, "dgrad_rows: too many taps");
            p.off[n] = tp.off[i]; p.c0[n] = 32 * j;
            const int tau = tp.chunk[i];
            if (m->use_tc) { p.wr0[n] = tau * L.cin_s; p.wc0[n] = 32 * j; }      // weff [Kflat][cout_s]: rows ci, K = co contiguous
            else { p.wr0[n] = 32 * j; p.wc0[n] = tau * L.cin_s; }                 // weffT [cout_s][Kflat]: rows co (K), ci contiguous
        }
    p.ntap = n;
    if (m->use_tc) { p.w = m->weff + L.weff_off; p.w_rows = Kflat; p.w_cols = L.cout_s; p.w_kmajor = 1; }
    else { p.w = m->weffT + L.weff_off; p.w_rows = L.cout_s; p.w_cols = Kflat; p.w_kmajor = 0; }
    p.bias = nullptr; p.residual = residual; p.relumask = relumask; p.relu = 0;
    p.round_tf32 = m->use_tc ? 1 : 0;
    p.flops = 2.0 * B * L.Ho * L.Wo * L.To * L.taps() * L.cin * L.cout;
    p.tag = tag;
    if (m->x3 && gz_pack && rowconv3_tc_supported(p)) {
        p.x = gz_pack; p.w = m->weff_pack + L.weff_off; p.f16_pack = 2; p.y_pack = gx_pack;      // gx_pack: gx as bf16 pair rows for the next one
        return launch_rowconv_tc(p, st);
    }
    // precision 4 runs every 3x3x3 data gradient with split weights (the weight rounding is the one systematic error of the data-gradient
    // chain, profiles/r02_tf32_numerics_study.md); the pointwise ones live in the fused block kernels
    if (m->x3) return set_error(PV_ERR_BAD_ARG, "dgrad_rows: the error-compensated engine needs bf16 pair rows of the gradient and a 3x3x3 lattice");
    return m->use_tc ? launch_rowconv_tc(p, st) : launch_rowconv_simt(p, st);
}

// Error-compensated forward convolution (precision 4), x w ~= x_hi w_hi + (x_lo w_hi + x_hi w_lo), in ONE launch of the conv3 kernel over
// packed fp16 pair rows (rows.h): the main product and the two corrections accumulate in separate TMEM accumulators (conv3_tc.cu, F16),
//   v = act(x_hi w_hi + 2^-12 (2^12 x_lo w_hi + x_hi 2^12 w_lo) + bias (+ res_hi + res_lo))  ->  y_hi = tf32(v), y_lo = v - y_hi, y_pack.
// (Round 2's earlier versions: three tf32 passes, then a packed correction pass + a tf32 main pass chained through an fp32 partial buffer.)
// y_lo / y_pack nullable; y_lo == y_pack == nullptr stores v itself (un-rounded fp32: the upscale conv, whose output feeds the CUDA-core tail).
int conv_rows_x3(pv_model* m, const Layer& L, const Taps& tp, const float* x_pack, const RowGeom& ig, float* y_hi, float* y_lo,
                 float* y_pack, const RowGeom& og, const float* res_hi, const float* res_lo, int B, const char* tag, cudaStream_t st) {
    RowConvP p;
    memset(&p, 0, sizeof p);
    p.xc = 32; p.n = L.cout_s; p.B = B;
    p.in_lead = ig.lead; p.in_pstride = ig.pstride; p.og = og;
    p.ntap = tp.n; p.kc = 32;
    const int Kflat = L.taps() * L.cin_s;
    for (int i = 0; i < tp.n; ++i) { p.off[i] = tp.off[i]; p.c0[i] = tp.c0[i]; p.wr0[i] = 0; p.wc0[i] = 32 * tp.chunk[i]; }
    p.w_rows = L.cout_s; p.w_cols = Kflat; p.w_kmajor = 1;
    p.tag = tag;
    p.x = x_pack; p.w = m->weffT_pack + L.weff_off; p.f16_pack = 1; p.bias = m->bias_s + L.bias_s_off;
    p.residual = res_hi; p.residual2 = res_hi ? res_lo : nullptr;
    p.relu = L.relu; p.y = y_hi; p.y_lo = y_lo; p.y_pack = y_pack;
    p.flops = 2.0 * B * L.Ho * L.Wo * L.To * L.taps() * L.cin * L.cout;
    return launch_rowconv_tc(p, st);
}

// weight gradient of layer L:  dweff[Kflat][cout_s] += x^T gz,  dbias += column sums of gz
int wgrad_rows(pv_trainer* t, const Layer& L, const Taps& tp /* forward offsets */, const float* x, int xc, const RowGeom& ig,
               const float* gz, const RowGeom& og, int B, const char* tag, cudaStream_t st) {
    pv_model* m = t->m;
    RowWgradP p;
    memset(&p, 0, sizeof p);
    p.x = x; p.xc = xc; p.gz = gz; p.n = L.cout_s; p.B = B;
    p.dw = t->dweff + L.weff_off; p.dw_cols = L.cout_s; p.db = t->dbias_s + L.bias_s_off;
    p.in_lead = ig.lead; p.in_pstride = ig.pstride; p.og = og;
    p.ntap = tp.n; p.kc = 32;
    for (int i = 0; i < tp.n; ++i) { p.off[i] = tp.off[i]; p.c0[i] = tp.c0[i]; p.dwr0[i] = 32 * tp.chunk[i]; p.dwc0[i] = 0; }
    p.flops = 2.0 * B * L.Ho * L.Wo * L.To * L.taps() * L.cin * L.cout;
    p.tag = tag;
    if (m->use_tc) return launch_rowwgrad_tc(p, st, t->wg_partials, t->wg_partial_floats, &t->rq);
    return launch_rowwgrad_simt(p, st);
}

}  // namespace

// ------------------------------------------------------------------------------------------ self-test
// Runs each tensor-core kernel configuration of the graph against the CUDA-core kernel on the same random row
// buffers (values on a coarse dyadic grid, so tf32 products are exact and both paths must agree to fp32 rounding).
// `ig` (nullable): geometry of the input rows; rows outside its valid extent are zeroed, which is the engine's layout
// invariant (rows.h) and what the N = 96 kernel's halo-free first lane relies on.
static int selftest_one(const char* name, RowConvP p, size_t in_floats, size_t out_floats, size_t w_floats, std::string& rep,
                        const RowGeom* ig = nullptr) {
    float *x = nullptr, *w = nullptr, *wt = nullptr, *b = nullptr, *res = nullptr, *msk = nullptr, *y0 = nullptr, *y1 = nullptr;
    PV_CUDA(cudaMalloc(&x, in_floats * 4)); PV_CUDA(cudaMalloc(&w, w_floats * 4)); PV_CUDA(cudaMalloc(&wt, w_floats * 4));
    PV_CUDA(cudaMalloc(&b, 256 * 4)); PV_CUDA(cudaMalloc(&res, out_floats * 4)); PV_CUDA(cudaMalloc(&msk, out_floats * 4));
    PV_CUDA(cudaMalloc(&y0, out_floats * 4)); PV_CUDA(cudaMalloc(&y1, out_floats * 4));
    std::vector<float> h(std::max(std::max(in_floats, out_floats), w_floats));
    unsigned s = 12345u;
    auto fill = [&](float* d, size_t n, int range, float scale) {
        for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h[i] = (float)((int)((s >> 16) % (2 * range + 1)) - range) * scale; }
        return cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    };
    if (ig) {
        for (size_t i = 0; i < in_floats; ++i) { s = s * 1664525u + 1013904223u; h[i] = (float)((int)((s >> 16) % 9) - 4) * 0.125f; }
        for (long long r = 0; r < (long long)(in_floats / p.xc); ++r) {
            const long long q = r - ig->lead;
            const bool ok = q >= 0 && q < (long long)p.B * ig->pstride && row_valid(*ig, (int)(q % ig->pstride));
            if (!ok) for (int c = 0; c < p.xc; ++c) h[(size_t)r * p.xc + c] = 0.f;
        }
        PV_CUDA(cudaMemcpy(x, h.data(), in_floats * 4, cudaMemcpyHostToDevice));
    } else {
        PV_CUDA(fill(x, in_floats, 4, 0.125f));
    }
    PV_CUDA(fill(b, 256, 4, 0.25f));
    PV_CUDA(fill(res, out_floats, 4, 0.25f)); PV_CUDA(fill(msk, out_floats, 1, 1.0f));
    // weights: [rows = p.w_rows][cols = p.w_cols] K-major for the tensor cores, transposed copy for the CUDA cores
    PV_CUDA(fill(wt, w_floats, 3, 0.125f));
    std::vector<float> hw(w_floats);
    for (int r = 0; r < p.w_rows; ++r) for (int cc = 0; cc < p.w_cols; ++cc) hw[(size_t)cc * p.w_rows + r] = h[(size_t)r * p.w_cols + cc];
    PV_CUDA(cudaMemcpy(w, hw.data(), w_floats * 4, cudaMemcpyHostToDevice));
    PV_CUDA(cudaMemset(y0, 0, out_floats * 4)); PV_CUDA(cudaMemset(y1, 0, out_floats * 4));
    const bool use_res = p.residual != nullptr, use_msk = p.relumask != nullptr, use_bias = p.bias != nullptr;
    p.x = x; p.bias = use_bias ? b : nullptr; p.residual = use_res ? res : nullptr; p.relumask = use_msk ? msk : nullptr;
    RowConvP q = p;                       // CUDA-core twin: transposed weight matrix, swapped box coordinates
    q.w = w; q.w_kmajor = 0; q.w_rows = p.w_cols; q.w_cols = p.w_rows;
    for (int i = 0; i < p.ntap; ++i) { q.wr0[i] = p.wc0[i]; q.wc0[i] = p.wr0[i]; }
    q.y = y0; p.w = wt; p.y = y1;
    int rc = launch_rowconv_simt(q, 0);
    if (!rc) rc = launch_rowconv_tc(p, 0);
    if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = set_error(PV_ERR_CUDA, "selftest %s: %s", name, cudaGetErrorString(cudaGetLastError()));
    double worst = 0; size_t bad = 0;
    if (!rc) {
        std::vector<float> a(out_floats), c2(out_floats);
        cudaMemcpy(a.data(), y0, out_floats * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(c2.data(), y1, out_floats * 4, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < out_floats; ++i) { const double d = std::fabs((double)a[i] - c2[i]); if (!(d <= 1e-3)) ++bad; if (d > worst || d != d) worst = d; }
    }
    char line[256];
    snprintf(line, sizeof line, "%-34s %s max|tc - simt| = %.3g, mismatches %zu of %zu%s%s\n", name, (!rc && bad == 0) ? "PASS" : "FAIL",
             worst, bad, out_floats, rc ? " : " : "", rc ? last_error().c_str() : "");
    rep += line;
    cudaFree(x); cudaFree(w); cudaFree(wt); cudaFree(b); cudaFree(res); cudaFree(msk); cudaFree(y0); cudaFree(y1);
    return (!rc && bad == 0) ? 0 : 1;
}

// fused expand -> ReLU -> decay against the two CUDA-core kernels run back to back
static int selftest_resfront(std::string& rep) {
    const int B = 3;
    const RowGeom pr = pr_geom();
    const size_t rows = (size_t)(pr.lead + (long long)B * pr.pstride + ROW_TAIL);
    float *x, *we, *weT, *wd, *wdT, *be, *bd, *E, *d0, *d1;
    PV_CUDA(cudaMalloc(&x, rows * 32 * 4)); PV_CUDA(cudaMalloc(&we, 8192 * 4)); PV_CUDA(cudaMalloc(&weT, 8192 * 4));
    PV_CUDA(cudaMalloc(&wd, 8192 * 4)); PV_CUDA(cudaMalloc(&wdT, 8192 * 4)); PV_CUDA(cudaMalloc(&be, 1024)); PV_CUDA(cudaMalloc(&bd, 128));
    PV_CUDA(cudaMalloc(&E, rows * 256 * 4)); PV_CUDA(cudaMalloc(&d0, rows * 32 * 4)); PV_CUDA(cudaMalloc(&d1, rows * 32 * 4));
    std::vector<float> h(rows * 32), t(8192);
    unsigned s = 4242u;
    auto gen = [&](std::vector<float>& v, size_t n, int range, float scale) {
        for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; v[i] = (float)((int)((s >> 16) % (2 * range + 1)) - range) * scale; }
    };
    gen(h, rows * 32, 4, 0.125f); PV_CUDA(cudaMemcpy(x, h.data(), rows * 32 * 4, cudaMemcpyHostToDevice));
    gen(h, 8192, 3, 0.125f);                                       // We [32][256] (k rows, n cols)
    for (int k = 0; k < 32; ++k) for (int n = 0; n < 256; ++n) t[(size_t)n * 32 + k] = h[(size_t)k * 256 + n];
    PV_CUDA(cudaMemcpy(we, h.data(), 8192 * 4, cudaMemcpyHostToDevice)); PV_CUDA(cudaMemcpy(weT, t.data(), 8192 * 4, cudaMemcpyHostToDevice));
    gen(h, 8192, 3, 0.125f);                                       // Wd [256][32]
    for (int k = 0; k < 256; ++k) for (int n = 0; n < 32; ++n) t[(size_t)n * 256 + k] = h[(size_t)k * 32 + n];
    PV_CUDA(cudaMemcpy(wd, h.data(), 8192 * 4, cudaMemcpyHostToDevice)); PV_CUDA(cudaMemcpy(wdT, t.data(), 8192 * 4, cudaMemcpyHostToDevice));
    gen(h, 256, 4, 0.25f); PV_CUDA(cudaMemcpy(be, h.data(), 1024, cudaMemcpyHostToDevice));
    gen(h, 32, 4, 0.25f); PV_CUDA(cudaMemcpy(bd, h.data(), 128, cudaMemcpyHostToDevice));
    PV_CUDA(cudaMemset(d0, 0, rows * 32 * 4)); PV_CUDA(cudaMemset(d1, 0, rows * 32 * 4)); PV_CUDA(cudaMemset(E, 0, rows * 256 * 4));
    RowConvP p;
    memset(&p, 0, sizeof p);
    p.x = x; p.xc = 32; p.w = we; p.w_rows = 32; p.w_cols = 256; p.w_kmajor = 0; p.bias = be; p.y = E; p.n = 256; p.B = B;
    p.in_lead = pr.lead; p.in_pstride = pr.pstride; p.og = pr; p.ntap = 1; p.kc = 32; p.relu = 1;
    int rc = launch_rowconv_simt(p, 0);
    RowConvP q;
    memset(&q, 0, sizeof q);
    q.x = E; q.xc = 256; q.w = wd; q.w_rows = 256; q.w_cols = 32; q.w_kmajor = 0; q.bias = bd; q.y = d0; q.n = 32; q.B = B;
    q.in_lead = pr.lead; q.in_pstride = pr.pstride; q.og = pr; q.ntap = 8; q.kc = 32;
    for (int j = 0; j < 8; ++j) { q.c0[j] = 32 * j; q.wr0[j] = 32 * j; }
    if (!rc) rc = launch_rowconv_simt(q, 0);
    // both forward variants: inference (no ReLU bit mask) and training (bit mask written)
    uint32_t* bits = nullptr;
    PV_CUDA(cudaMalloc(&bits, rows * 32));
    size_t bad = 0;
    for (int variant = 0; variant < 2; ++variant) {
        PV_CUDA(cudaMemset(d1, 0, rows * 32 * 4));
        if (!rc) rc = launch_resfront_fwd_tc(x, weT, wdT, be, bd, d1, variant ? bits : nullptr, pr, B, 0, 0.0, 0);
        if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = set_error(PV_ERR_CUDA, "selftest resfront: %s", cudaGetErrorString(cudaGetLastError()));
        double worst = 0; size_t badv = 0;
        if (!rc) {
            std::vector<float> a(rows * 32), c2(rows * 32);
            cudaMemcpy(a.data(), d0, rows * 32 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(c2.data(), d1, rows * 32 * 4, cudaMemcpyDeviceToHost);
            for (size_t i = 0; i < rows * 32; ++i) { const double dd = std::fabs((double)a[i] - c2[i]); if (!(dd <= 1e-3)) ++badv; if (dd > worst || dd != dd) worst = dd; }
        }
        char line[256];
        snprintf(line, sizeof line, "%-34s %s max|tc - simt| = %.3g, mismatches %zu of %zu%s%s\n",
                 variant ? "fused exp->relu->dec fwd (train)" : "fused exp->relu->dec fwd (infer)", (!rc && badv == 0) ? "PASS" : "FAIL",
                 worst, badv, rows * 32, rc ? " : " : "", rc ? last_error().c_str() : "");
        rep += line;
        bad += badv;
    }
    cudaFree(bits);
    cudaFree(x); cudaFree(we); cudaFree(weT); cudaFree(wd); cudaFree(wdT); cudaFree(be); cudaFree(bd); cudaFree(E); cudaFree(d0); cudaFree(d1);
    return (!rc && bad == 0) ? 0 : 1;
}

// fused backward of the expand/decay chain (data + weight kernels) against the CUDA-core kernels chained through HBM
static int selftest_resback(std::string& rep) {
    const int B = 3;
    const RowGeom pr = pr_geom();
    const size_t rows = (size_t)(pr.lead + (long long)B * pr.pstride + ROW_TAIL);
    const size_t part_floats = (size_t)148 * (9 * 4096 + 1024);
    float *x, *gd, *G, *M, *we, *weT, *wd, *wdT, *be, *E, *gZ, *ga0, *ga1, *dwd0, *dwd1, *dwe0, *dwe1, *db0, *db1, *part;
    PV_CUDA(cudaMalloc(&x, rows * 128)); PV_CUDA(cudaMalloc(&gd, rows * 128)); PV_CUDA(cudaMalloc(&G, rows * 128)); PV_CUDA(cudaMalloc(&M, rows * 128));
    PV_CUDA(cudaMalloc(&we, 32768)); PV_CUDA(cudaMalloc(&weT, 32768)); PV_CUDA(cudaMalloc(&wd, 32768)); PV_CUDA(cudaMalloc(&wdT, 32768));
    PV_CUDA(cudaMalloc(&be, 1024)); PV_CUDA(cudaMalloc(&E, rows * 1024)); PV_CUDA(cudaMalloc(&gZ, rows * 1024));
    PV_CUDA(cudaMalloc(&ga0, rows * 128)); PV_CUDA(cudaMalloc(&ga1, rows * 128));
    PV_CUDA(cudaMalloc(&dwd0, 32768)); PV_CUDA(cudaMalloc(&dwd1, 32768)); PV_CUDA(cudaMalloc(&dwe0, 32768)); PV_CUDA(cudaMalloc(&dwe1, 32768));
    PV_CUDA(cudaMalloc(&db0, 2048)); PV_CUDA(cudaMalloc(&db1, 2048)); PV_CUDA(cudaMalloc(&part, part_floats * 4));
    std::vector<float> h(rows * 32), t(8192);
    unsigned s = 99u;
    auto gen = [&](std::vector<float>& v, size_t n, int range, float scale) {
        for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; v[i] = (float)((int)((s >> 16) % (2 * range + 1)) - range) * scale; }
    };
    auto mask_rows = [&](std::vector<float>& v) {
        for (size_t r = 0; r < rows; ++r) {
            const long long q = (long long)r - pr.lead;
            const bool ok = q >= 0 && q < (long long)B * pr.pstride && row_valid(pr, (int)(q % pr.pstride));
            if (!ok) for (int c = 0; c < 32; ++c) v[r * 32 + c] = 0.f;
        }
    };
    gen(h, rows * 32, 4, 0.125f); mask_rows(h); PV_CUDA(cudaMemcpy(x, h.data(), rows * 128, cudaMemcpyHostToDevice));
    gen(h, rows * 32, 4, 0.125f); mask_rows(h); PV_CUDA(cudaMemcpy(gd, h.data(), rows * 128, cudaMemcpyHostToDevice));
    gen(h, rows * 32, 4, 0.25f); PV_CUDA(cudaMemcpy(G, h.data(), rows * 128, cudaMemcpyHostToDevice));
    gen(h, rows * 32, 1, 1.0f); PV_CUDA(cudaMemcpy(M, h.data(), rows * 128, cudaMemcpyHostToDevice));
    gen(h, 8192, 3, 0.125f);                                       // We [32 ci][256 ch]
    for (int k = 0; k < 32; ++k) for (int n = 0; n < 256; ++n) t[(size_t)n * 32 + k] = h[(size_t)k * 256 + n];
    PV_CUDA(cudaMemcpy(we, h.data(), 32768, cudaMemcpyHostToDevice)); PV_CUDA(cudaMemcpy(weT, t.data(), 32768, cudaMemcpyHostToDevice));
    gen(h, 8192, 3, 0.125f);                                       // Wd [256 ch][32 co]
    for (int k = 0; k < 256; ++k) for (int n = 0; n < 32; ++n) t[(size_t)n * 256 + k] = h[(size_t)k * 32 + n];
    PV_CUDA(cudaMemcpy(wd, h.data(), 32768, cudaMemcpyHostToDevice)); PV_CUDA(cudaMemcpy(wdT, t.data(), 32768, cudaMemcpyHostToDevice));
    gen(h, 256, 4, 0.25f); PV_CUDA(cudaMemcpy(be, h.data(), 1024, cudaMemcpyHostToDevice));
    for (float* p : {E, gZ}) PV_CUDA(cudaMemset(p, 0, rows * 1024));
    for (float* p : {ga0, ga1}) PV_CUDA(cudaMemset(p, 0, rows * 128));
    for (float* p : {dwd0, dwd1, dwe0, dwe1}) PV_CUDA(cudaMemset(p, 0, 32768));
    PV_CUDA(cudaMemset(db0, 0, 2048)); PV_CUDA(cudaMemset(db1, 0, 2048));
    auto conv = [&](const float* in, int xc, const float* w, int wr, int wc, float* out, int n) {
        RowConvP p;
        memset(&p, 0, sizeof p);
        p.x = in; p.xc = xc; p.w = w; p.w_rows = wr; p.w_cols = wc; p.w_kmajor = 0; p.y = out; p.n = n; p.B = B;
        p.in_lead = pr.lead; p.in_pstride = pr.pstride; p.og = pr; p.ntap = xc / 32; p.kc = 32;
        for (int j = 0; j < p.ntap; ++j) { p.c0[j] = 32 * j; p.wr0[j] = 32 * j; }
        return p;
    };
    RowConvP pe = conv(x, 32, we, 32, 256, E, 256); pe.bias = be; pe.relu = 1;
    int rc = launch_rowconv_simt(pe, 0);
    RowConvP pz = conv(gd, 32, wdT, 32, 256, gZ, 256); pz.relumask = E;              // gZ = (gD Wd^T) .* (E > 0)
    if (!rc) rc = launch_rowconv_simt(pz, 0);
    RowConvP pa = conv(gZ, 256, weT, 256, 32, ga0, 32); pa.residual = G; pa.relumask = M;
    if (!rc) rc = launch_rowconv_simt(pa, 0);
    auto wg = [&](const float* in, int xc, const float* gz, int n, float* dw, float* db) {
        RowWgradP p;
        memset(&p, 0, sizeof p);
        p.x = in; p.xc = xc; p.gz = gz; p.n = n; p.dw = dw; p.dw_cols = n; p.db = db; p.B = B;
        p.in_lead = pr.lead; p.in_pstride = pr.pstride; p.og = pr; p.ntap = xc / 32; p.kc = 32;
        for (int j = 0; j < p.ntap; ++j) { p.c0[j] = 32 * j; p.dwr0[j] = 32 * j; }
        return p;
    };
    if (!rc) rc = launch_rowwgrad_simt(wg(E, 256, gd, 32, dwd0, db0 + 256), 0);
    if (!rc) rc = launch_rowwgrad_simt(wg(x, 32, gZ, 256, dwe0, db0), 0);
    // the backward-data kernel consumes the ReLU bits the fused forward emits
    float *dtmp = nullptr, *bd0 = nullptr; uint32_t* bits = nullptr;
    PV_CUDA(cudaMalloc(&dtmp, rows * 128)); PV_CUDA(cudaMalloc(&bd0, 128)); PV_CUDA(cudaMalloc(&bits, rows * 32));
    PV_CUDA(cudaMemset(bd0, 0, 128)); PV_CUDA(cudaMemset(bits, 0, rows * 32));
    if (!rc) rc = launch_resfront_fwd_tc(x, weT, wdT, be, bd0, dtmp, bits, pr, B, 0, 0.0, 0);
    if (!rc) rc = launch_resfront_bwd_data_tc(gd, wd, we, bits, G, M, ga1, pr, B, 0, 0.0, 0);
    if (!rc) rc = launch_resfront_bwd_weight_tc(x, gd, weT, wd, be, dwd1, dwe1, db1, db1 + 256, pr, B, part, part_floats, 0.0, 0);
    if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = set_error(PV_ERR_CUDA, "selftest resback: %s", cudaGetErrorString(cudaGetLastError()));
    auto cmp = [&](const char* name, const float* d0, const float* d1, size_t n) {
        std::vector<float> a(n), c2(n);
        cudaMemcpy(a.data(), d0, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(c2.data(), d1, n * 4, cudaMemcpyDeviceToHost);
        double worst = 0, ref = 0; size_t bad = 0;
        for (size_t i = 0; i < n; ++i) ref = std::max(ref, (double)std::fabs(a[i]));
        for (size_t i = 0; i < n; ++i) { const double dd = std::fabs((double)a[i] - c2[i]); if (!(dd <= 1e-5 * ref + 1e-6)) ++bad; if (dd > worst || dd != dd) worst = dd; }
        char line[256];
        snprintf(line, sizeof line, "%-34s %s max|tc - simt| = %.3g (max |ref| %.3g), mismatches %zu of %zu\n", name, bad == 0 ? "PASS" : "FAIL", worst, ref, bad, n);
        rep += line;
        return bad == 0 ? 0 : 1;
    };
    int fails = 0;
    if (rc) { rep += std::string("fused exp/dec backward              FAIL : ") + last_error() + "\n"; fails = 1; }
    else {
        fails += cmp("fused bwd: gA (+skip, relu mask)", ga0, ga1, rows * 32);
        fails += cmp("fused bwd: dW decConv", dwd0, dwd1, 8192);
        fails += cmp("fused bwd: dW expConv", dwe0, dwe1, 8192);
        fails += cmp("fused bwd: db expConv | db decConv", db0, db1, 288);
    }
    if (!rc) {
        // split-weight variant of the backward-data kernel (precision 4): with W_lo := W_hi both products double, so the result
        // must be exactly four times the single-weight one (all values dyadic; the tf32 rounding of H commutes with x2)
        PV_CUDA(cudaMemset(ga0, 0, rows * 128)); PV_CUDA(cudaMemset(ga1, 0, rows * 128));
        rc = launch_resfront_bwd_data_tc(gd, wd, we, bits, nullptr, nullptr, ga0, pr, B, 0, 0.0, 0);
        if (!rc) rc = launch_resfront_bwd_data_tc(gd, wd, we, bits, nullptr, nullptr, ga1, pr, B, 0, 0.0, 0, wd, we);
        if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = set_error(PV_ERR_CUDA, "selftest resback split: %s", cudaGetErrorString(cudaGetLastError()));
        if (rc) { rep += std::string("fused bwd: split weights            FAIL : ") + last_error() + "\n"; fails += 1; }
        else {
            std::vector<float> a(rows * 32), c2(rows * 32);
            cudaMemcpy(a.data(), ga0, rows * 128, cudaMemcpyDeviceToHost); cudaMemcpy(c2.data(), ga1, rows * 128, cudaMemcpyDeviceToHost);
            size_t bad = 0; double ref = 0;
            for (size_t i = 0; i < rows * 32; ++i) { ref = std::max(ref, (double)std::fabs(a[i])); if (4.0f * a[i] != c2[i]) ++bad; }
            char line[256];
            snprintf(line, sizeof line, "%-34s %s gA(W, W) == 4 gA(W) (max |ref| %.3g), mismatches %zu of %zu\n", "fused bwd: split weights (W+W)",
                     (bad == 0 && ref > 0) ? "PASS" : "FAIL", ref, bad, rows * 32);
            rep += line;
            fails += (bad == 0 && ref > 0) ? 0 : 1;
        }
    }
    for (float* p : {x, gd, G, M, we, weT, wd, wdT, be, E, gZ, ga0, ga1, dwd0, dwd1, dwe0, dwe1, db0, db1, part, dtmp, bd0}) cudaFree(p);
    cudaFree(bits);
    return fails;
}

// ---- error-compensated kernels (precision 4) against the fp32 CUDA-core kernels on NON-dyadic random data: the three-MMA
// products must agree with fp32 arithmetic to ~1e-6 (a missing lo term would show up at ~1e-3)
static float host_tf32(float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    u = (u + 0x1000u) & 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}
struct SplitBuf {
    float *full = nullptr, *hi = nullptr, *lo = nullptr, *pack = nullptr;
    // pack_mode 1: activation rows [fp16(hi) | fp16(PACK_SCALE lo)], 2: weight rows [fp16(PACK_SCALE lo) | fp16(hi)] per 32 values (rows.h),
    // 3: bf16 pair rows [bf16(v) | bf16(v - bf16(v))] (gradients and the data gradient's weights)
    int upload(const std::vector<float>& h, int pack_mode = 0) {
        const size_t n = h.size();
        std::vector<float> a(n), b(n);
        for (size_t i = 0; i < n; ++i) { a[i] = host_tf32(h[i]); b[i] = h[i] - a[i]; }
        if (pack_mode == 3) {
            std::vector<__nv_bfloat16> pk(2 * n);
            for (size_t r = 0; r < n / 32; ++r)
                for (int c = 0; c < 32; ++c) {
                    const __nv_bfloat16 hh = __float2bfloat16_rn(h[r * 32 + c]);
                    pk[r * 64 + c] = hh;
                    pk[r * 64 + 32 + c] = __float2bfloat16_rn(h[r * 32 + c] - __bfloat162float(hh));
                }
            PV_CUDA(cudaMalloc(&pack, n * 4));
            PV_CUDA(cudaMemcpy(pack, pk.data(), n * 4, cudaMemcpyHostToDevice));
        } else if (pack_mode) {
            std::vector<__half> pk(2 * n);
            for (size_t r = 0; r < n / 32; ++r)
                for (int c = 0; c < 32; ++c) {
                    const __half hh = __float2half_rn(a[r * 32 + c]);                  // lo' = v - fp16(hi): hi16 + lo' = v exactly
                    const __half ll = __float2half_rn((h[r * 32 + c] - __half2float(hh)) * PACK_SCALE);
                    pk[r * 64 + (pack_mode == 1 ? c : 32 + c)] = hh;
                    pk[r * 64 + (pack_mode == 1 ? 32 + c : c)] = ll;
                }
            PV_CUDA(cudaMalloc(&pack, n * 4));
            PV_CUDA(cudaMemcpy(pack, pk.data(), n * 4, cudaMemcpyHostToDevice));
        }
        PV_CUDA(cudaMalloc(&full, n * 4)); PV_CUDA(cudaMalloc(&hi, n * 4)); PV_CUDA(cudaMalloc(&lo, n * 4));
        PV_CUDA(cudaMemcpy(full, h.data(), n * 4, cudaMemcpyHostToDevice));
        PV_CUDA(cudaMemcpy(hi, a.data(), n * 4, cudaMemcpyHostToDevice));
        PV_CUDA(cudaMemcpy(lo, b.data(), n * 4, cudaMemcpyHostToDevice));
        return 0;
    }
    void release() { cudaFree(full); cudaFree(hi); cudaFree(lo); cudaFree(pack); }
};

static int selftest_x3(std::string& rep) {
    const int B = 3;
    const RowGeom pr = pr_geom();
    const size_t rows = (size_t)(pr.lead + (long long)B * pr.pstride + ROW_TAIL);
    unsigned s = 20261017u;
    auto rnd = [&](float scale) { s = s * 1664525u + 1013904223u; return ((float)(s >> 8) / 16777216.0f - 0.5f) * 2.0f * scale; };
    auto rows_rand = [&](std::vector<float>& v, float scale) {
        v.resize(rows * 32);
        for (size_t r = 0; r < rows; ++r) {
            const long long q = (long long)r - pr.lead;
            const bool ok = q >= 0 && q < (long long)B * pr.pstride && row_valid(pr, (int)(q % pr.pstride));
            for (int c = 0; c < 32; ++c) v[r * 32 + c] = ok ? rnd(scale) : 0.f;
        }
    };
    auto compare = [&](const char* name, const float* ref_d, const float* hi_d, const float* lo_d, size_t n, double tol) {
        std::vector<float> a(n), h(n), l(n);
        cudaMemcpy(a.data(), ref_d, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h.data(), hi_d, n * 4, cudaMemcpyDeviceToHost);
        if (lo_d) cudaMemcpy(l.data(), lo_d, n * 4, cudaMemcpyDeviceToHost); else std::fill(l.begin(), l.end(), 0.f);
        double ref = 0, worst = 0; size_t bad = 0, nontf32 = 0;
        for (size_t i = 0; i < n; ++i) ref = std::max(ref, (double)std::fabs(a[i]));
        for (size_t i = 0; i < n; ++i) {
            const double d = std::fabs((double)h[i] + (double)l[i] - (double)a[i]);
            if (!(d <= tol * ref)) ++bad;
            if (d > worst || d != d) worst = d;
            if (lo_d && host_tf32(h[i]) != h[i]) ++nontf32;          // the hi half must be tf32-exact
        }
        char line[256];
        snprintf(line, sizeof line, "%-34s %s max|x3 - fp32| = %.3g (%.2g of max |ref| %.3g), over tolerance %zu, hi not tf32 %zu of %zu\n", name,
                 (bad == 0 && nontf32 == 0 && ref > 0) ? "PASS" : "FAIL", worst, ref > 0 ? worst / ref : 0.0, ref, bad, nontf32, n);
        rep += line;
        return (bad == 0 && nontf32 == 0 && ref > 0) ? 0 : 1;
    };
    int fails = 0, rc = 0;
    std::vector<float> h;
    SplitBuf X, RES, W3, WE, WD;
    rows_rand(h, 1.0f); PV_TRY(X.upload(h, 1));
    rows_rand(h, 1.0f); PV_TRY(RES.upload(h));
    // ---------------- conv3 'same' forward, bias + skip connection: weights [co = 32][K = 27 * 32] K-major for the tensor cores
    {
        std::vector<float> wk(32 * 864), wt(864 * 32), bb(32);
        for (auto& v : wk) v = rnd(0.08f);
        for (int co = 0; co < 32; ++co) for (int k = 0; k < 864; ++k) wt[(size_t)k * 32 + co] = wk[(size_t)co * 864 + k];
        for (auto& v : bb) v = rnd(0.5f);
        PV_TRY(W3.upload(wk, 2));
        float *wt_d, *b_d, *y0, *yh, *yl;
        PV_CUDA(cudaMalloc(&wt_d, wt.size() * 4)); PV_CUDA(cudaMalloc(&b_d, 128));
        PV_CUDA(cudaMemcpy(wt_d, wt.data(), wt.size() * 4, cudaMemcpyHostToDevice)); PV_CUDA(cudaMemcpy(b_d, bb.data(), 128, cudaMemcpyHostToDevice));
        for (float** p : {&y0, &yh, &yl}) { PV_CUDA(cudaMalloc(p, rows * 128)); PV_CUDA(cudaMemset(*p, 0, rows * 128)); }
        const Taps tp = conv3_taps(pr.plane, pr.pw, true, +1);
        RowConvP q;
        memset(&q, 0, sizeof q);
        q.x = X.full; q.xc = 32; q.w = wt_d; q.w_rows = 864; q.w_cols = 32; q.w_kmajor = 0; q.bias = b_d; q.residual = RES.full; q.y = y0; q.n = 32; q.B = B;
        q.in_lead = pr.lead; q.in_pstride = pr.pstride; q.og = pr; q.ntap = 27; q.kc = 32;
        for (int i = 0; i < 27; ++i) { q.off[i] = tp.off[i]; q.wr0[i] = 32 * i; q.wc0[i] = 0; }
        rc = launch_rowconv_simt(q, 0);
        pv_model fm;                                   // just enough of a model for conv_rows_x3
        fm.weffT = W3.hi; fm.weffT_lo = W3.lo; fm.weffT_pack = W3.pack; fm.bias_s = b_d; fm.use_tc = true; fm.x3 = true;
        Layer L;
        L.k[0] = L.k[1] = L.k[2] = 3; L.cin = L.cout = L.cin_s = L.cout_s = 32; L.relu = 0; L.weff_off = 0; L.bias_s_off = 0;
        L.Ho = L.Wo = 22; L.To = 9;
        float* ypk;
        PV_CUDA(cudaMalloc(&ypk, rows * 128)); PV_CUDA(cudaMemset(ypk, 0, rows * 128));
        if (!rc) rc = conv_rows_x3(&fm, L, tp, X.pack, pr, yh, yl, ypk, pr, RES.hi, RES.lo, B, "selftest_x3", 0);
        if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = set_error(PV_ERR_CUDA, "selftest x3 conv: %s", cudaGetErrorString(cudaGetLastError()));
        if (rc) { rep += std::string("x3 conv3 same fwd                  FAIL : ") + last_error() + "\n"; ++fails; }
        else {
            fails += compare("x3 conv3 same fwd (+bias +skip)", y0, yh, yl, rows * 32, 2e-5);
            // the packed output row must decode to the same (hi, lo) pair, to fp16's 11 bits of the lo half
            std::vector<float> hh(rows * 32), ll(rows * 32);
            std::vector<__half> pk(rows * 64);
            cudaMemcpy(hh.data(), yh, rows * 128, cudaMemcpyDeviceToHost); cudaMemcpy(ll.data(), yl, rows * 128, cudaMemcpyDeviceToHost);
            cudaMemcpy(pk.data(), ypk, rows * 128, cudaMemcpyDeviceToHost);
            size_t bad = 0;
            for (size_t r = 0; r < rows; ++r)
                for (int c = 0; c < 32; ++c) {
                    const float dh = __half2float(pk[r * 64 + c]), dl = __half2float(pk[r * 64 + 32 + c]) / PACK_SCALE;
                    // hi16 + lo' = hi + lo to fp16's 11 bits of the lo half (values below fp16's normal range, 6e-5, keep fewer bits: absolute floors)
                    const float v = hh[r * 32 + c] + ll[r * 32 + c], lo2 = v - dh;
                    if (std::fabs(dh - hh[r * 32 + c]) > 6e-8f || std::fabs(dl - lo2) > 1e-3f * std::fabs(lo2) + 2e-11f) ++bad;
                }
            char line[256];
            snprintf(line, sizeof line, "%-34s %s %zu mismatching elements of %zu\n", "x3 conv3: packed fp16 pair output", bad == 0 ? "PASS" : "FAIL", bad, rows * 32);
            rep += line;
            fails += bad == 0 ? 0 : 1;
        }
        // ---------------- the inference form (conv3_tc.cu MODE 3): hi halves of the fp16 pair rows only = the single-pass product of the tf32 values
        {
            float* wth_d;
            std::vector<float> wth(864 * 32);
            for (int co = 0; co < 32; ++co) for (int k = 0; k < 864; ++k) wth[(size_t)k * 32 + co] = host_tf32(wk[(size_t)co * 864 + k]);
            PV_CUDA(cudaMalloc(&wth_d, wth.size() * 4));
            PV_CUDA(cudaMemcpy(wth_d, wth.data(), wth.size() * 4, cudaMemcpyHostToDevice));
            RowConvP r0 = q;
            r0.x = X.hi; r0.w = wth_d; r0.residual = RES.hi;
            PV_CUDA(cudaMemset(y0, 0, rows * 128)); PV_CUDA(cudaMemset(yh, 0, rows * 128));
            rc = launch_rowconv_simt(r0, 0);
            RowConvP g = q;
            g.x = X.pack; g.w = W3.pack; g.w_rows = 32; g.w_cols = 864; g.w_kmajor = 1; g.f16_pack = 3; g.residual = RES.hi; g.y = yh; g.round_tf32 = 0;
            for (int i = 0; i < 27; ++i) { g.wr0[i] = 0; g.wc0[i] = 32 * i; }
            if (!rc) rc = launch_rowconv_tc(g, 0);
            if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = set_error(PV_ERR_CUDA, "selftest fp16-row conv: %s", cudaGetErrorString(cudaGetLastError()));
            if (rc) { rep += std::string("conv3 fp16 rows (inference)        FAIL : ") + last_error() + "\n"; ++fails; }
            else fails += compare("conv3 fp16 rows (inference)", y0, yh, nullptr, rows * 32, 2e-6);      // same tf32 values, fp32 accumulation order differs
            cudaFree(wth_d);
        }
        // ---------------- the same lattice as a split-weight data gradient in one launch: bf16 pair rows x fp16 pair weights (conv3_tc.cu MODE 2)
        {
            SplitBuf GZ;
            rows_rand(h, 1e-4f);                           // gradient-sized values: far below fp16's normal range, fine in bf16
            PV_TRY(GZ.upload(h, 3));
            q.x = GZ.full; q.bias = nullptr; q.residual = nullptr;
            PV_CUDA(cudaMemset(y0, 0, rows * 128)); PV_CUDA(cudaMemset(yh, 0, rows * 128));
            rc = launch_rowconv_simt(q, 0);
            RowConvP g = q;
            SplitBuf WB;
            PV_TRY(WB.upload(wk, 3));
            g.x = GZ.pack; g.w = WB.pack; g.w_rows = 32; g.w_cols = 864; g.w_kmajor = 1; g.f16_pack = 2; g.y = yh; g.round_tf32 = 0;
            for (int i = 0; i < 27; ++i) { g.wr0[i] = 0; g.wc0[i] = 32 * i; }
            if (!rc) rc = launch_rowconv_tc(g, 0);
            if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = set_error(PV_ERR_CUDA, "selftest x3 dgrad: %s", cudaGetErrorString(cudaGetLastError()));
            if (rc) { rep += std::string("x3 conv3 split-weight dgrad        FAIL : ") + last_error() + "\n"; ++fails; }
            else fails += compare("x3 conv3 split-weight dgrad", y0, yh, nullptr, rows * 32, 5e-5);     // 16 bits of each operand
            GZ.release(); WB.release();
        }
        for (float* p : {wt_d, b_d, y0, yh, yl, ypk}) cudaFree(p);
    }
    // ---------------- fused expand -> ReLU -> decay forward
    {
        std::vector<float> weT(256 * 32), we(32 * 256), wdT(32 * 256), wd(256 * 32), be(256), bd(32);
        for (int n = 0; n < 256; ++n) for (int k = 0; k < 32; ++k) { const float v = rnd(0.3f); weT[(size_t)n * 32 + k] = v; we[(size_t)k * 256 + n] = v; }
        for (int n = 0; n < 32; ++n) for (int k = 0; k < 256; ++k) { const float v = rnd(0.1f); wdT[(size_t)n * 256 + k] = v; wd[(size_t)k * 32 + n] = v; }
        for (auto& v : be) v = rnd(0.5f);
        for (auto& v : bd) v = rnd(0.5f);
        PV_TRY(WE.upload(weT)); PV_TRY(WD.upload(wdT, 2));
        float *we_d, *wd_d, *be_d, *bd_d, *Ebuf, *d0, *dh, *dl; uint32_t* bits;
        PV_CUDA(cudaMalloc(&we_d, 32768)); PV_CUDA(cudaMalloc(&wd_d, 32768)); PV_CUDA(cudaMalloc(&be_d, 1024)); PV_CUDA(cudaMalloc(&bd_d, 128));
        PV_CUDA(cudaMemcpy(we_d, we.data(), 32768, cudaMemcpyHostToDevice)); PV_CUDA(cudaMemcpy(wd_d, wd.data(), 32768, cudaMemcpyHostToDevice));
        PV_CUDA(cudaMemcpy(be_d, be.data(), 1024, cudaMemcpyHostToDevice)); PV_CUDA(cudaMemcpy(bd_d, bd.data(), 128, cudaMemcpyHostToDevice));
        PV_CUDA(cudaMalloc(&Ebuf, rows * 1024)); PV_CUDA(cudaMemset(Ebuf, 0, rows * 1024));
        for (float** p : {&d0, &dh, &dl}) { PV_CUDA(cudaMalloc(p, rows * 128)); PV_CUDA(cudaMemset(*p, 0, rows * 128)); }
        PV_CUDA(cudaMalloc(&bits, rows * 32)); PV_CUDA(cudaMemset(bits, 0, rows * 32));
        const int tpp = cdiv(pr.nrows, 128);
        uint32_t* bits_t;
        PV_CUDA(cudaMalloc(&bits_t, (size_t)B * tpp * 4096)); PV_CUDA(cudaMemset(bits_t, 0, (size_t)B * tpp * 4096));
        RowConvP pe;
        memset(&pe, 0, sizeof pe);
        pe.x = X.full; pe.xc = 32; pe.w = we_d; pe.w_rows = 32; pe.w_cols = 256; pe.w_kmajor = 0; pe.bias = be_d; pe.y = Ebuf; pe.n = 256; pe.B = B;
        pe.in_lead = pr.lead; pe.in_pstride = pr.pstride; pe.og = pr; pe.ntap = 1; pe.kc = 32; pe.relu = 1;
        rc = launch_rowconv_simt(pe, 0);
        RowConvP pd;
        memset(&pd, 0, sizeof pd);
        pd.x = Ebuf; pd.xc = 256; pd.w = wd_d; pd.w_rows = 256; pd.w_cols = 32; pd.w_kmajor = 0; pd.bias = bd_d; pd.y = d0; pd.n = 32; pd.B = B;
        pd.in_lead = pr.lead; pd.in_pstride = pr.pstride; pd.og = pr; pd.ntap = 8; pd.kc = 32;
        for (int j = 0; j < 8; ++j) { pd.c0[j] = 32 * j; pd.wr0[j] = 32 * j; }
        if (!rc) rc = launch_rowconv_simt(pd, 0);
        for (int variant = 0; variant < 2 && !rc; ++variant) {
            PV_CUDA(cudaMemset(dh, 0, rows * 128)); PV_CUDA(cudaMemset(dl, 0, rows * 128));
            rc = launch_resfront_fwd_x3_tc(X.hi, X.lo, WE.hi, WE.lo, WD.pack, be_d, bd_d, dh, dl, variant ? bits : nullptr, variant ? bits_t : nullptr, pr, B, 0.0, 0, variant);
            if (!rc && variant) {       // the training variant writes packed fp16 pair rows: decode the lo half back to fp32 for the comparison
                if (cudaDeviceSynchronize() != cudaSuccess) rc = set_error(PV_ERR_CUDA, "selftest x3 resfront: %s", cudaGetErrorString(cudaGetLastError()));
                std::vector<__half> pk(rows * 64);
                std::vector<float> ll(rows * 32);
                cudaMemcpy(pk.data(), dl, rows * 128, cudaMemcpyDeviceToHost);
                std::vector<float> hv(rows * 32);
                cudaMemcpy(hv.data(), dh, rows * 128, cudaMemcpyDeviceToHost);
                for (size_t r = 0; r < rows; ++r)                                      // v = hi16 + lo'; the comparison wants v - hi
                    for (int c = 0; c < 32; ++c)
                        ll[r * 32 + c] = (__half2float(pk[r * 64 + c]) - hv[r * 32 + c]) + __half2float(pk[r * 64 + 32 + c]) / PACK_SCALE;
                cudaMemcpy(dl, ll.data(), rows * 128, cudaMemcpyHostToDevice);
            }
            if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = set_error(PV_ERR_CUDA, "selftest x3 resfront: %s", cudaGetErrorString(cudaGetLastError()));
            if (!rc) fails += compare(variant ? "x3 fused exp->relu->dec (train)" : "x3 fused exp->relu->dec (infer)", d0, dh, dl, rows * 32, 2e-5);
        }
        if (!rc) {      // the ReLU bit mask against the fp32 expanded tensor (bit 31 - e of word c / 32 <=> E[row][c] > 0)
            std::vector<float> Eh(rows * 256);
            std::vector<uint32_t> bh(rows * 8), bt((size_t)B * tpp * 1024);
            cudaMemcpy(Eh.data(), Ebuf, rows * 1024, cudaMemcpyDeviceToHost); cudaMemcpy(bh.data(), bits, rows * 32, cudaMemcpyDeviceToHost);
            cudaMemcpy(bt.data(), bits_t, bt.size() * 4, cudaMemcpyDeviceToHost);
            size_t bad = 0, set = 0, nvalid = 0, bad_t = 0;
            for (size_t r = 0; r < rows; ++r) {
                const long long q = (long long)r - pr.lead;
                if (!(q >= 0 && q < (long long)B * pr.pstride && row_valid(pr, (int)(q % pr.pstride)))) continue;
                ++nvalid;
                for (int c2 = 0; c2 < 256; ++c2) {
                    const bool bit = (bh[r * 8 + c2 / 32] >> (31 - (c2 & 31))) & 1u, pos = Eh[r * 256 + c2] > 0.f;
                    if (bit) ++set;
                    if (bit != pos && std::fabs(Eh[r * 256 + c2]) > 1e-5f) ++bad;
                    // the transposed copy must hold the same bit: tile = (patch, row / 128), block = (row % 128) / 32, bit = row % 32
                    const int b2 = (int)(q / pr.pstride), rr = (int)(q % pr.pstride) - pr.row0;
                    const size_t w = ((size_t)(b2 * tpp + rr / 128) * 4 + (rr % 128) / 32) * 256 + c2;
                    if ((((bt[w] >> (rr % 32)) & 1u) != 0u) != bit) ++bad_t;
                }
            }
            bad += bad_t;
            char line[256];
            snprintf(line, sizeof line, "%-34s %s %zu mismatching bits of %zu (%zu set)\n", "x3 fused fwd: ReLU bit mask", (bad == 0 && set > 0) ? "PASS" : "FAIL", bad, nvalid * 256, set);
            rep += line;
            fails += (bad == 0 && set > 0) ? 0 : 1;
        }
        if (rc) { rep += std::string("x3 fused exp->relu->dec            FAIL : ") + last_error() + "\n"; ++fails; }
        for (float* p : {we_d, wd_d, be_d, bd_d, Ebuf, d0, dh, dl}) cudaFree(p);
        cudaFree(bits); cudaFree(bits_t);
    }
    X.release(); RES.release(); W3.release(); WE.release(); WD.release();
    return fails;
}

static int selftest_wgrad(const char* name, RowWgradP p, const RowGeom& ig, int B, std::string& rep) {
    const size_t in_floats = (size_t)(ig.lead + (long long)B * ig.pstride + ROW_TAIL) * p.xc;
    const size_t gz_floats = (size_t)(p.og.lead + (long long)B * p.og.pstride + ROW_TAIL) * p.n;
    const size_t dw_floats = (size_t)p.ntap * 32 * p.n;
    const size_t part_floats = (size_t)148 * (9 * 4096 + 1024);
    float *x = nullptr, *gz = nullptr, *dw0 = nullptr, *dw1 = nullptr, *db0 = nullptr, *db1 = nullptr, *part = nullptr;
    PV_CUDA(cudaMalloc(&x, in_floats * 4)); PV_CUDA(cudaMalloc(&gz, gz_floats * 4));
    PV_CUDA(cudaMalloc(&dw0, dw_floats * 4)); PV_CUDA(cudaMalloc(&dw1, dw_floats * 4));
    PV_CUDA(cudaMalloc(&db0, 256 * 4)); PV_CUDA(cudaMalloc(&db1, 256 * 4)); PV_CUDA(cudaMalloc(&part, part_floats * 4));
    std::vector<float> h(std::max(in_floats, gz_floats));
    unsigned s = 777u;
    auto gen = [&](size_t n, int range, float scale) {
        for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h[i] = (float)((int)((s >> 16) % (2 * range + 1)) - range) * scale; }
    };
    gen(in_floats, 4, 0.125f);
    PV_CUDA(cudaMemcpy(x, h.data(), in_floats * 4, cudaMemcpyHostToDevice));
    gen(gz_floats, 4, 0.125f);
    for (long long r = 0; r < (long long)(gz_floats / p.n); ++r) {          // the engine's invariant: gz is zero outside the valid extent
        const long long q = r - p.og.lead;
        const bool ok = q >= 0 && q < (long long)B * p.og.pstride && row_valid(p.og, (int)(q % p.og.pstride));
        if (!ok) for (int c = 0; c < p.n; ++c) h[(size_t)r * p.n + c] = 0.f;
    }
    PV_CUDA(cudaMemcpy(gz, h.data(), gz_floats * 4, cudaMemcpyHostToDevice));
    PV_CUDA(cudaMemset(dw0, 0, dw_floats * 4)); PV_CUDA(cudaMemset(dw1, 0, dw_floats * 4));
    PV_CUDA(cudaMemset(db0, 0, 1024)); PV_CUDA(cudaMemset(db1, 0, 1024));
    p.x = x; p.gz = gz; p.B = B; p.in_lead = ig.lead; p.in_pstride = ig.pstride;
    RowWgradP q = p;
    q.dw = dw0; q.db = db0; p.dw = dw1; p.db = db1;
    int rc = launch_rowwgrad_simt(q, 0);
    if (!rc) rc = launch_rowwgrad_tc(p, 0, part, part_floats);
    if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = set_error(PV_ERR_CUDA, "selftest %s: %s", name, cudaGetErrorString(cudaGetLastError()));
    double worst = 0; size_t bad = 0;
    if (!rc) {
        std::vector<float> a(dw_floats + 256), c2(dw_floats + 256);
        cudaMemcpy(a.data(), dw0, dw_floats * 4, cudaMemcpyDeviceToHost); cudaMemcpy(c2.data(), dw1, dw_floats * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(a.data() + dw_floats, db0, p.n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(c2.data() + dw_floats, db1, p.n * 4, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < dw_floats + p.n; ++i) { const double d = std::fabs((double)a[i] - c2[i]); if (!(d <= 1e-2)) ++bad; if (d > worst || d != d) worst = d; }
    }
    char line[256];
    snprintf(line, sizeof line, "%-34s %s max|tc - simt| = %.3g, mismatches %zu of %zu%s%s\n", name, (!rc && bad == 0) ? "PASS" : "FAIL",
             worst, bad, dw_floats + p.n, rc ? " : " : "", rc ? last_error().c_str() : "");
    rep += line;
    cudaFree(x); cudaFree(gz); cudaFree(dw0); cudaFree(dw1); cudaFree(db0); cudaFree(db1); cudaFree(part);
    return (!rc && bad == 0) ? 0 : 1;
}

int tc_selftest(std::string& rep) {
    const int B = 3;
    int fails = 0;
    auto base = [&](const RowGeom& ig, const RowGeom& og, int xc, int n, const Taps& tp, int nchunk) {
        RowConvP p;
        memset(&p, 0, sizeof p);
        p.xc = xc; p.n = n; p.B = B; p.in_lead = ig.lead; p.in_pstride = ig.pstride; p.og = og; p.kc = 32; p.w_kmajor = 1;
        int k = 0;
        for (int i = 0; i < tp.n; ++i) for (int j = 0; j < nchunk; ++j, ++k) { p.off[k] = tp.off[i]; p.c0[k] = tp.c0[i] + 32 * j; p.wr0[k] = 0; p.wc0[k] = 32 * k; }
        p.ntap = k; p.w_rows = n; p.w_cols = 32 * k;
        return p;
    };
    auto floats = [&](const RowGeom& g, int C) { return (size_t)(g.lead + (long long)B * g.pstride + ROW_TAIL) * C; };
    const RowGeom pr = pr_geom();
    float* dummy = reinterpret_cast<float*>(1);
    {   // normConv forward: 27 centred taps, bias + residual
        RowConvP p = base(pr, pr, 32, 32, conv3_taps(529, 23, true, +1), 1);
        p.bias = dummy; p.residual = dummy;
        fails += selftest_one("conv3 same fwd (+bias +residual)", p, floats(pr, 32), floats(pr, 32), (size_t)p.w_rows * p.w_cols, rep, &pr);
    }
    {   // normConv data gradient: negated taps, ReLU mask, tf32-rounded output
        RowConvP p = base(pr, pr, 32, 32, conv3_taps(529, 23, true, -1), 1);
        p.relumask = dummy; p.round_tf32 = 1;
        fails += selftest_one("conv3 same dgrad (+relu mask)", p, floats(pr, 32), floats(pr, 32), (size_t)p.w_rows * p.w_cols, rep, &pr);
    }
    {   // reducer forward (valid taps, G1 -> G2) with ReLU
        RowConvP p = base(g_geom(1), g_geom(2), 32, 32, conv3_taps(576, 24, false, +1), 1);
        p.bias = dummy; p.relu = 1;
        fails += selftest_one("conv3 valid fwd G1->G2 (+relu)", p, floats(g_geom(1), 32), floats(g_geom(2), 32), (size_t)p.w_rows * p.w_cols, rep);
    }
    {   // reducer data gradient (G2 -> G1)
        RowConvP p = base(g_geom(2), g_geom(1), 32, 32, conv3_taps(576, 24, false, -1), 1);
        p.relumask = dummy;
        fails += selftest_one("conv3 valid dgrad G2->G1", p, floats(g_geom(2), 32), floats(g_geom(1), 32), (size_t)p.w_rows * p.w_cols, rep);
    }
    {   // expConv forward: 1 tap, N = 256, ReLU
        RowConvP p = base(pr, pr, 32, 256, chunk_taps(32), 1);
        p.bias = dummy; p.relu = 1;
        fails += selftest_one("pointwise 32 -> 256 (+relu)", p, floats(pr, 32), floats(pr, 256), (size_t)p.w_rows * p.w_cols, rep);
    }
    {   // decConv forward: K = 256 as 8 chunks, N = 32
        RowConvP p = base(pr, pr, 256, 32, chunk_taps(256), 1);
        p.bias = dummy;
        fails += selftest_one("pointwise 256 -> 32", p, floats(pr, 256), floats(pr, 32), (size_t)p.w_rows * p.w_cols, rep);
    }
    fails += selftest_resfront(rep);
    fails += selftest_resback(rep);
    fails += selftest_x3(rep);
    auto wbase = [&](const RowGeom& og, int xc, int n, const Taps& tp) {
        RowWgradP p;
        memset(&p, 0, sizeof p);
        p.xc = xc; p.n = n; p.og = og; p.kc = 32; p.ntap = tp.n; p.dw_cols = n;
        for (int i = 0; i < tp.n; ++i) { p.off[i] = tp.off[i]; p.c0[i] = tp.c0[i]; p.dwr0[i] = 32 * tp.chunk[i]; p.dwc0[i] = 0; }
        return p;
    };
    fails += selftest_wgrad("wgrad conv3 same (PR)", wbase(pr, 32, 32, conv3_taps(529, 23, true, +1)), pr, B, rep);
    fails += selftest_wgrad("wgrad conv3 valid (G1 -> G2)", wbase(g_geom(2), 32, 32, conv3_taps(576, 24, false, +1)), g_geom(1), B, rep);
    fails += selftest_wgrad("wgrad pointwise x=256 (decConv)", wbase(pr, 256, 32, chunk_taps(256)), pr, B, rep);
    fails += selftest_wgrad("wgrad pointwise gz=256 (expConv)", wbase(pr, 32, 256, chunk_taps(32)), pr, B, rep);
    return fails;
}

// ------------------------------------------------------------------------------------------ plan
int tc_build_plan(pv_model* m) {
    const pv_cfg& c = m->cfg;
    const RowGeom pr = pr_geom(m->T);
    const int F = m->F, EX = F * c.exp_rate;
    if (!tail_plan_supported(m))
        return set_error(PV_ERR_BAD_CONFIG, "the row engine runs the T = 7, 9 and 13 reducer tails (modelsTF.py:62-67); "
                                            "num_low_res_imgs=%d needs precision fp32", m->T);
    const std::vector<TailStep> tail = tail_plan(m);
    const RowGeom ug = g_dims(1, m->P, m->P);
    for (int tr = 0; tr < 2; ++tr) {
        Pool& P = tr ? m->pool_train : m->pool_infer;
        P.add("xn", (size_t)m->S * m->S * m->T);
        P.add("mn", (size_t)m->S * m->S);
        for (int i = 0; i <= m->R; ++i) P.add(m->A(i, tr), rows_per(pr, F), rows_extra(pr, F));
        for (int i = 0; i < m->R; ++i) {
            if (!m->use_tc) P.add(m->E(i, tr), rows_per(pr, EX), rows_extra(pr, EX));   // tensor-core engine keeps E in TMEM
            else if (tr) {
                P.add("M" + std::to_string(i), rows_per(pr, 8), rows_extra(pr, 8));    // ... and only its ReLU bits (32 B / row)
                if (m->x3) P.add("MT" + std::to_string(i), (size_t)cdiv(pr.nrows, 128) * 1024);   // transposed copy for the weight-gradient kernel
            }
            P.add(m->D(i, tr), rows_per(pr, F), rows_extra(pr, F));
        }
        for (const TailStep& ts : tail) {
            if (ts.copy) P.add(ts.in, rows_per(ts.ig, F), rows_extra(ts.ig, F));
            P.add(ts.out, rows_per(ts.og, F), rows_extra(ts.og, F));
        }
        P.add("U", rows_per(ug, F), rows_extra(ug, F));
        if (m->x3) {                    // remainders (v - tf32(v)) of the tensors a compensated forward product reads, and their packed pair rows
            P.add("a_lo0", rows_per(pr, F), rows_extra(pr, F));
            P.add("a_lo1", rows_per(pr, F), rows_extra(pr, F));
            P.add("D_pack", rows_per(pr, F), rows_extra(pr, F));         // packed fp16 pair rows (rows.h): what a compensated conv3 reads
            for (const TailStep& ts : tail) {
                if (ts.copy) P.add(ts.in + "_pack", rows_per(ts.ig, F), rows_extra(ts.ig, F));
                P.add(ts.out + "_lo", rows_per(ts.og, F), rows_extra(ts.og, F));
                P.add(ts.out + "_pack", rows_per(ts.og, F), rows_extra(ts.og, F));
            }
        }
        for (int i = 0; i < c.scale; ++i) {
            const Layer& L = m->layers[m->li("residConv" + std::to_string(i + 1))];
            P.add("q" + std::to_string(i + 1), (size_t)L.Ho * L.Wo * L.cout_s);
        }
        if (tr) {
            P.add("g_a0", rows_per(pr, F), rows_extra(pr, F));
            P.add("g_a1", rows_per(pr, F), rows_extra(pr, F));
            P.add("g_D", rows_per(pr, F), rows_extra(pr, F));
            if (m->x3) P.add("g_pack", rows_per(pr, F), rows_extra(pr, F));      // bf16 pair rows of the block-input gradient (conv3_tc.cu MODE 2)
            if (!m->use_tc) P.add("g_E", rows_per(pr, EX), rows_extra(pr, EX));
            for (const TailStep& ts : tail) {
                if (ts.copy) P.add("g_" + ts.in, rows_per(ts.ig, F), rows_extra(ts.ig, F));
                P.add("g_" + ts.out, rows_per(ts.og, F), rows_extra(ts.og, F));
                if (m->x3) P.add("g_" + ts.out + "_pack", rows_per(ts.og, F), rows_extra(ts.og, F));
            }
            if (m->x3) P.add("g_U_pack", rows_per(ug, F), rows_extra(ug, F));
            P.add("g_U", rows_per(ug, F), rows_extra(ug, F));
            P.add("g_tail", (size_t)m->P * m->P * c.scale * c.scale);
            for (int i = 0; i + 1 < c.scale; ++i) {
                const Layer& L = m->layers[m->li("residConv" + std::to_string(i + 1))];
                P.add("g_q" + std::to_string(i + 1), (size_t)L.Ho * L.Wo * L.cout_s);
            }
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------ forward
// low-frequency skip path + depth_to_space + add + denormalise (modelsTF.py:38-53,73), shared by both forward variants
static int tc_forward_tail(pv_model* m, int B, float* sr, bool tr, int clip_round, cudaStream_t st) {
    Pool& P = tr ? m->pool_train : m->pool_infer;
    const pv_cfg& c = m->cfg;
    const RowGeom ug = g_dims(1, m->P, m->P);
    const float* q = P["mn"];
    if (c.scale == 3 && skip2d_supported(m->S, c.scale * c.scale)) {   // WDSRNetLRResidualPath, modelsTF.py:45-53: one fused kernel
        const Layer &R1 = m->layers[m->li("residConv1")], &R2 = m->layers[m->li("residConv2")], &R3 = m->layers[m->li("residConv3")];
        // ... with depth_to_space + add + de-normalise fused in: ONE kernel after the upscale conv (north_star bullet 2)
        return launch_skip2d_fwd_tail(P["mn"], m->weff + R1.weff_off, m->bias_s + R1.bias_s_off, m->weff + R2.weff_off, m->bias_s + R2.bias_s_off,
                                      m->weff + R3.weff_off, m->bias_s + R3.bias_s_off, B, m->S, c.scale * c.scale, P["q1"], P["q2"], P["q3"],
                                      P["U"], ug, m->F, c.scale, c.mean, c.std, clip_round, sr, st);
    } else {
        for (int i = 0; i < c.scale; ++i) {
            float* out = P["q" + std::to_string(i + 1)];
            PV_TRY(conv_fwd(m, m->li("residConv" + std::to_string(i + 1)), q, out, nullptr, B, st));
            q = out;
        }
    }
    return launch_tail_rows(P["U"], ug, m->F, q, B, m->P, c.scale, c.mean, c.std, clip_round, sr, st);
}

// precision 4 ("tf32x3"): the same graph with every tensor-core product compensated (resblock_x3_tc.cu, conv_rows_x3).  The hi
// arrays (tf32(v)) are the ones the backward pass reads, exactly as in the single-pass engine; the lo arrays live only here.
static int tc_forward_x3(pv_model* m, int B, float* sr, bool tr, int clip_round, cudaStream_t st) {
    Pool& P = tr ? m->pool_train : m->pool_infer;
    const RowGeom pr = pr_geom(m->T);
    const int F = m->F;
    const Taps same = conv3_taps(pr.plane, pr.pw, true, +1);
    const std::vector<TailStep> tail = tail_plan(m);
    const RowGeom ug = g_dims(1, m->P, m->P);
    auto alo = [&](int i) { return P[(i & 1) ? "a_lo1" : "a_lo0"]; };
    const Layer& L0 = m->layers[m->li("mainConv1")];
    PV_TRY(launch_first_conv_pr(P["xn"], m->weff + L0.weff_off, m->bias_s + L0.bias_s_off, B, m->S, m->T, P[m->A(0, tr)], pr, st, alo(0)));
    for (int i = 0; i < m->R; ++i) {                                   // ResConv3D, modelsTF.py:177-189
        const int e = m->li("expConv_" + std::to_string(i));
        const Layer &Le = m->layers[e], &Ld = m->layers[e + 1];
        const double fl = 2.0 * B * Le.Ho * Le.Wo * Le.To * ((double)Le.cin * Le.cout + (double)Ld.cin * Ld.cout);
        PV_TRY(launch_resfront_fwd_x3_tc(P[m->A(i, tr)], alo(i), m->weffT + Le.weff_off, m->weffT_lo + Le.weff_off, m->weffT_pack + Ld.weff_off,
                                         m->bias_s + Le.bias_s_off, m->bias_s + Ld.bias_s_off, P[m->D(i, tr)], P["D_pack"],
                                         tr ? reinterpret_cast<uint32_t*>(P["M" + std::to_string(i)]) : nullptr,
                                         tr ? reinterpret_cast<uint32_t*>(P["MT" + std::to_string(i)]) : nullptr, pr, B, fl, st, 1));
        PV_TRY(conv_rows_x3(m, m->layers[e + 2], same, P["D_pack"], pr, P[m->A(i + 1, tr)], alo(i + 1), nullptr, pr, P[m->A(i, tr)], alo(i), B,
                            "norm_fwd_x3", st));
    }
    const Taps valid = conv3_taps(576, 24, false, +1);
    for (size_t k = 0; k < tail.size(); ++k) {
        const TailStep& ts = tail[k];
        const float* in_pack;
        if (k == 0 || ts.copy) {
            // re-layout (+ reflect pad) of the (hi, lo) pair in one launch: hi rows for the backward pass, packed pair rows for the conv
            // (rows the copy does not write stay zero in both)
            const float* src_hi = k == 0 ? P[m->A(m->R, tr)] : P[tail[k - 1].out];
            const float* src_lo = k == 0 ? alo(m->R) : P[tail[k - 1].out + "_lo"];
            const RowGeom sg = k == 0 ? pr : tail[k - 1].og;
            PV_TRY(launch_pr_to_g_reflect(src_hi, sg, P[ts.in], ts.ig, B, F, st, ts.pad, src_lo, P[ts.in + "_pack"]));
            in_pack = P[ts.in + "_pack"];
        } else {
            in_pack = P[tail[k - 1].out + "_pack"];
        }
        const Layer& L = m->layers[m->li("convReducer_" + std::to_string(k + 1))];
        // a reducer's output feeds the next conv directly (packed rows) or through a reflect-pad copy (which needs the fp32 lo half)
        const bool next_copies = k + 1 < tail.size() && tail[k + 1].copy;
        PV_TRY(conv_rows_x3(m, L, valid, in_pack, ts.ig, P[ts.out], next_copies ? P[ts.out + "_lo"] : nullptr, P[ts.out + "_pack"], ts.og, nullptr,
                            nullptr, B, "reducer_fwd_x3", st));
    }
    PV_TRY(conv_rows_x3(m, m->layers[m->li("upscaleConv1")], valid, P[tail.back().out + "_pack"], tail.back().og, P["U"], nullptr, nullptr, ug, nullptr,
                        nullptr, B, "upscale_fwd_x3", st));
    return tc_forward_tail(m, B, sr, tr, clip_round, st);
}

int tc_forward(pv_model* m, const float* lr, int B, float* sr, bool tr, int clip_round, cudaStream_t st) {
    Pool& P = tr ? m->pool_train : m->pool_infer;
    PV_TRY(P.ensure(B, st));
    PV_TRY(refresh_weights(m, st));
    const pv_cfg& c = m->cfg;
    const RowGeom pr = pr_geom(m->T);
    const int F = m->F, EX = F * c.exp_rate;
    const Taps same = conv3_taps(pr.plane, pr.pw, true, +1);
    const Taps one = chunk_taps(32), wide = chunk_taps(EX);
    const std::vector<TailStep> tail = tail_plan(m);
    const RowGeom ug = g_dims(1, m->P, m->P);

    PV_TRY(launch_prep(lr, B, m->S * m->S, m->T, c.mean, c.std, P["xn"], P["mn"], st));
    const Layer& L0 = m->layers[m->li("mainConv1")];
    if (m->x3) return tc_forward_x3(m, B, sr, tr, clip_round, st);
    PV_TRY(launch_first_conv_pr(P["xn"], m->weff + L0.weff_off, m->bias_s + L0.bias_s_off, B, m->S, m->T, P[m->A(0, tr)], pr, st, nullptr,
                                m->use_tc ? 1 : 0));
    // Inference on the tensor cores: the decay output only feeds the block's 3x3x3 conv, so the fused kernel stores it as fp16 rows
    // (tf32-exact values are fp16-exact) and the conv runs K = 16 MMAs over them: half the operand fetch of the tf32 form, same result.
    // (Training keeps fp32 rows: the weight-gradient kernels read D as a tf32 operand.)  PV_INFER_TF32_ROWS=1 keeps the tf32 form (A/B).
    static const bool infer_tf32_rows = getenv("PV_INFER_TF32_ROWS") != nullptr;
    const bool f16_rows = !tr && m->use_tc && m->weffT_pack && !infer_tf32_rows;
    for (int i = 0; i < m->R; ++i) {                                   // ResConv3D, modelsTF.py:177-189
        const int e = m->li("expConv_" + std::to_string(i));
        const Layer &Le = m->layers[e], &Ld = m->layers[e + 1];
        if (m->use_tc) {                // expand -> ReLU -> decay in one kernel; E never leaves TMEM
            const double fl = 2.0 * B * Le.Ho * Le.Wo * Le.To * ((double)Le.cin * Le.cout + (double)Ld.cin * Ld.cout);
            PV_TRY(launch_resfront_fwd_tc(P[m->A(i, tr)], m->weffT + Le.weff_off, m->weffT + Ld.weff_off, m->bias_s + Le.bias_s_off,
                                          m->bias_s + Ld.bias_s_off, P[m->D(i, tr)],
                                          tr ? reinterpret_cast<uint32_t*>(P["M" + std::to_string(i)]) : nullptr, pr, B, 1, fl, st, f16_rows ? 1 : 0));
        } else {
            PV_TRY(conv_rows(m, Le, one, P[m->A(i, tr)], F, pr, P[m->E(i, tr)], pr, nullptr, B, "exp_fwd", st));
            PV_TRY(conv_rows(m, Ld, wide, P[m->E(i, tr)], EX, pr, P[m->D(i, tr)], pr, nullptr, B, "dec_fwd", st));
        }
        PV_TRY(conv_rows(m, m->layers[e + 2], same, P[m->D(i, tr)], F, pr, P[m->A(i + 1, tr)], pr, P[m->A(i, tr)], B, "norm_fwd", st, true, f16_rows));
    }
    // ConvReduceAndUpscale (T = 9: modelsTF.py:152-164, reflect pad H,W by 1 before reducer 1), v2 (T = 7: :166-175, no pad),
    // v3 (T = 13: :123-150, reflect pad before reducers 1-3): valid 3x3x3 + ReLU reducers, then the upscale conv
    const Taps valid = conv3_taps(576, 24, false, +1);
    for (size_t k = 0; k < tail.size(); ++k) {
        const TailStep& ts = tail[k];
        if (k == 0) PV_TRY(launch_pr_to_g_reflect(P[m->A(m->R, tr)], pr, P[ts.in], ts.ig, B, F, st, ts.pad));
        else if (ts.copy) PV_TRY(launch_pr_to_g_reflect(P[tail[k - 1].out], tail[k - 1].og, P[ts.in], ts.ig, B, F, st, ts.pad));
        const Layer& L = m->layers[m->li("convReducer_" + std::to_string(k + 1))];
        PV_TRY(conv_rows(m, L, valid, P[ts.in], F, ts.ig, P[ts.out], ts.og, nullptr, B, "reducer_fwd", st));
    }
    PV_TRY(conv_rows(m, m->layers[m->li("upscaleConv1")], valid, P[tail.back().out], F, tail.back().og, P["U"], ug, nullptr, B, "upscale_fwd", st, false));
    return tc_forward_tail(m, B, sr, tr, clip_round, st);
}

// ------------------------------------------------------------------------------------------ backward
int tc_bucket_split_layer(const pv_model* m) { return m->li("expConv_" + std::to_string(m->R / 2)); }

int tc_backward(pv_trainer* t, const float* g_sr, int B, cudaStream_t st, int stage) {
    pv_model* m = t->m;
    Pool& P = m->pool_train;
    const pv_cfg& c = m->cfg;
    const RowGeom pr = pr_geom(m->T);
    const int F = m->F, EX = F * c.exp_rate, R = m->R;
    const std::vector<TailStep> tail = tail_plan(m);
    const RowGeom ug = g_dims(1, m->P, m->P);
    const Taps same = conv3_taps(pr.plane, pr.pw, true, +1), same_T = conv3_taps(pr.plane, pr.pw, true, -1);
    const Taps valid = conv3_taps(576, 24, false, +1), valid_T = conv3_taps(576, 24, false, -1);
    const Taps one = chunk_taps(32), wide = chunk_taps(EX);

    const int isplit = R / 2;                  // blocks R-1 .. isplit belong to bucket 0
    const int lsplit = tc_bucket_split_layer(m), nlayers = (int)m->layers.size();
    // finalises the gradients of layers [la, lb): pending partial reductions, then the weight-norm backward of that range
    auto finish = [&](int la, int lb) -> int {
        PV_TRY(launch_deferred_reduce(t->rq, st));
        return launch_wn_bwd(m->wn_tab, nlayers, m->wn_first[lb] - m->wn_first[la], m->params, m->scale, t->dweff, t->dbias_s, t->grads, st,
                             m->wn_first[la]);
    };
    if (stage != 1) {
    t->rq.reset(t->wg_partial_floats);        // deferred partial reductions of this pass: one launch per bucket, before wn_bwd
    PV_CUDA(cudaMemsetAsync(t->dweff, 0, m->nweff * sizeof(float), st));
    PV_CUDA(cudaMemsetAsync(t->dbias_s, 0, m->nbias_s * sizeof(float), st));
    // precision 4: every gradient a 3x3x3 data gradient reads also exists as bf16 pair rows ("..._pack"), written by its producer
    auto gpack = [&](const std::string& name) -> float* { return m->x3 ? P[name + "_pack"] : nullptr; };
    PV_TRY(launch_tail_bwd_rows(g_sr, B, m->P, c.scale, c.std, P["g_U"], ug, F, P["g_tail"], st, m->use_tc ? 1 : 0, gpack("g_U")));
    if (c.scale == 3 && skip2d_supported(m->S, c.scale * c.scale) &&
        skip2d_partial_floats(B, m->S, c.scale * c.scale) <= t->wg_partial_floats) {   // ---- 2-D skip path: one fused kernel + a fixed-order reduction
        const Layer &R1 = m->layers[m->li("residConv1")], &R2 = m->layers[m->li("residConv2")], &R3 = m->layers[m->li("residConv3")];
        PV_TRY(launch_skip2d_bwd(P["mn"], P["q1"], P["q2"], P["g_tail"], m->weff + R2.weff_off, m->weff + R3.weff_off, B, m->S,
                                 c.scale * c.scale, t->wg_partials, t->wg_partial_floats, t->dweff + R1.weff_off, t->dweff + R2.weff_off,
                                 t->dweff + R3.weff_off, t->dbias_s + R1.bias_s_off, t->dbias_s + R2.bias_s_off, t->dbias_s + R3.bias_s_off, st, &t->rq));
    } else {   // ---- 2-D skip path (dense kernels)
        const float* gout = P["g_tail"];
        for (int i = c.scale; i >= 1; --i) {
            const int id = m->li("residConv" + std::to_string(i));
            const float* in = (i == 1) ? P["mn"] : P["q" + std::to_string(i - 1)];
            const float* ref = m->layers[id].relu ? P["q" + std::to_string(i)] : nullptr;
            PV_TRY(conv_wgrad(t, id, in, gout, ref, B, st));
            if (i > 1) {
                float* gin = P["g_q" + std::to_string(i - 1)];
                PV_TRY(conv_dgrad(m, id, gout, ref, gin, nullptr, B, st));
                gout = gin;
            }
        }
    }
    {   // ---- upscale conv, then the reducers (G layouts); each dgrad applies the ReLU mask of the layer below
        const Layer& U = m->layers[m->li("upscaleConv1")];
        const TailStep& last = tail.back();
        PV_TRY(wgrad_rows(t, U, valid, P[last.out], F, last.og, P["g_U"], ug, B, "upscale_wgrad", st));
        PV_TRY(dgrad_rows(m, U, valid_T, 1, P["g_U"], ug, P["g_" + last.out], last.og, nullptr, P[last.out], B, "upscale_dgrad", st,
                          gpack("g_U"), gpack("g_" + last.out)));
        for (int k = (int)tail.size() - 1; k >= 0; --k) {
            const TailStep& ts = tail[k];
            const Layer& L = m->layers[m->li("convReducer_" + std::to_string(k + 1))];
            PV_TRY(wgrad_rows(t, L, valid, P[ts.in], F, ts.ig, P["g_" + ts.out], ts.og, B, "reducer_wgrad", st));
            // the input is the previous reducer's ReLU output itself (mask here) or a padded copy of it (mask in the pad adjoint)
            // (a padded copy's gradient goes through the pad adjoint next, which reads fp32 rows: no pair rows of it)
            PV_TRY(dgrad_rows(m, L, valid_T, 1, P["g_" + ts.out], ts.og, P["g_" + ts.in], ts.ig, nullptr, (k > 0 && !ts.copy) ? P[ts.in] : nullptr,
                              B, "reducer_dgrad", st, gpack("g_" + ts.out), ts.copy ? nullptr : gpack("g_" + ts.in)));
            if (k > 0 && ts.copy)
                PV_TRY(launch_pr_to_g_reflect_bwd(P["g_" + ts.in], ts.ig, P["g_" + tail[k - 1].out], tail[k - 1].og, B, F, st, ts.pad, P[tail[k - 1].out],
                                                  m->use_tc ? 1 : 0, gpack("g_" + tail[k - 1].out)));
        }
        PV_TRY(launch_pr_to_g_reflect_bwd(P["g_" + tail[0].in], tail[0].ig, P["g_a" + std::to_string(R & 1)], pr, B, F, st, tail[0].pad, nullptr,
                                          m->use_tc ? 1 : 0, m->x3 ? P["g_pack"] : nullptr));
    }
    }   // stage != 1
    const int i_hi = stage == 1 ? isplit - 1 : R - 1, i_lo = stage == 0 ? isplit : 0;
    for (int i = i_hi; i >= i_lo; --i) {        // ---- residual blocks, last to first
        const int e = m->li("expConv_" + std::to_string(i));
        const Layer &Le = m->layers[e], &Ld = m->layers[e + 1], &Ln = m->layers[e + 2];
        const float* G = P["g_a" + std::to_string((i + 1) & 1)];
        float* gin = P["g_a" + std::to_string(i & 1)];
        PV_TRY(wgrad_rows(t, Ln, same, P[m->D(i, true)], F, pr, G, pr, B, "norm_wgrad", st));
        // (precision 4: block i + 1's backward-data kernel, or the tail's pad adjoint, left G as bf16 pair rows too)
        PV_TRY(dgrad_rows(m, Ln, same_T, 1, G, pr, P["g_D"], pr, nullptr, nullptr, B, "norm_dgrad", st,
                          m->x3 ? P["g_pack"] : nullptr));
        if (m->use_tc) {                // expand/decay backward on chip: E and gZ are recomputed in TMEM, never stored
            const double fl = 2.0 * B * Le.Ho * Le.Wo * Le.To * ((double)Le.cin * Le.cout + (double)Ld.cin * Ld.cout);
            PV_TRY(launch_resfront_bwd_weight_tc(P[m->A(i, true)], P["g_D"], m->weffT + Le.weff_off, m->weff + Ld.weff_off,
                                                 m->bias_s + Le.bias_s_off, t->dweff + Ld.weff_off, t->dweff + Le.weff_off,
                                                 t->dbias_s + Le.bias_s_off, t->dbias_s + Ld.bias_s_off, pr, B, t->wg_partials,
                                                 t->wg_partial_floats, fl, st, &t->rq,
                                                 m->x3 ? reinterpret_cast<const uint32_t*>(P["MT" + std::to_string(i)]) : nullptr));
            PV_TRY(launch_resfront_bwd_data_tc(P["g_D"], m->weff + Ld.weff_off, m->weff + Le.weff_off,
                                               reinterpret_cast<const uint32_t*>(P["M" + std::to_string(i)]), G,
                                               i == 0 ? P[m->A(0, true)] : nullptr, gin, pr, B, 1, fl, st,
                                               m->x3 ? m->weff_lo + Ld.weff_off : nullptr, m->x3 ? m->weff_lo + Le.weff_off : nullptr,
                                               (m->x3 && i > 0) ? P["g_pack"] : nullptr));
            continue;
        }
        PV_TRY(wgrad_rows(t, Ld, wide, P[m->E(i, true)], EX, pr, P["g_D"], pr, B, "dec_wgrad", st));
        PV_TRY(dgrad_rows(m, Ld, one, 1, P["g_D"], pr, P["g_E"], pr, nullptr, P[m->E(i, true)], B, "dec_dgrad", st));
        PV_TRY(wgrad_rows(t, Le, one, P[m->A(i, true)], F, pr, P["g_E"], pr, B, "exp_wgrad", st));
        // expConv data gradient + the skip connection; block 0 flows into mainConv1's ReLU
        PV_TRY(dgrad_rows(m, Le, one, EX / 32, P["g_E"], pr, gin, pr, G, i == 0 ? P[m->A(0, true)] : nullptr, B, "exp_dgrad", st));
    }
    if (stage == 0) return finish(lsplit, nlayers);
    const Layer& L0 = m->layers[m->li("mainConv1")];
    PV_TRY(launch_first_conv_pr_wgrad(P["xn"], P["g_a0"], B, m->S, m->T, pr, t->dweff + L0.weff_off, t->dbias_s + L0.bias_s_off,
                                      t->wg_partials, t->wg_partial_floats, st, &t->rq));
    return finish(0, stage == 1 ? lsplit : nlayers);
}

}  // namespace pv
