// engine.cu -- graph plan, arenas, forward / backward / optimizer sequencing, and the extern "C" ABI
// declared in include/probav_b200.h.
//
// Replaces, behind the C-ABI:
//   WDSRConv3D.build                     reference models/modelsTF.py:15-43   -> pv_model_create (plan below)
//   WDSRNetHRResidualPath / ResConv3D    modelsTF.py:55-74, 177-189           -> Model::forward trunk
//   ConvReduceAndUpscale{,v2,v3,Ex}      modelsTF.py:76-175                   -> reducer plan (T = 9, 7, 13, 19)
//   WDSRNetLRResidualPath                modelsTF.py:45-53                    -> 2-D skip (conv layers with T = kt = 1)
//   ModelTrainer.trainStep / testStep    models/trainClass.py:124-143         -> Trainer::forward_backward / apply / eval
//   tf.keras.optimizers.Nadam/Adam/SGD   train.py:76-83                       -> Trainer::apply (host scalars + one kernel)
//   resolve / resolveByBatch / reconstruct_from_patches   test.py:114-160     -> pv_resolve*, pv_predict_*
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <atomic>

#include "kernels.h"

namespace pv {

// ------------------------------------------------------------------------------------------ error plumbing
std::string& last_error() {
    static thread_local std::string e;
    return e;
}
int set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches += n; }
int64_t launch_count() { return g_launches.load(); }
bool pdl_simple_enabled() { static const bool on = [] { const char* e = getenv("PV_PDL_SIMPLE"); return e && e[0] == '1'; }(); return on; }
bool pdl_enabled() { static const bool on = getenv("PV_NO_PDL") == nullptr; return on; }

// ------------------------------------------------------------------------------------------ kernel timing
namespace {
struct TimedLaunch { const char* name; cudaEvent_t a, b; double flops, bytes, exec_flops; };
std::vector<TimedLaunch> g_timed;
bool g_timing = false;
}  // namespace

KernelTimer::KernelTimer(const char* name, cudaStream_t s, double flops, double bytes, double exec_flops) : slot(-1), st(s) {
    if (!g_timing) return;
    TimedLaunch t{name, nullptr, nullptr, flops, bytes, exec_flops > 0.0 ? exec_flops : flops};
    if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) return;
    cudaEventRecord(t.a, st);
    slot = (int)g_timed.size();
    g_timed.push_back(t);
}
KernelTimer::~KernelTimer() {
    if (slot >= 0) cudaEventRecord(g_timed[slot].b, st);
}

}  // namespace pv

#include "engine.h"

using namespace pv;

namespace pv {

// ------------------------------------------------------------------------------------------ plan
static int add_layer(pv_model* m, const std::string& name, int kh, int kw, int kt, int cin, int cout, bool same,
                     int relu, int Hi, int Wi, int Ti) {
    Layer L;
    L.name = name;
    L.k[0] = kh; L.k[1] = kw; L.k[2] = kt;
    L.cin = cin; L.cout = cout;
    L.cin_s = storage_channels(cin); L.cout_s = storage_channels(cout);
    const bool is3d = name.rfind("residConv", 0) != 0;
    if (m->rows && is3d) {              // row engine: every trunk tensor is [rows][32k] and taps run (dt,dh,dw)
        L.wn_mode = 1;
        if (cout > 1) L.cout_s = ((cout + 31) / 32) * 32;
    }
    for (int a = 0; a < 3; ++a) L.pad[a] = same ? L.k[a] / 2 : 0;
    L.relu = relu;
    L.Hi = Hi; L.Wi = Wi; L.Ti = Ti;
    L.Ho = Hi + 2 * L.pad[0] - (kh - 1); L.Wo = Wi + 2 * L.pad[1] - (kw - 1); L.To = Ti + 2 * L.pad[2] - (kt - 1);
    const long long K = (long long)L.taps() * cin;
    L.v_off = m->nparams; m->nparams += K * cout;
    L.g_off = m->nparams; m->nparams += cout;
    L.b_off = m->nparams; m->nparams += cout;
    L.weff_off = m->nweff; m->nweff += (long long)L.taps() * L.cin_s * L.cout_s;
    L.bias_s_off = m->nbias_s; m->nbias_s += L.cout_s;
    L.scale_off = m->nscale; m->nscale += 2 * cout;
    const bool is2d = (kt == 1 && Ti == 1 && name.rfind("residConv", 0) == 0);
    PInfo pv_;
    pv_.name = name + "/v";
    if (is2d) { pv_.rank = 4; pv_.shape[0] = kh; pv_.shape[1] = kw; pv_.shape[2] = cin; pv_.shape[3] = cout; pv_.shape[4] = 0; }
    else { pv_.rank = 5; pv_.shape[0] = kh; pv_.shape[1] = kw; pv_.shape[2] = kt; pv_.shape[3] = cin; pv_.shape[4] = cout; }
    pv_.off = L.v_off; pv_.numel = K * cout;
    m->pinfo.push_back(pv_);
    PInfo pg; pg.name = name + "/g"; pg.rank = 1; pg.shape[0] = cout; pg.shape[1] = pg.shape[2] = pg.shape[3] = pg.shape[4] = 0;
    pg.off = L.g_off; pg.numel = cout;
    m->pinfo.push_back(pg);
    PInfo pb = pg; pb.name = name + "/bias"; pb.off = L.b_off;
    m->pinfo.push_back(pb);
    m->layers.push_back(L);
    return (int)m->layers.size() - 1;
}

static int build_plan(pv_model* m) {
    const pv_cfg& c = m->cfg;
    if (c.num_res_blocks < 0 || c.scale < 1 || c.num_filters < 1 || c.exp_rate < 1 || c.patch_size < 1 || c.max_shift < 0)
        return set_error(PV_ERR_BAD_CONFIG, "bad cfg value");
    if (c.kernel_size != 3)
        return set_error(PV_ERR_BAD_CONFIG, "kernel_size=%d: only 3 is supported (cfg/p16t9c85r12.cfg:20)", c.kernel_size);
    if (!c.is_grayscale)
        return set_error(PV_ERR_BAD_CONFIG, "is_grayscale=0 (3-channel input) is not built; PROBA-V bands are single-channel");
    const int ks = c.kernel_size;
    if (c.precision != 0 && c.precision != 1 && c.precision != 3 && c.precision != 4)
        return set_error(PV_ERR_BAD_CONFIG, "precision=%d: 0 = fp32 (CUDA cores), 1 = tf32 (tcgen05 tensor cores), 3 = fp32 on the row layouts, "
                                            "4 = error-compensated tf32 (3 x tf32 forward, split-weight data gradients)", c.precision);
    m->rows = c.precision != 0;
    m->use_tc = c.precision == 1 || c.precision == 4;
    m->x3 = c.precision == 4;
    if (m->rows && ((c.num_low_res_imgs != 7 && c.num_low_res_imgs != 9 && c.num_low_res_imgs != 13) || c.num_filters != 32 || c.num_filters * c.exp_rate != 256 || c.scale != 3 ||
                    c.patch_size != 16 || c.max_shift != 6 || (int)(c.num_filters * c.decay_rate) > 32))
        return set_error(PV_ERR_BAD_CONFIG, "the tensor-core engine is built for the p16 family (T in {7, 9, 13}, 32 filters, exp_rate 8, scale 3, "
                                            "patch 16, max_shift 6); use precision fp32 for other graphs");
    m->S = c.patch_size + c.max_shift;            // modelsTF.py:19
    m->T = c.num_low_res_imgs; m->P = c.patch_size; m->F = c.num_filters; m->R = c.num_res_blocks;
    const int S = m->S, T = m->T, F = m->F;
    const int dec = (int)(F * c.decay_rate);      // int(numFilters*decayRate), modelsTF.py:182
    if (dec < 1) return set_error(PV_ERR_BAD_CONFIG, "int(num_filters*decay_rate) = %d", dec);
    // reducer table per modelsTF.py:62-69
    struct Red { int k; int p[3]; };
    std::vector<Red> reds;
    if (T == 9) { for (int i = 0; i < T / c.scale; ++i) reds.push_back({ks, {i == 0 ? 1 : 0, i == 0 ? 1 : 0, 0}}); }
    else if (T == 7) { for (int i = 0; i < T / c.scale; ++i) reds.push_back({ks, {0, 0, 0}}); }
    else if (T == 13) { for (int i = 0; i < 5; ++i) reds.push_back({ks, {i < 3 ? 1 : 0, i < 3 ? 1 : 0, 0}}); }
    else if (T == 19) {
        const int pp[10][3] = {{2, 2, 2}, {2, 2, 1}, {2, 2, 0}, {2, 2, 0}, {1, 1, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int i = 0; i < 10; ++i) reds.push_back({i == 0 ? 5 : ks, {pp[i][0], pp[i][1], pp[i][2]}});
    } else
        return set_error(PV_ERR_BAD_CONFIG, "num_low_res_imgs=%d: the reference graph only has reducers for 7, 9, 13, 19 "
                                            "(modelsTF.py:62-69)", T);
    m->nred = (int)reds.size();

    add_layer(m, "mainConv1", ks, ks, ks, 1, F, true, 1, S, S, T);
    for (int i = 0; i < m->R; ++i) {
        add_layer(m, "expConv_" + std::to_string(i), 1, 1, 1, F, F * c.exp_rate, true, 1, S, S, T);
        add_layer(m, "decConv_" + std::to_string(i), 1, 1, 1, F * c.exp_rate, dec, true, 0, S, S, T);
        add_layer(m, "normConv_" + std::to_string(i), ks, ks, ks, dec, F, true, 0, S, S, T);
    }
    int H = S, W = S, TT = T;
    for (int i = 0; i < m->nred; ++i) {
        for (int a = 0; a < 3; ++a) m->red_pad.push_back(reds[i].p[a]);
        H += 2 * reds[i].p[0]; W += 2 * reds[i].p[1]; TT += 2 * reds[i].p[2];
        if (H < reds[i].k || W < reds[i].k || TT < reds[i].k) return set_error(PV_ERR_BAD_CONFIG, "reducer %d collapses", i + 1);
        const int id = add_layer(m, "convReducer_" + std::to_string(i + 1), reds[i].k, reds[i].k, reds[i].k, F, F, false, 1, H, W, TT);
        H = m->layers[id].Ho; W = m->layers[id].Wo; TT = m->layers[id].To;
    }
    if (H < ks || W < ks || TT < ks) return set_error(PV_ERR_BAD_CONFIG, "upscale conv input %dx%dx%d too small", H, W, TT);
    const int up = add_layer(m, "upscaleConv1", ks, ks, ks, F, c.scale * c.scale, false, 0, H, W, TT);
    if (m->layers[up].Ho != m->P || m->layers[up].Wo != m->P || m->layers[up].To != 1)
        return set_error(PV_ERR_BAD_CONFIG, "main path ends at %dx%dx%d but Reshape((patch,patch,scale^2)) needs %dx%dx1 "
                         "(modelsTF.py:71): max_shift must be 6 for this reducer", m->layers[up].Ho, m->layers[up].Wo,
                         m->layers[up].To, m->P, m->P);
    int h2 = S, cin = 1;
    for (int i = 0; i < c.scale; ++i) {
        if (h2 < ks) return set_error(PV_ERR_BAD_CONFIG, "2-D skip path collapses");
        const int id = add_layer(m, "residConv" + std::to_string(i + 1), ks, ks, 1, cin, c.scale * c.scale, false, i == 0, h2, h2, 1);
        h2 = m->layers[id].Ho; cin = c.scale * c.scale;
    }
    if (h2 != m->P)
        return set_error(PV_ERR_BAD_CONFIG, "2-D skip path ends at %dx%d, main path at %dx%d (Add at modelsTF.py:38 would fail)", h2, h2, m->P, m->P);

    if (m->rows) return tc_build_plan(m);
    // activation pools
    for (int tr = 0; tr < 2; ++tr) {
        Pool& P = tr ? m->pool_train : m->pool_infer;
        const size_t vox = (size_t)S * S * T;
        P.add("xn", vox);
        P.add("mn", (size_t)S * S);
        for (int i = 0; i <= m->R; ++i) P.add(m->A(i, tr), vox * F);
        for (int i = 0; i < m->R; ++i) {
            P.add(m->E(i, tr), vox * F * c.exp_rate);
            P.add(m->D(i, tr), vox * storage_channels(dec));
        }
        for (int i = 0; i < m->nred; ++i) {
            const Layer& L = m->layers[m->li("convReducer_" + std::to_string(i + 1))];
            if (reds[i].p[0] || reds[i].p[1] || reds[i].p[2]) P.add("pad" + std::to_string(i + 1), (size_t)L.Hi * L.Wi * L.Ti * F);
            P.add("r" + std::to_string(i + 1), (size_t)L.Ho * L.Wo * L.To * F);
        }
        P.add("U", (size_t)m->P * m->P * c.scale * c.scale);
        for (int i = 0; i < c.scale; ++i) {
            const Layer& L = m->layers[m->li("residConv" + std::to_string(i + 1))];
            P.add("q" + std::to_string(i + 1), (size_t)L.Ho * L.Wo * L.cout_s);
        }
        if (tr) {   // gradient buffers
            P.add("g_a0", vox * F); P.add("g_a1", vox * F);
            P.add("g_E", vox * F * c.exp_rate);
            P.add("g_D", vox * storage_channels(dec));
            for (int i = 0; i < m->nred; ++i) {
                const Layer& L = m->layers[m->li("convReducer_" + std::to_string(i + 1))];
                if (reds[i].p[0] || reds[i].p[1] || reds[i].p[2]) P.add("g_pad" + std::to_string(i + 1), (size_t)L.Hi * L.Wi * L.Ti * F);
                P.add("g_r" + std::to_string(i + 1), (size_t)L.Ho * L.Wo * L.To * F);
            }
            P.add("g_tail", (size_t)m->P * m->P * c.scale * c.scale);
            for (int i = 0; i + 1 < c.scale; ++i) {
                const Layer& L = m->layers[m->li("residConv" + std::to_string(i + 1))];
                P.add("g_q" + std::to_string(i + 1), (size_t)L.Ho * L.Wo * L.cout_s);
            }
        }
    }
    return 0;
}

static ConvP conv_desc(const Layer& L, const float* w, const float* bias, int B) {
    ConvP p;
    p.x = nullptr; p.xmask = nullptr; p.w = w; p.bias = bias; p.residual = nullptr; p.y = nullptr;
    p.B = B; p.Hi = L.Hi; p.Wi = L.Wi; p.Ti = L.Ti; p.Ho = L.Ho; p.Wo = L.Wo; p.To = L.To;
    p.cin = L.cin_s; p.cout = L.cout_s; p.kh = L.k[0]; p.kw = L.k[1]; p.kt = L.k[2];
    p.ph = L.pad[0]; p.pw = L.pad[1]; p.pt = L.pad[2]; p.relu = L.relu;
    p.cin_r = L.cin; p.cout_r = L.cout; p.tag = nullptr;
    return p;
}

int conv_fwd(pv_model* m, int li, const float* in, float* out, const float* res, int B, cudaStream_t st) {
    const Layer& L = m->layers[li];
    ConvP p = conv_desc(L, m->weff + L.weff_off, m->bias_s + L.bias_s_off, B);
    p.x = in; p.y = out; p.residual = res;
    return launch_conv(p, st);
}

// data gradient: correlate dy (masked by the layer's ReLU output) with the flipped/transposed kernel, pad' = k-1-pad
int conv_dgrad(pv_model* m, int li, const float* gout, const float* relu_ref, float* gin, const float* res, int B, cudaStream_t st) {
    const Layer& L = m->layers[li];
    ConvP p;
    p.x = gout; p.xmask = relu_ref; p.w = m->weffT + L.weff_off; p.bias = nullptr; p.residual = res; p.y = gin;
    p.B = B; p.Hi = L.Ho; p.Wi = L.Wo; p.Ti = L.To; p.Ho = L.Hi; p.Wo = L.Wi; p.To = L.Ti;
    p.cin = L.cout_s; p.cout = L.cin_s; p.kh = L.k[0]; p.kw = L.k[1]; p.kt = L.k[2];
    p.ph = L.k[0] - 1 - L.pad[0]; p.pw = L.k[1] - 1 - L.pad[1]; p.pt = L.k[2] - 1 - L.pad[2];
    p.relu = 0;
    p.cin_r = L.cout; p.cout_r = L.cin; p.tag = nullptr;
    return launch_conv(p, st);
}

int conv_wgrad(pv_trainer* t, int li, const float* in, const float* gout, const float* relu_ref, int B, cudaStream_t st) {
    const Layer& L = t->m->layers[li];
    WgradP p;
    p.x = in; p.dy = gout; p.ymask = relu_ref; p.dw = t->dweff + L.weff_off; p.db = t->dbias_s + L.bias_s_off;
    p.B = B; p.Hi = L.Hi; p.Wi = L.Wi; p.Ti = L.Ti; p.Ho = L.Ho; p.Wo = L.Wo; p.To = L.To;
    p.cin = L.cin_s; p.cout = L.cout_s; p.kh = L.k[0]; p.kw = L.k[1]; p.kt = L.k[2];
    p.ph = L.pad[0]; p.pw = L.pad[1]; p.pt = L.pad[2];
    p.cin_r = L.cin; p.cout_r = L.cout; p.tag = nullptr;
    p.partials = t->dense_partials; p.partial_floats = t->dense_partials ? WGRAD_PARTIAL_FLOATS : 0;   // fixed-order split reduction
    return launch_wgrad(p, st);
}

int refresh_weights(pv_model* m, cudaStream_t st) {
    if (!m->weff_dirty) return 0;
    PV_TRY(launch_wn_prep(m->wn_tab, (int)m->layers.size(), m->wn_blocks, m->params, m->weff, m->weffT, m->bias_s, m->scale, st,
                          m->weff_lo, m->weffT_lo, m->weffT_pack, m->weff_pack));
    m->weff_dirty = false;
    return 0;
}

// model(x): modelsTF.py:15-43
static int model_forward(pv_model* m, const float* lr, int B, float* sr, bool tr, int clip_round, cudaStream_t st) {
    if (!lr || !sr || B <= 0) return set_error(PV_ERR_BAD_ARG, "forward: null buffer or B=%d", B);
    PV_CUDA(cudaSetDevice(m->device));
    if (m->rows) return tc_forward(m, lr, B, sr, tr, clip_round, st);
    Pool& P = tr ? m->pool_train : m->pool_infer;
    PV_TRY(P.ensure(B, st));
    PV_TRY(refresh_weights(m, st));
    const pv_cfg& c = m->cfg;
    PV_TRY(launch_prep(lr, B, m->S * m->S, m->T, c.mean, c.std, P["xn"], P["mn"], st));
    PV_TRY(conv_fwd(m, m->li("mainConv1"), P["xn"], P[m->A(0, tr)], nullptr, B, st));
    for (int i = 0; i < m->R; ++i) {                                   // ResConv3D, modelsTF.py:177-189
        const int e = m->li("expConv_" + std::to_string(i));
        PV_TRY(conv_fwd(m, e, P[m->A(i, tr)], P[m->E(i, tr)], nullptr, B, st));
        PV_TRY(conv_fwd(m, e + 1, P[m->E(i, tr)], P[m->D(i, tr)], nullptr, B, st));
        PV_TRY(conv_fwd(m, e + 2, P[m->D(i, tr)], P[m->A(i + 1, tr)], P[m->A(i, tr)], B, st));
    }
    const float* cur = P[m->A(m->R, tr)];
    for (int i = 0; i < m->nred; ++i) {                                // ConvReduceAndUpscale*, modelsTF.py:76-175
        const int id = m->li("convReducer_" + std::to_string(i + 1));
        const Layer& L = m->layers[id];
        const int* rp = &m->red_pad[3 * i];
        if (rp[0] || rp[1] || rp[2]) {
            float* pd = P["pad" + std::to_string(i + 1)];
            PV_TRY(launch_reflect_pad(cur, pd, B, L.Hi - 2 * rp[0], L.Wi - 2 * rp[1], L.Ti - 2 * rp[2], m->F, rp[0], rp[1], rp[2], st));
            cur = pd;
        }
        float* out = P["r" + std::to_string(i + 1)];
        PV_TRY(conv_fwd(m, id, cur, out, nullptr, B, st));
        cur = out;
    }
    PV_TRY(conv_fwd(m, m->li("upscaleConv1"), cur, P["U"], nullptr, B, st));
    const float* q = P["mn"];
    for (int i = 0; i < c.scale; ++i) {                                // WDSRNetLRResidualPath, modelsTF.py:45-53
        float* out = P["q" + std::to_string(i + 1)];
        PV_TRY(conv_fwd(m, m->li("residConv" + std::to_string(i + 1)), q, out, nullptr, B, st));
        q = out;
    }
    PV_TRY(launch_tail(P["U"], q, B, m->P, c.scale, c.mean, c.std, clip_round, sr, st));
    return 0;
}

// tape.gradient(loss, trainable_variables): trainClass.py:131.  g_sr = dL/dSR [B, sP, sP].
static int model_backward(pv_trainer* t, const float* g_sr, int B, cudaStream_t st, int stage = -1) {
    pv_model* m = t->m;
    if (m->rows) return tc_backward(t, g_sr, B, st, stage);
    if (stage == 1) return 0;                  // dense engine: one bucket, everything is final after stage 0
    Pool& P = m->pool_train;
    const pv_cfg& c = m->cfg;
    PV_CUDA(cudaMemsetAsync(t->dweff, 0, m->nweff * sizeof(float), st));
    PV_CUDA(cudaMemsetAsync(t->dbias_s, 0, m->nbias_s * sizeof(float), st));
    PV_TRY(launch_tail_bwd(g_sr, B, m->P, c.scale, c.std, P["g_tail"], st));
    // ---- 2-D skip path
    {
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
    // ---- upscale conv + reducers
    const int R = m->R;
    {
        const int up = m->li("upscaleConv1");
        const float* in = m->nred ? P["r" + std::to_string(m->nred)] : P[m->A(R, true)];
        float* gin = m->nred ? P["g_r" + std::to_string(m->nred)] : P["g_a" + std::to_string(R & 1)];
        PV_TRY(conv_wgrad(t, up, in, P["g_tail"], nullptr, B, st));
        PV_TRY(conv_dgrad(m, up, P["g_tail"], nullptr, gin, nullptr, B, st));
    }
    for (int i = m->nred; i >= 1; --i) {
        const int id = m->li("convReducer_" + std::to_string(i));
        const Layer& L = m->layers[id];
        const int* rp = &m->red_pad[3 * (i - 1)];
        const bool padded = rp[0] || rp[1] || rp[2];
        const float* prev = (i == 1) ? P[m->A(R, true)] : P["r" + std::to_string(i - 1)];
        float* gprev = (i == 1) ? P["g_a" + std::to_string(R & 1)] : P["g_r" + std::to_string(i - 1)];
        const float* in = padded ? P["pad" + std::to_string(i)] : prev;
        const float* gout = P["g_r" + std::to_string(i)];
        const float* ref = P["r" + std::to_string(i)];
        PV_TRY(conv_wgrad(t, id, in, gout, ref, B, st));
        if (padded) {
            float* gp = P["g_pad" + std::to_string(i)];
            PV_TRY(conv_dgrad(m, id, gout, ref, gp, nullptr, B, st));
            PV_TRY(launch_reflect_pad_bwd(gp, gprev, B, L.Hi - 2 * rp[0], L.Wi - 2 * rp[1], L.Ti - 2 * rp[2], m->F, rp[0], rp[1], rp[2], st));
        } else {
            PV_TRY(conv_dgrad(m, id, gout, ref, gprev, nullptr, B, st));
        }
    }
    // ---- residual blocks, last to first
    for (int i = R - 1; i >= 0; --i) {
        const int e = m->li("expConv_" + std::to_string(i));
        const float* gout = P["g_a" + std::to_string((i + 1) & 1)];
        float* gin = P["g_a" + std::to_string(i & 1)];
        PV_TRY(conv_wgrad(t, e + 2, P[m->D(i, true)], gout, nullptr, B, st));
        PV_TRY(conv_dgrad(m, e + 2, gout, nullptr, P["g_D"], nullptr, B, st));
        PV_TRY(conv_wgrad(t, e + 1, P[m->E(i, true)], P["g_D"], nullptr, B, st));
        PV_TRY(conv_dgrad(m, e + 1, P["g_D"], nullptr, P["g_E"], nullptr, B, st));
        PV_TRY(conv_wgrad(t, e, P[m->A(i, true)], P["g_E"], P[m->E(i, true)], B, st));
        PV_TRY(conv_dgrad(m, e, P["g_E"], P[m->E(i, true)], gin, gout, B, st));      // + skip connection
    }
    PV_TRY(conv_wgrad(t, m->li("mainConv1"), P["xn"], P["g_a0"], P[m->A(0, true)], B, st));
    PV_TRY(launch_wn_bwd(m->wn_tab, (int)m->layers.size(), m->wn_blocks, m->params, m->scale, t->dweff, t->dbias_s, t->grads, st));
    return 0;
}

static int trainer_ensure(pv_trainer* t, int B) {
    if (B <= t->capB) return 0;
    cudaFree(t->sr); cudaFree(t->dsr); cudaFree(t->loss_ps); cudaFree(t->cpsnr_ps); cudaFree(t->best); cudaFree(t->cnt);
    const size_t hw = (size_t)t->m->P * t->m->cfg.scale * t->m->P * t->m->cfg.scale;
    PV_CUDA(cudaMalloc(&t->sr, hw * B * sizeof(float)));
    PV_CUDA(cudaMalloc(&t->dsr, hw * B * sizeof(float)));
    PV_CUDA(cudaMalloc(&t->loss_ps, B * sizeof(float)));
    PV_CUDA(cudaMalloc(&t->cpsnr_ps, B * sizeof(float)));
    PV_CUDA(cudaMalloc(&t->best, B * sizeof(int32_t)));
    PV_CUDA(cudaMalloc(&t->cnt, B * sizeof(int32_t)));
    t->capB = B;
    return 0;
}

// fwd -> loss (+ cPSNR metric in the same pass) [-> grads]
static int trainer_fwd_loss(pv_trainer* t, const float* lr, const float* hr, const uint8_t* mask, int B, float grad_scale,
                            bool backward, float* out_dev, cudaStream_t st, int stage = -1) {
    if (!lr || !hr || !mask || !out_dev || B <= 0) return set_error(PV_ERR_BAD_ARG, "step: null buffer or B=%d", B);
    pv_model* m = t->m;
    PV_TRY(trainer_ensure(t, B));
    // the training pool (which keeps every activation for the backward pass) only when gradients follow: evaluation runs on the
    // inference pool's ping-pong buffers (a validation batch must not grow the training pool)
    PV_TRY(model_forward(m, lr, B, t->sr, backward, 0, st));
    const int HW = m->P * m->cfg.scale;
    PV_TRY(shift_loss_device(t->loss_kind, hr, mask, t->sr, B, HW, HW, 3, grad_scale, t->loss_ps, t->best, t->cnt,
                             t->cpsnr_ps, out_dev, backward ? t->dsr : nullptr, nullptr, st));
    PV_TRY(launch_mean(t->cpsnr_ps, B, out_dev + 1, st));
    if (backward) PV_TRY(model_backward(t, t->dsr, B, st, stage));
    return 0;
}

// optimizer.apply_gradients: trainClass.py:132 (Keras Nadam per SURVEY Appendix B.7)
static int trainer_apply(pv_trainer* t, cudaStream_t st) {
    pv_model* m = t->m;
    const long long n = m->nparams;
    const double b1 = 0.9, b2 = 0.999, eps = 1e-7;
    const long long step = t->iter + 1;
    if (t->opt == PV_OPT_NADAM) {
        const double decay = 0.004;
        const double mu_t = b1 * (1.0 - 0.5 * std::pow(0.96, decay * step));
        const double mu_t1 = b1 * (1.0 - 0.5 * std::pow(0.96, decay * (step + 1)));
        const double Pt = t->momentum_cache * mu_t, Pt1 = Pt * mu_t1;
        t->momentum_cache = Pt;
        NadamScalars s;
        s.lr = t->lr; s.b1 = (float)b1; s.b2 = (float)b2; s.eps = (float)eps; s.mu_t = (float)mu_t; s.mu_t1 = (float)mu_t1;
        s.one_minus_Pt = (float)(1.0 - Pt); s.one_minus_Pt1 = (float)(1.0 - Pt1);
        s.one_minus_b2t = (float)(1.0 - std::pow(b2, (double)step));
        PV_TRY(launch_nadam(m->params, t->grads, t->m1, t->m2, n, s, st));
    } else if (t->opt == PV_OPT_ADAM) {
        const double lr_t = t->lr * std::sqrt(1.0 - std::pow(b2, (double)step)) / (1.0 - std::pow(b1, (double)step));
        PV_TRY(launch_adam(m->params, t->grads, t->m1, t->m2, n, (float)lr_t, (float)b1, (float)b2, (float)eps, st));
    } else {
        PV_TRY(launch_sgd(m->params, t->grads, n, t->lr, st));
    }
    t->iter = step;
    m->weff_dirty = true;
    return 0;
}

static cudaStream_t S_(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace pv

// =========================================================================================== extern "C"
extern "C" {

int pv_abi_version(void) { return PV_ABI_VERSION; }
const char* pv_last_error(void) { return pv::last_error().c_str(); }
int64_t pv_launch_count(void) { return pv::launch_count(); }

int pv_selftest(char* buf, int cap) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return set_error(PV_ERR_NO_DEVICE, "no CUDA device: no CPU fallback"); }
    std::string rep;
    const int fails = tc_selftest(rep);
    if (buf && cap > 0) { std::strncpy(buf, rep.c_str(), cap - 1); buf[cap - 1] = 0; }
    return fails;
}

int64_t pv_debug_read_buffer(pv_model* m, const char* name, int train, float* host, int64_t n) {
    if (!m || !name) return set_error(PV_ERR_BAD_ARG, "pv_debug_read_buffer: null argument");
    Pool& P = train ? m->pool_train : m->pool_infer;
    float* d = P[name];
    if (!d) return set_error(PV_ERR_BAD_ARG, "pv_debug_read_buffer: no buffer '%s' in the %s pool", name, train ? "training" : "inference");
    int64_t len = 0;
    for (auto& s : P.spec) if (s.name == name) len = (int64_t)(s.per * (size_t)P.cap + s.extra);
    PV_CUDA(cudaSetDevice(m->device));
    PV_CUDA(cudaDeviceSynchronize());
    if (host && n > 0) PV_CUDA(cudaMemcpy(host, d, (size_t)std::min(n, len) * sizeof(float), cudaMemcpyDeviceToHost));
    return len;
}

int pv_timing_enable(int on) { pv::g_timing = on != 0; return 0; }

int pv_timing_reset(void) {
    for (auto& t : pv::g_timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    pv::g_timed.clear();
    return 0;
}

int pv_timing_report(char* buf, int cap) {
    if (cudaDeviceSynchronize() != cudaSuccess) return set_error(PV_ERR_CUDA, "timing_report: device sync failed");
    struct Acc { long long n = 0; double ms = 0, flops = 0, bytes = 0, exec_flops = 0; };
    std::map<std::string, Acc> acc;
    for (auto& t : pv::g_timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.a, t.b) != cudaSuccess) { cudaGetLastError(); continue; }
        Acc& a = acc[t.name];
        a.n++; a.ms += ms; a.flops += t.flops; a.bytes += t.bytes; a.exec_flops += t.exec_flops;
    }
    std::string out;
    char line[256];
    for (auto& kv : acc) {
        snprintf(line, sizeof line, "%s %lld %.6f %.6e %.6e %.6e\n", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes, kv.second.exec_flops);
        out += line;
    }
    if (buf && cap > 0) { std::strncpy(buf, out.c_str(), cap - 1); buf[cap - 1] = 0; }
    return (int)out.size() + 1;
}

int pv_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

int pv_model_create(const pv_cfg* cfg, int device, pv_model** out) {
    if (!cfg || !out) return set_error(PV_ERR_BAD_ARG, "pv_model_create: null argument");
    *out = nullptr;
    pv_model* m = new pv_model();
    m->cfg = *cfg;
    m->device = device;
    int s = build_plan(m);
    if (s) { delete m; return s; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        delete m;
        return set_error(PV_ERR_NO_DEVICE, "no CUDA device: libprobav_b200 has no CPU fallback");
    }
    cudaDeviceProp prop;
    if (device < 0 || device >= ndev || cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        delete m;
        return set_error(PV_ERR_NO_DEVICE, "device %d is not an sm_100 (B200) GPU: this library is built for sm_100a only", device);
    }
    auto fail = [&](int code) { pv_model_destroy(m); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return fail(set_error(PV_ERR_CUDA, "cudaSetDevice(%d) failed", device));
    if (cudaMalloc(&m->params, m->nparams * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&m->weff, m->nweff * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&m->weffT, m->nweff * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&m->bias_s, m->nbias_s * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&m->scale, m->nscale * sizeof(float)) != cudaSuccess)
        return fail(set_error(PV_ERR_CUDA, "cudaMalloc of the weight arenas failed"));
    cudaMemset(m->params, 0, m->nparams * sizeof(float));
    cudaMemset(m->weff, 0, m->nweff * sizeof(float));
    cudaMemset(m->weffT, 0, m->nweff * sizeof(float));
    cudaMemset(m->bias_s, 0, m->nbias_s * sizeof(float));
    if (m->x3) {
        if (cudaMalloc(&m->weff_lo, m->nweff * sizeof(float)) != cudaSuccess || cudaMalloc(&m->weffT_lo, m->nweff * sizeof(float)) != cudaSuccess ||
            cudaMalloc(&m->weffT_pack, m->nweff * sizeof(float)) != cudaSuccess ||
            cudaMalloc(&m->weff_pack, m->nweff * sizeof(float)) != cudaSuccess)
            return fail(set_error(PV_ERR_CUDA, "cudaMalloc of the weight remainder arenas failed"));
        cudaMemset(m->weff_lo, 0, m->nweff * sizeof(float));
        cudaMemset(m->weffT_lo, 0, m->nweff * sizeof(float));
        cudaMemset(m->weffT_pack, 0, m->nweff * sizeof(float));
        cudaMemset(m->weff_pack, 0, m->nweff * sizeof(float));
    } else if (m->use_tc) {             // single-pass tensor-core engine: fp16 copies of the 3x3x3 weights for the inference convs (conv3_tc.cu MODE 3)
        if (cudaMalloc(&m->weffT_pack, m->nweff * sizeof(float)) != cudaSuccess)
            return fail(set_error(PV_ERR_CUDA, "cudaMalloc of the fp16 weight arena failed"));
        cudaMemset(m->weffT_pack, 0, m->nweff * sizeof(float));
    }
    std::vector<WnLayer> tab(m->layers.size());
    int blocks = 0;
    for (size_t i = 0; i < m->layers.size(); ++i) {
        const Layer& L = m->layers[i];
        WnLayer& w = tab[i];
        w.v_off = L.v_off; w.g_off = L.g_off; w.b_off = L.b_off;
        w.weff_off = L.weff_off; w.weffT_off = L.weff_off; w.bias_s_off = L.bias_s_off; w.scale_off = L.scale_off;
        w.taps = L.taps(); w.cin = L.cin; w.cout = L.cout; w.cin_s = L.cin_s; w.cout_s = L.cout_s;
        w.first_block = blocks;
        m->wn_first.push_back(blocks);
        w.mode = L.wn_mode;
        w.round_tf32 = (m->use_tc && L.wn_mode == 1 && L.cin > 1) ? 1 : 0;
        blocks += L.cout;
    }
    m->wn_blocks = blocks;
    m->wn_first.push_back(blocks);
    if (cudaMalloc(&m->wn_tab, tab.size() * sizeof(WnLayer)) != cudaSuccess ||
        cudaMemcpy(m->wn_tab, tab.data(), tab.size() * sizeof(WnLayer), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(set_error(PV_ERR_CUDA, "weight-norm table upload failed"));
    *out = m;
    return 0;
}

void pv_model_destroy(pv_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    m->pool_infer.release();
    m->pool_train.release();
    cudaFree(m->params); cudaFree(m->weff); cudaFree(m->weffT); cudaFree(m->bias_s); cudaFree(m->scale); cudaFree(m->wn_tab);
    cudaFree(m->weff_lo); cudaFree(m->weffT_lo); cudaFree(m->weffT_pack); cudaFree(m->weff_pack);
    cudaFree(m->stage_lr); cudaFree(m->stage_sr); cudaFree(m->stage_scene);
    m->scene_pipe.release();
    delete m;
}

int pv_model_param_count(const pv_model* m) { return m ? (int)m->pinfo.size() : set_error(PV_ERR_BAD_ARG, "null model"); }

int pv_model_param_info(const pv_model* m, int idx, char* name, int name_cap, int* rank, int64_t shape[5],
                        int64_t* offset, int64_t* numel) {
    if (!m || idx < 0 || idx >= (int)m->pinfo.size()) return set_error(PV_ERR_BAD_ARG, "param index %d out of range", idx);
    const PInfo& p = m->pinfo[idx];
    if (name && name_cap > 0) { std::strncpy(name, p.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (rank) *rank = p.rank;
    if (shape) for (int i = 0; i < 5; ++i) shape[i] = p.shape[i];
    if (offset) *offset = p.off;
    if (numel) *numel = p.numel;
    return 0;
}

int64_t pv_model_param_numel(const pv_model* m) { return m ? m->nparams : 0; }

int pv_model_set_params(pv_model* m, const float* flat_host, int64_t n) {
    if (!m || !flat_host || n != m->nparams) return set_error(PV_ERR_BAD_ARG, "set_params: expected %lld floats", m ? m->nparams : 0LL);
    PV_CUDA(cudaSetDevice(m->device));
    PV_CUDA(cudaMemcpy(m->params, flat_host, n * sizeof(float), cudaMemcpyHostToDevice));
    m->weff_dirty = true;
    return 0;
}

int pv_model_get_params(pv_model* m, float* flat_host, int64_t n) {
    if (!m || !flat_host || n != m->nparams) return set_error(PV_ERR_BAD_ARG, "get_params: expected %lld floats", m ? m->nparams : 0LL);
    PV_CUDA(cudaSetDevice(m->device));
    PV_CUDA(cudaDeviceSynchronize());
    PV_CUDA(cudaMemcpy(flat_host, m->params, n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

int pv_model_param_arena(pv_model* m, float** dev_ptr, int64_t* n) {
    if (!m || !dev_ptr || !n) return set_error(PV_ERR_BAD_ARG, "param_arena: null argument");
    *dev_ptr = m->params; *n = m->nparams;
    m->weff_dirty = true;       // the caller may write through the pointer (DP broadcast)
    return 0;
}

int pv_model_init_g_from_v(pv_model* m) {
    if (!m) return set_error(PV_ERR_BAD_ARG, "null model");
    PV_CUDA(cudaSetDevice(m->device));
    PV_TRY(launch_g_from_v(m->wn_tab, (int)m->layers.size(), m->wn_blocks, m->params, 0));
    PV_CUDA(cudaDeviceSynchronize());
    m->weff_dirty = true;
    return 0;
}

int pv_forward(pv_model* m, const float* lr_dev, int B, float* sr_dev, void* stream) {
    if (!m) return set_error(PV_ERR_BAD_ARG, "null model");
    return model_forward(m, lr_dev, B, sr_dev, false, 0, S_(stream));
}

int pv_resolve(pv_model* m, const float* lr_dev, int B, float* sr_dev, void* stream) {
    if (!m) return set_error(PV_ERR_BAD_ARG, "null model");
    return model_forward(m, lr_dev, B, sr_dev, false, 1, S_(stream));
}

static int stage_ensure(float** p, size_t* cap, size_t n) {
    if (n <= *cap) return 0;
    cudaFree(*p); *p = nullptr; *cap = 0;
    PV_CUDA(cudaMalloc(p, n * sizeof(float)));
    *cap = n;
    return 0;
}

static int forward_host(pv_model* m, const float* lr_host, int B, float* sr_host, int clip_round) {
    if (!m || !lr_host || !sr_host || B <= 0) return set_error(PV_ERR_BAD_ARG, "forward_host: null buffer or B=%d", B);
    PV_CUDA(cudaSetDevice(m->device));
    const size_t nin = (size_t)m->S * m->S * m->T, nout = (size_t)m->P * m->cfg.scale * m->P * m->cfg.scale;
    const int chunk = 256;
    PV_TRY(stage_ensure(&m->stage_lr, &m->stage_lr_n, nin * std::min(B, chunk)));
    PV_TRY(stage_ensure(&m->stage_sr, &m->stage_sr_n, nout * std::min(B, chunk)));
    for (int s = 0; s < B; s += chunk) {
        const int b = std::min(chunk, B - s);
        PV_CUDA(cudaMemcpyAsync(m->stage_lr, lr_host + nin * s, nin * b * sizeof(float), cudaMemcpyHostToDevice, 0));
        PV_TRY(model_forward(m, m->stage_lr, b, m->stage_sr, false, clip_round, 0));
        PV_CUDA(cudaMemcpyAsync(sr_host + nout * s, m->stage_sr, nout * b * sizeof(float), cudaMemcpyDeviceToHost, 0));
    }
    PV_CUDA(cudaStreamSynchronize(0));
    return 0;
}

int pv_forward_host(pv_model* m, const float* lr_host, int B, float* sr_host) { return forward_host(m, lr_host, B, sr_host, 0); }
int pv_resolve_host(pv_model* m, const float* lr_host, int B, float* sr_host) { return forward_host(m, lr_host, B, sr_host, 1); }

int pv_predict_scenes_host(pv_model* m, const float* lr_patches_host, int nscenes, int pps, float* sr_scenes_host) {
    if (!m || !lr_patches_host || !sr_scenes_host || nscenes <= 0 || pps <= 0) return set_error(PV_ERR_BAD_ARG, "predict_scenes: bad argument");
    const int n = (int)std::lround(std::sqrt((double)pps));
    if (n * n != pps) return set_error(PV_ERR_BAD_ARG, "predict_scenes: %d patches per scene is not a square (test.py:152)", pps);
    PV_CUDA(cudaSetDevice(m->device));
    const size_t nin = (size_t)m->S * m->S * m->T, PP = (size_t)m->P * m->cfg.scale, nout = PP * PP;
    const int chunk = std::max(1, 256 / pps);          // scenes per launch
    PV_TRY(stage_ensure(&m->stage_lr, &m->stage_lr_n, nin * pps * chunk));
    PV_TRY(stage_ensure(&m->stage_sr, &m->stage_sr_n, nout * pps * chunk));
    PV_TRY(stage_ensure(&m->stage_scene, &m->stage_scene_n, nout * pps * chunk));
    for (int s = 0; s < nscenes; s += chunk) {
        const int ns = std::min(chunk, nscenes - s);
        PV_CUDA(cudaMemcpyAsync(m->stage_lr, lr_patches_host + nin * pps * s, nin * pps * ns * sizeof(float), cudaMemcpyHostToDevice, 0));
        PV_TRY(model_forward(m, m->stage_lr, ns * pps, m->stage_sr, false, 1, 0));
        PV_TRY(launch_stitch(m->stage_sr, ns, n, (int)PP, m->stage_scene, 0));
        PV_CUDA(cudaMemcpyAsync(sr_scenes_host + nout * pps * s, m->stage_scene, nout * pps * ns * sizeof(float), cudaMemcpyDeviceToHost, 0));
    }
    PV_CUDA(cudaStreamSynchronize(0));
    return 0;
}

// device-resident scenes in, device-resident stitched scenes out; chunks of SCENE_CHUNK_PATCHES patches per forward pass
static const int SCENE_CHUNK_PATCHES = 512;

static int predict_chunk(pv_model* m, const float* scn_dev, int ns, int H, int W, float* out_dev, cudaStream_t st) {
    const int n = H / m->P;
    const size_t PP = (size_t)m->P * m->cfg.scale;
    PV_TRY(launch_scene_to_patches(scn_dev, ns, m->T, H, W, m->P, m->cfg.max_shift, m->stage_lr, st));
    PV_TRY(model_forward(m, m->stage_lr, ns * n * n, m->stage_sr, false, 1, st));
    return launch_stitch(m->stage_sr, ns, n, (int)PP, out_dev, st);
}

static int scene_args_ok(pv_model* m, const void* a, const void* b, int nscenes, int H, int W) {
    if (!m || !a || !b || nscenes <= 0) return set_error(PV_ERR_BAD_ARG, "predict_from_scenes: bad argument");
    if (H != W || H <= 0 || H % m->P) return set_error(PV_ERR_BAD_ARG, "predict_from_scenes: scene %dx%d must be square and a multiple of patch_size", H, W);
    return 0;
}

int pv_predict_from_scenes(pv_model* m, const float* lr_scenes_dev, int nscenes, int H, int W, float* sr_scenes_dev, void* stream) {
    PV_TRY(scene_args_ok(m, lr_scenes_dev, sr_scenes_dev, nscenes, H, W));
    PV_CUDA(cudaSetDevice(m->device));
    const int n = H / m->P, pps = n * n;
    const size_t nin = (size_t)m->S * m->S * m->T, PP = (size_t)m->P * m->cfg.scale, nout = PP * PP, nsc = (size_t)m->T * H * W;
    const int chunk = std::max(1, SCENE_CHUNK_PATCHES / pps);
    PV_TRY(stage_ensure(&m->stage_lr, &m->stage_lr_n, nin * pps * chunk));
    PV_TRY(stage_ensure(&m->stage_sr, &m->stage_sr_n, nout * pps * chunk));
    for (int s = 0; s < nscenes; s += chunk)
        PV_TRY(predict_chunk(m, lr_scenes_dev + nsc * s, std::min(chunk, nscenes - s), H, W, sr_scenes_dev + nout * pps * s, S_(stream)));
    return 0;
}

// Host scenes in, host scenes out.  Two slots of pinned staging + device buffers: while chunk c computes, chunk c+1 is copied
// into pinned memory and up to the device on a copy-in stream, and chunk c-1 comes back on a copy-out stream and is copied to
// the caller's (pageable) array -- the PCIe transfers and both host memcpys overlap the forward pass.
int pv_predict_from_scenes_host(pv_model* m, const float* lr_scenes_host, int nscenes, int H, int W, float* sr_scenes_host) {
    PV_TRY(scene_args_ok(m, lr_scenes_host, sr_scenes_host, nscenes, H, W));
    PV_CUDA(cudaSetDevice(m->device));
    PV_CUDA(cudaDeviceSynchronize());     // the private non-blocking streams below do not order against work already in flight
    const int n = H / m->P, pps = n * n;
    const size_t nin = (size_t)m->S * m->S * m->T, PP = (size_t)m->P * m->cfg.scale, nout = PP * PP, nsc = (size_t)m->T * H * W;
    const int chunk = std::max(1, SCENE_CHUNK_PATCHES / pps);
    pv::ScenePipe& sp = m->scene_pipe;
    PV_TRY(sp.ensure(nsc * chunk, nout * pps * chunk));
    PV_TRY(stage_ensure(&m->stage_lr, &m->stage_lr_n, nin * pps * chunk));
    PV_TRY(stage_ensure(&m->stage_sr, &m->stage_sr_n, nout * pps * chunk));
    const int nchunks = (nscenes + chunk - 1) / chunk;
    int rc = 0;
    auto drain = [&](int c) {            // chunk c's result: wait for its D2H, then copy pinned -> caller
        const int slot = c & 1, s0 = c * chunk, ns = std::min(chunk, nscenes - s0);
        if (cudaEventSynchronize(sp.ev_out[slot]) != cudaSuccess) return set_error(PV_ERR_CUDA, "predict_from_scenes: D2H failed");
        std::memcpy(sr_scenes_host + nout * pps * s0, sp.pin_out[slot], nout * pps * ns * sizeof(float));
        return 0;
    };
    for (int c = 0; c < nchunks && !rc; ++c) {
        const int slot = c & 1, s0 = c * chunk, ns = std::min(chunk, nscenes - s0);
        if (c >= 2 && cudaEventSynchronize(sp.ev_in[slot]) != cudaSuccess) { rc = set_error(PV_ERR_CUDA, "predict_from_scenes: H2D failed"); break; }
        std::memcpy(sp.pin_in[slot], lr_scenes_host + nsc * s0, nsc * ns * sizeof(float));
        if (c >= 2) cudaStreamWaitEvent(sp.st_in, sp.ev_patched[slot], 0);       // scn[slot] was consumed by chunk c-2's patching
        if (cudaMemcpyAsync(sp.scn[slot], sp.pin_in[slot], nsc * ns * sizeof(float), cudaMemcpyHostToDevice, sp.st_in) != cudaSuccess) {
            rc = set_error(PV_ERR_CUDA, "predict_from_scenes: H2D failed"); break;
        }
        cudaEventRecord(sp.ev_in[slot], sp.st_in);
        cudaStreamWaitEvent(sp.st_c, sp.ev_in[slot], 0);
        if (c >= 2) cudaStreamWaitEvent(sp.st_c, sp.ev_out[slot], 0);            // out[slot] has left the device (chunk c-2)
        if ((rc = launch_scene_to_patches(sp.scn[slot], ns, m->T, H, W, m->P, m->cfg.max_shift, m->stage_lr, sp.st_c))) break;
        cudaEventRecord(sp.ev_patched[slot], sp.st_c);
        if ((rc = model_forward(m, m->stage_lr, ns * pps, m->stage_sr, false, 1, sp.st_c))) break;
        if ((rc = launch_stitch(m->stage_sr, ns, n, (int)PP, sp.out[slot], sp.st_c))) break;
        cudaEventRecord(sp.ev_done[slot], sp.st_c);
        if (c >= 1) rc = drain(c - 1);                                           // overlaps chunk c's forward pass
        if (rc) break;
        cudaStreamWaitEvent(sp.st_out, sp.ev_done[slot], 0);
        if (cudaMemcpyAsync(sp.pin_out[slot], sp.out[slot], nout * pps * ns * sizeof(float), cudaMemcpyDeviceToHost, sp.st_out) != cudaSuccess) {
            rc = set_error(PV_ERR_CUDA, "predict_from_scenes: D2H failed"); break;
        }
        cudaEventRecord(sp.ev_out[slot], sp.st_out);
    }
    if (!rc) rc = drain(nchunks - 1);
    if (cudaStreamSynchronize(sp.st_c) != cudaSuccess && !rc) rc = set_error(PV_ERR_CUDA, "predict_from_scenes: %s", cudaGetErrorString(cudaGetLastError()));
    cudaStreamSynchronize(sp.st_in); cudaStreamSynchronize(sp.st_out);
    return rc;
}

// ------------------------------------------------------------------------------------------ losses
int pv_shift_loss(int kind, const float* hr, const uint8_t* mask, const float* sr, int B, int H, int W, int border,
                  float grad_scale, float* loss_ps, int32_t* best_shift, int32_t* clear_count, float* cpsnr_ps,
                  float* mean_loss, float* dsr, float* stack_out, void* stream) {
    return shift_loss_device(kind, hr, mask, sr, B, H, W, border, grad_scale, loss_ps, best_shift, clear_count, cpsnr_ps,
                             mean_loss, dsr, stack_out, S_(stream));
}

int pv_shift_loss_host(int kind, const float* hr, const uint8_t* mask, const float* sr, int B, int H, int W, int border,
                       float grad_scale, float* loss_ps, int32_t* best_shift, int32_t* clear_count, float* cpsnr_ps,
                       float* mean_loss, float* dsr, float* stack_out) {
    if (!hr || !mask || !sr || !loss_ps || !best_shift || !clear_count || B <= 0 || H <= 0 || W <= 0)
        return set_error(PV_ERR_BAD_ARG, "pv_shift_loss_host: null buffer or bad shape");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return set_error(PV_ERR_NO_DEVICE, "no CUDA device: no CPU fallback"); }
    const size_t n = (size_t)B * H * W;
    float *d_hr = nullptr, *d_sr = nullptr, *d_f = nullptr;
    uint8_t* d_m = nullptr;
    int32_t* d_i = nullptr;
    // sub-buffer offsets rounded to 4 floats: the kernel stores dSR as float4
    const size_t o_mean = ((size_t)B * 2 + 3) & ~(size_t)3, o_dsr = o_mean + 4, o_stack = o_dsr + (dsr ? n : 0);
    const size_t nf = o_stack + (stack_out ? (size_t)B * 49 * 4 : 0);
    int rc = 0;
    do {
        if (cudaMalloc(&d_hr, n * 4) != cudaSuccess || cudaMalloc(&d_sr, n * 4) != cudaSuccess || cudaMalloc(&d_m, n) != cudaSuccess ||
            cudaMalloc(&d_f, nf * 4) != cudaSuccess || cudaMalloc(&d_i, (size_t)B * 2 * 4) != cudaSuccess) { rc = set_error(PV_ERR_CUDA, "cudaMalloc failed"); break; }
        cudaMemcpyAsync(d_hr, hr, n * 4, cudaMemcpyHostToDevice, 0);
        cudaMemcpyAsync(d_sr, sr, n * 4, cudaMemcpyHostToDevice, 0);
        cudaMemcpyAsync(d_m, mask, n, cudaMemcpyHostToDevice, 0);
        float* d_loss = d_f; float* d_cp = d_f + B; float* d_mean = d_f + o_mean;
        float* d_dsr = dsr ? d_f + o_dsr : nullptr;
        float* d_stack = stack_out ? d_f + o_stack : nullptr;
        if ((rc = shift_loss_device(kind, d_hr, d_m, d_sr, B, H, W, border, grad_scale, d_loss, d_i, d_i + B, d_cp, d_mean, d_dsr, d_stack, 0))) break;
        cudaMemcpyAsync(loss_ps, d_loss, (size_t)B * 4, cudaMemcpyDeviceToHost, 0);
        cudaMemcpyAsync(best_shift, d_i, (size_t)B * 4, cudaMemcpyDeviceToHost, 0);
        cudaMemcpyAsync(clear_count, d_i + B, (size_t)B * 4, cudaMemcpyDeviceToHost, 0);
        if (cpsnr_ps) cudaMemcpyAsync(cpsnr_ps, d_cp, (size_t)B * 4, cudaMemcpyDeviceToHost, 0);
        if (mean_loss) cudaMemcpyAsync(mean_loss, d_mean, 4, cudaMemcpyDeviceToHost, 0);
        if (dsr) cudaMemcpyAsync(dsr, d_dsr, n * 4, cudaMemcpyDeviceToHost, 0);
        if (stack_out) cudaMemcpyAsync(stack_out, d_stack, (size_t)B * 49 * 16, cudaMemcpyDeviceToHost, 0);
        if (cudaStreamSynchronize(0) != cudaSuccess) rc = set_error(PV_ERR_CUDA, "shift loss failed: %s", cudaGetErrorString(cudaGetLastError()));
    } while (0);
    cudaFree(d_hr); cudaFree(d_sr); cudaFree(d_m); cudaFree(d_f); cudaFree(d_i);
    return rc;
}

// ------------------------------------------------------------------------------------------ trainer
int pv_trainer_create(pv_model* m, int opt_kind, float learning_rate, int loss_kind, pv_trainer** out) {
    if (!m || !out) return set_error(PV_ERR_BAD_ARG, "pv_trainer_create: null argument");
    *out = nullptr;
    if (opt_kind < PV_OPT_SGD || opt_kind > PV_OPT_NADAM) return set_error(PV_ERR_BAD_ARG, "unknown optimizer %d", opt_kind);
    if (loss_kind != PV_LOSS_L1 && loss_kind != PV_LOSS_L2 && loss_kind != PV_LOSS_L1EDGE)
        return set_error(PV_ERR_BAD_ARG, "unknown loss kind %d", loss_kind);
    if (m->P * m->cfg.scale != 48)
        return set_error(PV_ERR_BAD_CONFIG, "training needs scale*patch_size = 48 (fused loss backward); got %d", m->P * m->cfg.scale);
    PV_CUDA(cudaSetDevice(m->device));
    pv_trainer* t = new pv_trainer();
    t->m = m; t->opt = opt_kind; t->lr = learning_rate; t->loss_kind = loss_kind;
    if (cudaMalloc(&t->grads, m->nparams * 4) != cudaSuccess || cudaMalloc(&t->m1, m->nparams * 4) != cudaSuccess ||
        cudaMalloc(&t->m2, m->nparams * 4) != cudaSuccess || cudaMalloc(&t->dweff, m->nweff * 4) != cudaSuccess ||
        cudaMalloc(&t->dbias_s, m->nbias_s * 4) != cudaSuccess || cudaMalloc(&t->out2, 2 * 4) != cudaSuccess) {
        pv_trainer_destroy(t);
        return set_error(PV_ERR_CUDA, "cudaMalloc of the trainer arenas failed");
    }
    // split-reduction scratch of the CUDA-core weight-gradient kernels (the dense engine, and the 2-D skip path's fallback): makes the
    // exact engines bit-reproducible
    if (cudaMalloc(&t->dense_partials, WGRAD_PARTIAL_FLOATS * sizeof(float)) != cudaSuccess) {
        pv_trainer_destroy(t);
        return set_error(PV_ERR_CUDA, "cudaMalloc of the weight-gradient split scratch failed");
    }
    if (m->rows) {
        t->wg_partial_floats = (size_t)148 * (9 * 4096 + 1024);
        // deferred reductions: one private region per conv3 layer (R norm convs + reducers + upscale), per fused block,
        // for mainConv1 and for the 2-D skip path (sized for batches up to 4096; larger batches fall back to immediate mode)
        const size_t conv3_layers = (size_t)m->R + m->nred + 1;
        const size_t arena = t->wg_partial_floats + conv3_layers * ((size_t)148 * (9 * 4096 + 128) + 64) +
                             (size_t)m->R * ((size_t)148 * (4 * 4096 + pv::RESBW_DBP) + 64) + (size_t)148 * 4 * 28 * 32 + 64 +
                             pv::skip2d_partial_floats(4096, m->S, m->cfg.scale * m->cfg.scale) + 64;
        t->rq.arena_floats = arena;
        if (cudaMalloc(&t->wg_partials, arena * 4) != cudaSuccess) {
            pv_trainer_destroy(t);
            return set_error(PV_ERR_CUDA, "cudaMalloc of the weight-gradient partials failed");
        }
    }
    t->rq.arena = t->wg_partials;
    cudaMemset(t->grads, 0, m->nparams * 4);
    cudaMemset(t->m1, 0, m->nparams * 4);
    cudaMemset(t->m2, 0, m->nparams * 4);
    // legacy-stream memsets are not ordered against non-blocking streams: make them complete before the handle is used
    if (cudaDeviceSynchronize() != cudaSuccess) {
        pv_trainer_destroy(t);
        return set_error(PV_ERR_CUDA, "pv_trainer_create: device synchronisation failed");
    }
    *out = t;
    return 0;
}

void pv_trainer_destroy(pv_trainer* t) {
    if (!t) return;
    cudaSetDevice(t->m->device);
    cudaFree(t->grads); cudaFree(t->dweff); cudaFree(t->dbias_s); cudaFree(t->m1); cudaFree(t->m2);
    cudaFree(t->sr); cudaFree(t->dsr); cudaFree(t->loss_ps); cudaFree(t->cpsnr_ps); cudaFree(t->best); cudaFree(t->cnt);
    cudaFree(t->out2); cudaFree(t->s_lr); cudaFree(t->s_hr); cudaFree(t->s_mask); cudaFree(t->wg_partials); cudaFree(t->dense_partials);
    delete t;
}

int pv_train_forward_backward(pv_trainer* t, const float* lr, const float* hr, const uint8_t* mask, int B, float grad_scale,
                              float* out_dev, void* stream) {
    if (!t) return set_error(PV_ERR_BAD_ARG, "null trainer");
    if (B == 0) {       // empty data-parallel shard: zero gradients (see pv_train_forward_backward_staged)
        PV_CUDA(cudaSetDevice(t->m->device));
        PV_CUDA(cudaMemsetAsync(t->grads, 0, (size_t)t->m->nparams * sizeof(float), S_(stream)));
        if (out_dev) PV_CUDA(cudaMemsetAsync(out_dev, 0, 2 * sizeof(float), S_(stream)));
        return 0;
    }
    return trainer_fwd_loss(t, lr, hr, mask, B, grad_scale, true, out_dev, S_(stream));
}

int pv_train_forward_backward_staged(pv_trainer* t, const float* lr, const float* hr, const uint8_t* mask, int B, float grad_scale,
                                     float* out_dev, int stage, int64_t* grad_lo, int64_t* grad_hi, void* stream) {
    if (!t || !grad_lo || !grad_hi) return set_error(PV_ERR_BAD_ARG, "pv_train_forward_backward_staged: null argument");
    if (stage != 0 && stage != 1) return set_error(PV_ERR_BAD_ARG, "pv_train_forward_backward_staged: stage must be 0 or 1, got %d", stage);
    pv_model* m = t->m;
    // the parameter arena is laid out in layer order, so each bucket is one contiguous range of the gradient arena
    const int split = m->rows ? pv::tc_bucket_split_layer(m) : 0;
    const int64_t cut = split < (int)m->layers.size() ? m->layers[split].v_off : m->nparams;
    if (B == 0) {
        // A rank whose shard of the last partial global batch is empty (fewer samples than ranks) contributes zero gradients
        // and still joins both all-reduces (trainClass._dp_step): the reference's single process keeps the partial batch
        // (trainClass.py:84-93), so the data-parallel job must not dead-lock on it.
        PV_CUDA(cudaSetDevice(m->device));
        *grad_lo = stage == 0 ? cut : 0; *grad_hi = stage == 0 ? m->nparams : cut;
        PV_CUDA(cudaMemsetAsync(t->grads + *grad_lo, 0, (size_t)(*grad_hi - *grad_lo) * sizeof(float), S_(stream)));
        if (stage == 0 && out_dev) PV_CUDA(cudaMemsetAsync(out_dev, 0, 2 * sizeof(float), S_(stream)));
        return 0;
    }
    if (stage == 0) {
        PV_TRY(trainer_fwd_loss(t, lr, hr, mask, B, grad_scale, true, out_dev, S_(stream), 0));
        *grad_lo = cut; *grad_hi = m->nparams;
    } else {
        PV_CUDA(cudaSetDevice(m->device));
        PV_TRY(model_backward(t, t->dsr, B, S_(stream), 1));
        *grad_lo = 0; *grad_hi = cut;
    }
    return 0;
}

int pv_trainer_forward(pv_trainer* t, const float* lr, int B, float* sr, void* stream) {
    if (!t) return set_error(PV_ERR_BAD_ARG, "null trainer");
    t->fwd_B = 0;
    PV_TRY(model_forward(t->m, lr, B, sr, true, 0, S_(stream)));
    t->fwd_B = B;
    return 0;
}

int pv_trainer_backward(pv_trainer* t, const float* dsr, int B, void* stream) {
    if (!t || !dsr) return set_error(PV_ERR_BAD_ARG, "pv_trainer_backward: null argument");
    if (B <= 0 || B != t->fwd_B) return set_error(PV_ERR_STATE, "pv_trainer_backward: no training-mode forward of batch %d to differentiate (last: %d)", B, t->fwd_B);
    PV_CUDA(cudaSetDevice(t->m->device));
    return model_backward(t, dsr, B, S_(stream));
}

int pv_apply_gradients(pv_trainer* t, void* stream) {
    if (!t) return set_error(PV_ERR_BAD_ARG, "null trainer");
    PV_CUDA(cudaSetDevice(t->m->device));
    return trainer_apply(t, S_(stream));
}

int pv_train_step(pv_trainer* t, const float* lr, const float* hr, const uint8_t* mask, int B, float* out_dev, void* stream) {
    if (!t) return set_error(PV_ERR_BAD_ARG, "null trainer");
    PV_TRY(trainer_fwd_loss(t, lr, hr, mask, B, 1.0f / (float)B, true, out_dev, S_(stream)));
    return trainer_apply(t, S_(stream));
}

int pv_eval_step(pv_trainer* t, const float* lr, const float* hr, const uint8_t* mask, int B, float* out_dev, void* stream) {
    if (!t) return set_error(PV_ERR_BAD_ARG, "null trainer");
    return trainer_fwd_loss(t, lr, hr, mask, B, 0.f, false, out_dev, S_(stream));
}

static int step_host(pv_trainer* t, const float* lr, const float* hr, const uint8_t* mask, int B, float* out_host, bool train) {
    if (!t || !lr || !hr || !mask || !out_host || B <= 0) return set_error(PV_ERR_BAD_ARG, "step_host: null buffer or B=%d", B);
    pv_model* m = t->m;
    PV_CUDA(cudaSetDevice(m->device));
    const size_t nin = (size_t)m->S * m->S * m->T, nhr = (size_t)m->P * m->cfg.scale * m->P * m->cfg.scale;
    if (B > t->s_cap) {
        cudaFree(t->s_lr); cudaFree(t->s_hr); cudaFree(t->s_mask);
        t->s_lr = t->s_hr = nullptr; t->s_mask = nullptr; t->s_cap = 0;
        PV_CUDA(cudaMalloc(&t->s_lr, nin * B * 4));
        PV_CUDA(cudaMalloc(&t->s_hr, nhr * B * 4));
        PV_CUDA(cudaMalloc(&t->s_mask, nhr * B));
        t->s_cap = B;
    }
    PV_CUDA(cudaMemcpyAsync(t->s_lr, lr, nin * B * 4, cudaMemcpyHostToDevice, 0));
    PV_CUDA(cudaMemcpyAsync(t->s_hr, hr, nhr * B * 4, cudaMemcpyHostToDevice, 0));
    PV_CUDA(cudaMemcpyAsync(t->s_mask, mask, nhr * B, cudaMemcpyHostToDevice, 0));
    if (train) PV_TRY(pv_train_step(t, t->s_lr, t->s_hr, t->s_mask, B, t->out2, nullptr));
    else PV_TRY(pv_eval_step(t, t->s_lr, t->s_hr, t->s_mask, B, t->out2, nullptr));
    PV_CUDA(cudaMemcpyAsync(out_host, t->out2, 2 * 4, cudaMemcpyDeviceToHost, 0));
    PV_CUDA(cudaStreamSynchronize(0));
    return 0;
}

int pv_train_step_host(pv_trainer* t, const float* lr, const float* hr, const uint8_t* mask, int B, float* out_host) {
    return step_host(t, lr, hr, mask, B, out_host, true);
}
int pv_eval_step_host(pv_trainer* t, const float* lr, const float* hr, const uint8_t* mask, int B, float* out_host) {
    return step_host(t, lr, hr, mask, B, out_host, false);
}

int pv_trainer_grad_arena(pv_trainer* t, float** dev_ptr, int64_t* n) {
    if (!t || !dev_ptr || !n) return set_error(PV_ERR_BAD_ARG, "grad_arena: null argument");
    *dev_ptr = t->grads; *n = t->m->nparams;
    return 0;
}

int pv_trainer_get_state(pv_trainer* t, int64_t* iter, double* momentum_cache, float* m_host, float* v_host, int64_t n) {
    if (!t) return set_error(PV_ERR_BAD_ARG, "null trainer");
    PV_CUDA(cudaSetDevice(t->m->device));
    PV_CUDA(cudaDeviceSynchronize());
    if (iter) *iter = t->iter;
    if (momentum_cache) *momentum_cache = t->momentum_cache;
    if (m_host || v_host) {
        if (n != t->m->nparams) return set_error(PV_ERR_BAD_ARG, "get_state: expected %lld floats", t->m->nparams);
        if (m_host) PV_CUDA(cudaMemcpy(m_host, t->m1, n * 4, cudaMemcpyDeviceToHost));
        if (v_host) PV_CUDA(cudaMemcpy(v_host, t->m2, n * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int pv_trainer_set_state(pv_trainer* t, int64_t iter, double momentum_cache, const float* m_host, const float* v_host, int64_t n) {
    if (!t) return set_error(PV_ERR_BAD_ARG, "null trainer");
    PV_CUDA(cudaSetDevice(t->m->device));
    t->iter = iter; t->momentum_cache = momentum_cache;
    if (m_host || v_host) {
        if (n != t->m->nparams) return set_error(PV_ERR_BAD_ARG, "set_state: expected %lld floats", t->m->nparams);
        if (m_host) PV_CUDA(cudaMemcpy(t->m1, m_host, n * 4, cudaMemcpyHostToDevice));
        if (v_host) PV_CUDA(cudaMemcpy(t->m2, v_host, n * 4, cudaMemcpyHostToDevice));
    }
    return 0;
}

int pv_trainer_set_lr(pv_trainer* t, float learning_rate) {
    if (!t) return set_error(PV_ERR_BAD_ARG, "null trainer");
    t->lr = learning_rate;
    return 0;
}

}  // extern "C"
