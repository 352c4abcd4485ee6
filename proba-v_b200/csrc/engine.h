// engine.h -- handle structs shared by the dense fp32 engine (engine.cu) and the row-layout tensor-core engine (engine_tc.cu).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "kernels.h"
#include "rows.h"
#include "wgrad_reduce.cuh"

namespace pv {

static inline int storage_channels(int c) { return c > 16 ? ((c + 31) / 32) * 32 : c; }

struct Layer {
    std::string name;
    int k[3];            // kernel (H, W, T)
    int cin, cout, cin_s, cout_s;
    int pad[3];          // zero padding per side ('same' = k/2, 'valid' = 0)
    int relu;
    int Hi, Wi, Ti, Ho, Wo, To;
    long long v_off, g_off, b_off, weff_off, bias_s_off, scale_off;
    int wn_mode = 0;     // WnLayer::mode
    int taps() const { return k[0] * k[1] * k[2]; }
};

struct PInfo { std::string name; int rank; int64_t shape[5]; int64_t off, numel; };

struct Pool {
    std::map<std::string, float*> ptr;
    struct Spec { std::string name; size_t per, extra; };
    std::vector<Spec> spec;                             // floats per sample + fixed extra floats (row-layout lead/tail)
    int cap = 0;
    void add(const std::string& n, size_t per, size_t extra = 0) {
        for (auto& s : spec) if (s.name == n) { s.per = std::max(s.per, per); s.extra = std::max(s.extra, extra); return; }
        spec.push_back({n, per, extra});
    }
    // `st`: the stream the buffers' first kernels will be launched on.  The zero-fill must be ordered before them, and a
    // cudaMemset on the legacy stream is NOT ordered against cudaStreamNonBlocking streams (scene pipeline, torch side
    // streams), so it is issued on `st` itself.
    int ensure(int B, cudaStream_t st) {
        if (B <= cap) return 0;
        release();
        for (auto& s : spec) {
            float* d = nullptr;
            const size_t bytes = (s.per * (size_t)B + s.extra) * sizeof(float);
            PV_CUDA(cudaMalloc(&d, bytes));
            PV_CUDA(cudaMemsetAsync(d, 0, bytes, st));  // row layouts rely on never-written padding rows being zero
            ptr[s.name] = d;
        }
        cap = B;
        return 0;
    }
    void release() {
        for (auto& kv : ptr) cudaFree(kv.second);
        ptr.clear();
        cap = 0;
    }
    float* operator[](const std::string& n) {
        auto it = ptr.find(n);
        return it == ptr.end() ? nullptr : it->second;
    }
};

// double-buffered staging of the host-to-host scene prediction path (pv_predict_from_scenes_host)
struct ScenePipe {
    float *pin_in[2] = {nullptr, nullptr}, *pin_out[2] = {nullptr, nullptr};   // pinned host
    float *scn[2] = {nullptr, nullptr}, *out[2] = {nullptr, nullptr};           // device
    size_t cap_in = 0, cap_out = 0;
    cudaStream_t st_in = nullptr, st_c = nullptr, st_out = nullptr;
    cudaEvent_t ev_in[2] = {}, ev_patched[2] = {}, ev_done[2] = {}, ev_out[2] = {};
    bool made = false;
    int ensure(size_t n_in, size_t n_out) {
        if (!made) {
            PV_CUDA(cudaStreamCreateWithFlags(&st_in, cudaStreamNonBlocking));
            PV_CUDA(cudaStreamCreateWithFlags(&st_c, cudaStreamNonBlocking));
            PV_CUDA(cudaStreamCreateWithFlags(&st_out, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) {
                PV_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
                PV_CUDA(cudaEventCreateWithFlags(&ev_patched[i], cudaEventDisableTiming));
                PV_CUDA(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
                PV_CUDA(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
            }
            made = true;
        }
        if (n_in > cap_in) {
            for (int i = 0; i < 2; ++i) {
                cudaFreeHost(pin_in[i]); cudaFree(scn[i]); pin_in[i] = scn[i] = nullptr;
                PV_CUDA(cudaHostAlloc(&pin_in[i], n_in * sizeof(float), cudaHostAllocDefault));
                PV_CUDA(cudaMalloc(&scn[i], n_in * sizeof(float)));
            }
            cap_in = n_in;
        }
        if (n_out > cap_out) {
            for (int i = 0; i < 2; ++i) {
                cudaFreeHost(pin_out[i]); cudaFree(out[i]); pin_out[i] = out[i] = nullptr;
                PV_CUDA(cudaHostAlloc(&pin_out[i], n_out * sizeof(float), cudaHostAllocDefault));
                PV_CUDA(cudaMalloc(&out[i], n_out * sizeof(float)));
            }
            cap_out = n_out;
        }
        return 0;
    }
    void release() {
        for (int i = 0; i < 2; ++i) { cudaFreeHost(pin_in[i]); cudaFreeHost(pin_out[i]); cudaFree(scn[i]); cudaFree(out[i]); }
        if (made) {
            cudaStreamDestroy(st_in); cudaStreamDestroy(st_c); cudaStreamDestroy(st_out);
            for (int i = 0; i < 2; ++i) { cudaEventDestroy(ev_in[i]); cudaEventDestroy(ev_patched[i]); cudaEventDestroy(ev_done[i]); cudaEventDestroy(ev_out[i]); }
        }
        made = false; cap_in = cap_out = 0;
    }
};

}  // namespace pv

using pv::Layer;
using pv::PInfo;
using pv::Pool;
using pv::WnLayer;

struct pv_model {
    pv_cfg cfg;
    int device = 0;
    int S = 0, T = 0, P = 0, F = 0, R = 0, nred = 0;
    std::vector<Layer> layers;
    std::vector<PInfo> pinfo;
    std::vector<int> red_pad;          // 3 ints per reducer: reflect pad before it
    long long nparams = 0, nweff = 0, nbias_s = 0, nscale = 0;
    float *params = nullptr, *weff = nullptr, *weffT = nullptr, *bias_s = nullptr, *scale = nullptr;
    WnLayer* wn_tab = nullptr;
    int wn_blocks = 0;
    std::vector<int> wn_first;         // first wn block of layer i (prefix sum of cout), size layers + 1
    bool weff_dirty = true;
    Pool pool_infer, pool_train;
    bool rows = false;                 // row-layout engine (cfg.precision != 0), engine_tc.cu
    bool use_tc = false;               // tcgen05 kernels (precision 1, 4) vs the CUDA-core row kernels (precision 3)
    bool x3 = false;                   // precision 4: error-compensated tensor-core engine.  Every forward product is
                                       // x_hi w_hi + x_lo w_hi + x_hi w_lo (hi = tf32(v), lo = v - hi), data gradients use w_hi + w_lo
    float *weff_lo = nullptr, *weffT_lo = nullptr;   // w - tf32(w) in the layouts of weff / weffT
    float* weff_pack = nullptr;                      // bf16 pair rows [w_a | w - w_a] indexed like weff (data gradient: rows (tap, ci), K = co)
    float* weffT_pack = nullptr;                     // packed fp16 pair rows of the 3x3x3 layers (rows.h), indexed like weffT
    float *stage_lr = nullptr, *stage_sr = nullptr, *stage_scene = nullptr;   // host-API staging
    size_t stage_lr_n = 0, stage_sr_n = 0, stage_scene_n = 0;
    pv::ScenePipe scene_pipe;

    int li(const std::string& n) const {
        for (size_t i = 0; i < layers.size(); ++i) if (layers[i].name == n) return (int)i;
        return -1;
    }
    std::string A(int i, bool tr) const { return tr ? "a" + std::to_string(i) : "a" + std::to_string(i & 1); }
    std::string E(int i, bool tr) const { return tr ? "E" + std::to_string(i) : "E"; }
    std::string D(int i, bool tr) const { return tr ? "D" + std::to_string(i) : "D"; }
};

struct pv_trainer {
    pv_model* m = nullptr;
    int opt = PV_OPT_NADAM, loss_kind = PV_LOSS_L1;
    float lr = 1e-3f;
    long long iter = 0;
    int fwd_B = 0;                     // batch of the last pv_trainer_forward (what pv_trainer_backward may differentiate)
    double momentum_cache = 1.0;
    float *grads = nullptr, *dweff = nullptr, *dbias_s = nullptr, *m1 = nullptr, *m2 = nullptr;
    // per-batch loss workspace
    int capB = 0;
    float *sr = nullptr, *dsr = nullptr, *loss_ps = nullptr, *cpsnr_ps = nullptr, *out2 = nullptr;
    int32_t *best = nullptr, *cnt = nullptr;
    // host-API staging
    float *s_lr = nullptr, *s_hr = nullptr;
    uint8_t* s_mask = nullptr;
    int s_cap = 0;
    float* dense_partials = nullptr;   // split scratch of the CUDA-core wgrad kernels (kernels.h WGRAD_PARTIAL_FLOATS): fixed-order reduction
    float* wg_partials = nullptr;      // per-CTA partial weight gradients of the tensor-core wgrad kernels: the first
    size_t wg_partial_floats = 0;      // wg_partial_floats are the immediate-mode scratch, the rest of the arena is carved per layer
    pv::ReduceQueue rq;                // by the deferred reduction queue (one reduction launch per backward pass)
};


namespace pv {
// row-layout (tensor-core) engine, engine_tc.cu
int tc_build_plan(pv_model* m);
int tc_forward(pv_model* m, const float* lr, int B, float* sr, bool train, int clip_round, cudaStream_t st);
// stage -1: the whole backward pass; 0 / 1: its two gradient buckets (data-parallel overlap, pv_train_forward_backward_staged):
// stage 0 = tail, skip path, reducers and residual blocks R-1 .. R/2; stage 1 = blocks R/2-1 .. 0 and mainConv1
int tc_backward(pv_trainer* t, const float* g_sr, int B, cudaStream_t st, int stage = -1);
int tc_bucket_split_layer(const pv_model* m);      // first layer (index) whose gradients are final after stage 0
int refresh_weights(pv_model* m, cudaStream_t st);
int tc_selftest(std::string& report);
// dense-layout conv helpers of engine.cu (the 2-D low-frequency path uses them in both engines)
int conv_fwd(pv_model* m, int li, const float* in, float* out, const float* res, int B, cudaStream_t st);
int conv_dgrad(pv_model* m, int li, const float* gout, const float* relu_ref, float* gin, const float* res, int B, cudaStream_t st);
int conv_wgrad(pv_trainer* t, int li, const float* in, const float* gout, const float* relu_ref, int B, cudaStream_t st);
}  // namespace pv
