/* probav_b200.h -- C-ABI of libprobav_b200.so: the B200-native 3D-WDSR train/infer hot path of PROBA-V.
 *
 * The reference (mmbajo/PROBA-V) is pure Python on TensorFlow: it has no FFI of its own, its
 * boundary is Python call signatures.  Each entry point below names the reference interface it
 * replaces (file:line relative to the reference tree); the Python modules under proba-v_b200/ re-create those Python
 * signatures on top of this ABI (see INTEGRATION.md for the binding stub).
 *
 * Conventions
 *   - every function returns 0 on success or a negative pv_status; pv_last_error() gives the text
 *     (thread-local).  No C++ exception crosses the ABI.
 *   - *_dev pointers are device pointers on the handle's device; *_host pointers are host memory
 *     (pinned or pageable).  The caller owns every data buffer; the library owns only handles and
 *     their internal arenas.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls taking a stream
 *     are asynchronous on it; *_host variants synchronise before returning.
 *   - layouts are the reference's: LR [B,S,S,T,1] fp32 channels-last (models/modelsTF.py:19),
 *     HR/SR [B,Hh,Wh,1] fp32, mask [B,Hh,Wh,1] uint8/bool, 1 = clear (train.py:43).
 *   - weights are exchanged in TensorFlow layout: v [kh,kw,(kt,)Cin,Cout], g [Cout], bias [Cout]
 *     (TFA WeightNormalization variables), concatenated into one flat fp32
 *     arena whose per-tensor offsets pv_model_param_info reports.
 *   - a handle is bound to one device and is not thread-safe.
 */
#ifndef PROBAV_B200_H
#define PROBAV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PV_ABI_VERSION 1

typedef enum pv_status {
    PV_OK = 0,
    PV_ERR_BAD_CONFIG = -1,   /* cfg the reference graph cannot build (modelsTF.py:62-71)      */
    PV_ERR_BAD_ARG = -2,      /* NULL pointer, shape mismatch, unsupported size                  */
    PV_ERR_CUDA = -3,         /* CUDA runtime / launch failure                                   */
    PV_ERR_NO_DEVICE = -4,    /* no sm_100 device: there is NO CPU fallback                      */
    PV_ERR_STATE = -5         /* call order (e.g. backward before forward)                       */
} pv_status;

/* cfg-file fields the hot path consumes: [Net] + [Preprocessing] of cfg/p16t9c85r12.cfg
 * (utils/parseConfig.py:43-49,31-41) plus the per-band constants of train.py:47-52. */
typedef struct pv_cfg {
    int32_t num_res_blocks;
    int32_t num_low_res_imgs;   /* T: 7, 9, 13 or 19 (modelsTF.py:62-69)                      */
    int32_t scale;
    int32_t num_filters;
    int32_t kernel_size;
    int32_t exp_rate;
    float   decay_rate;
    int32_t is_grayscale;
    int32_t max_shift;
    int32_t patch_size;
    float   mean;               /* datasetAllMean (train.py:48,51)                             */
    float   std;                /* datasetAllStd  (train.py:49,52)                             */
    int32_t precision;          /* 0 = fp32 on CUDA cores (exact mode); 1 = tf32 on the tcgen05 tensor cores, fp32 accumulate
                                 * (what TensorFlow itself runs on Ampere-or-newer GPUs by default); 3 = fp32 CUDA-core
                                 * kernels on the tensor-core engine's row layouts (debug / cross-check);
                                 * 4 = error-compensated tf32 on the tensor cores ("tf32x3"): every forward product is computed as
                                 * x_hi*w_hi + x_lo*w_hi + x_hi*w_lo (hi = tf32(v), lo = v - hi) and the data gradients use
                                 * w_hi + w_lo, which brings SR to ~1e-6 and the gradients inside the 1e-3 bar */
} pv_cfg;

typedef enum pv_loss_kind {     /* cfg [Train] loss (train.py:93-100)                          */
    PV_LOSS_L1 = 0,             /* 'l1'            -> Losses.shiftCompensatedL1Loss  (loss.py:73-84)  */
    PV_LOSS_L2 = 1,             /* 'l2'            -> Losses.shiftCompensatedL2Loss  (loss.py:55-71)  */
    PV_LOSS_L1EDGE = 2          /* 'sobel_l1_mix'  -> Losses.shiftCompensatedL1EdgeLoss (loss.py:86-97) */
} pv_loss_kind;

typedef enum pv_opt_kind {      /* cfg [Train] optimizer (train.py:77-83)                      */
    PV_OPT_SGD = 0, PV_OPT_ADAM = 1, PV_OPT_NADAM = 2
} pv_opt_kind;

typedef struct pv_model pv_model;       /* replaces the tf.keras.Model returned by WDSRConv3D.build */
typedef struct pv_trainer pv_trainer;   /* replaces ModelTrainer's (model, loss, metric, optimizer)  */

/* ---- library ------------------------------------------------------------------------------- */
int         pv_abi_version(void);
const char* pv_last_error(void);
int         pv_device_count(void);                       /* number of sm_100 devices visible */

/* ---- graph builder: WDSRConv3D(name, band, mean, std, maxShift).build(...)  modelsTF.py:8-43 -- */
int  pv_model_create(const pv_cfg* cfg, int device, pv_model** out);
void pv_model_destroy(pv_model* m);
/* number of weight tensors (3 per weight-normalised conv: v, g, bias); order = arena order */
int  pv_model_param_count(const pv_model* m);
/* name is '<layer>/v' | '<layer>/g' | '<layer>/bias' with the reference's layer names
 * (mainConv1, expConv_i, decConv_i, normConv_i, convReducer_k, upscaleConv1, residConv1..3) */
int  pv_model_param_info(const pv_model* m, int idx, char* name, int name_cap,
                         int* rank, int64_t shape[5], int64_t* offset, int64_t* numel);
int64_t pv_model_param_numel(const pv_model* m);          /* flat arena length (535267 for p16t9c85r12) */
int  pv_model_set_params(pv_model* m, const float* flat_host, int64_t n);   /* replaces model.set_weights / ckpt.restore */
int  pv_model_get_params(pv_model* m, float* flat_host, int64_t n);         /* replaces model.get_weights / ckpt save    */
int  pv_model_param_arena(pv_model* m, float** dev_ptr, int64_t* n);        /* device view of the flat arena             */
/* TFA WeightNormalization first-call init g <- ||v|| */
int  pv_model_init_g_from_v(pv_model* m);
/* model(lr_batch, training=False)  -- modelsTF.py:15-43 called at trainClass.py:139, test.py:117 */
int  pv_forward(pv_model* m, const float* lr_dev, int B, float* sr_dev, void* stream);
int  pv_forward_host(pv_model* m, const float* lr_host, int B, float* sr_host);
/* test.py:114-122 resolve(): model -> clip_by_value(0, 2**16) -> round half-to-even.  Also testClass.py:24-30 */
int  pv_resolve(pv_model* m, const float* lr_dev, int B, float* sr_dev, void* stream);
int  pv_resolve_host(pv_model* m, const float* lr_host, int B, float* sr_host);
/* test.py:103-160 evaluate(): per scene, resolveByBatch over n*n patches + reconstruct_from_patches.
 * lr_patches [nscenes, n*n, S,S,T,1]; out [nscenes, n*P, n*P] row-major stitched (P = scale*patch). */
int  pv_predict_scenes_host(pv_model* m, const float* lr_patches_host, int nscenes, int patches_per_scene,
                            float* sr_scenes_host);
/* dataGenerator.py:108-121 + test.py:36-38 on device: [nscenes,T,H,W] LR scenes -> reflect-pad max_shift/2,
 * windows (patch+max_shift) stride patch -> model -> resolve -> stitched [nscenes, scale*H, scale*W]. */
int  pv_predict_from_scenes_host(pv_model* m, const float* lr_scenes_host, int nscenes, int H, int W,
                                 float* sr_scenes_host);
/* same with device-resident scenes (asynchronous on `stream`) */
int  pv_predict_from_scenes(pv_model* m, const float* lr_scenes_dev, int nscenes, int H, int W,
                            float* sr_scenes_dev, void* stream);

/* ---- losses: Losses(targetShape, cropBorder=3, bitDepth=16)  loss.py:13-35 ------------------- */
/* One pass over (hr, mask, sr) evaluates all (2*border+1)^2 shifts (loss.py:48-50 loop), the per-shift
 * bias (loss.py:182-187) and BOTH the selected loss and the cPSNR metric:
 *   loss_per_sample[b]  = min_shift score(kind)         (loss.py:83, :70, :96)
 *   best_shift[b]       = first arg-min, index i*(2*border+1)+j in the reference's stack order
 *   clear_count[b]      = N = sum(mask window) at that shift (loss.py:144)
 *   cpsnr_per_sample[b] = max_shift cPSNR               (loss.py:51-53)       (nullable)
 *   mean_loss[0]        = mean_b loss_per_sample        (loss.py:84)          (nullable)
 *   dsr                 = d(grad_scale * sum_b loss_per_sample)/d sr, [B,H,W] (nullable; fused backward)
 *   stack_out           = [B, S*S, 4] (L1, L2, N, bias) per shift             (nullable; parity/debug)
 * All pointers are device pointers. */
int  pv_shift_loss(int kind, const float* hr_dev, const uint8_t* mask_dev, const float* sr_dev,
                   int B, int H, int W, int border, float grad_scale,
                   float* loss_per_sample, int32_t* best_shift, int32_t* clear_count,
                   float* cpsnr_per_sample, float* mean_loss, float* dsr, float* stack_out, void* stream);
/* host-buffer convenience used by Losses.shiftCompensated* when handed numpy arrays
 * (also evaluate.py:76-87 whole-scene cPSNR with targetShape=(384,384,1)) */
int  pv_shift_loss_host(int kind, const float* hr_host, const uint8_t* mask_host, const float* sr_host,
                        int B, int H, int W, int border, float grad_scale,
                        float* loss_per_sample, int32_t* best_shift, int32_t* clear_count,
                        float* cpsnr_per_sample, float* mean_loss, float* dsr, float* stack_out);

/* ---- step loops: ModelTrainer  trainClass.py:25-143 ------------------------------------------ */
int  pv_trainer_create(pv_model* m, int opt_kind, float learning_rate, int loss_kind, pv_trainer** out);
void pv_trainer_destroy(pv_trainer* t);
/* trainStep (trainClass.py:124-135): fwd -> loss -> grads -> apply_gradients -> cPSNR metric.
 * out_dev[0] = loss (batch mean), out_dev[1] = mean cPSNR over the batch.  Asynchronous. */
int  pv_train_step(pv_trainer* t, const float* lr_dev, const float* hr_dev, const uint8_t* mask_dev,
                   int B, float* out_dev, void* stream);
/* same with host buffers: H2D of the batch and D2H of (loss, cPSNR) inside the call */
int  pv_train_step_host(pv_trainer* t, const float* lr_host, const float* hr_host, const uint8_t* mask_host,
                        int B, float* out_host);
/* testStep (trainClass.py:137-143): fwd -> loss -> metric */
int  pv_eval_step(pv_trainer* t, const float* lr_dev, const float* hr_dev, const uint8_t* mask_dev,
                  int B, float* out_dev, void* stream);
int  pv_eval_step_host(pv_trainer* t, const float* lr_host, const float* hr_host, const uint8_t* mask_host,
                       int B, float* out_host);
/* Split form of trainStep for data-parallel training: tape.gradient, then (after the caller has
 * all-reduced the flat gradient arena, debug/trainClassMultiGPU0.py:162-178) apply_gradients. */
int  pv_train_forward_backward(pv_trainer* t, const float* lr_dev, const float* hr_dev, const uint8_t* mask_dev,
                               int B, float grad_scale, float* out_dev, void* stream);
int  pv_trainer_grad_arena(pv_trainer* t, float** dev_ptr, int64_t* n);     /* same layout as the param arena */
/* The same in two stages, so that the caller can all-reduce the first gradient bucket while the second half of the
 * backward pass runs (TF's MirroredStrategy overlaps its per-variable all-reduces with backward the same way,
 * debug/trainMultiGPU.py:65-68).  stage 0: forward, loss, backward of the tail / reducers / last R/2 blocks;
 * stage 1: the remaining blocks and mainConv1 (lr/hr/mask/out are ignored).  On return [*grad_lo, *grad_hi) is the
 * contiguous range of the gradient arena that is final (in stream order).  Gradients are bit-identical to the
 * unstaged call. */
int  pv_train_forward_backward_staged(pv_trainer* t, const float* lr_dev, const float* hr_dev, const uint8_t* mask_dev,
                                      int B, float grad_scale, float* out_dev, int stage, int64_t* grad_lo,
                                      int64_t* grad_hi, void* stream);
int  pv_apply_gradients(pv_trainer* t, void* stream);
/* The two halves of tf.GradientTape for callers that bring their own loss (the PyTorch twin, models/modelsPyTorch.py):
 * pv_trainer_forward = model(lr, training=True) keeping the activations (trainClass.py:127); pv_trainer_backward =
 * tape.gradient(., model.trainable_variables) for a given dL/dSR [B,Hh,Wh,1] (trainClass.py:131): fills the gradient arena. */
int  pv_trainer_forward(pv_trainer* t, const float* lr_dev, int B, float* sr_dev, void* stream);
int  pv_trainer_backward(pv_trainer* t, const float* dsr_dev, int B, void* stream);
/* optimizer state for tf.train.Checkpoint parity (trainClass.py:33-39): iter, momentum_cache, m, v */
int  pv_trainer_get_state(pv_trainer* t, int64_t* iter, double* momentum_cache, float* m_host, float* v_host, int64_t n);
int  pv_trainer_set_state(pv_trainer* t, int64_t iter, double momentum_cache, const float* m_host, const float* v_host, int64_t n);
int  pv_trainer_set_lr(pv_trainer* t, float learning_rate);
/* number of kernels of this library launched so far (bench.py's gpu_launches) */
int64_t pv_launch_count(void);
/* Per-kernel-class device timing (bench.py's roofline; no reference counterpart).  While enabled every kernel
 * launch is bracketed by CUDA events on its stream.  pv_timing_report synchronises the device and writes one line
 * per kernel class: "name launches total_ms algorithmic_flops algorithmic_bytes executed_flops\n"; returns the number of bytes
 * written (or needed, if larger than cap). */
/* Device self-test: every tensor-core kernel configuration against the CUDA-core kernel on the same random buffers.
 * Returns the number of failing configurations (0 = all agree); the per-configuration report goes to buf. */
int  pv_selftest(char* buf, int cap);
/* Debug / numerics studies (no reference counterpart): copies the internal activation or gradient buffer `name` of the model's
 * training (train = 1) or inference pool to the host (device-synchronising).  Row-engine buffer names: a0..aR, D0.., g_a0, g_a1,
 * g_D, Gi1, Go1.., g_Go1.., U, g_U (csrc/engine_tc.cu).  Returns the buffer's length in floats (copying at most n), or a
 * negative pv_status when the buffer does not exist. */
int64_t pv_debug_read_buffer(pv_model* m, const char* name, int train, float* host, int64_t n);
int  pv_timing_enable(int on);
int  pv_timing_reset(void);
int  pv_timing_report(char* buf, int cap);

#ifdef __cplusplus
}
#endif
#endif /* PROBAV_B200_H */
