// common.cuh -- error plumbing and small device helpers shared by every translation unit.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>

#include "../../include/probav_b200.h"

namespace pv {

// thread-local last-error text behind pv_last_error()
std::string& last_error();
int set_error(int code, const char* fmt, ...);
// every kernel launch of this library goes through here so bench.py can report gpu_launches
void count_launch(int n = 1);
int64_t launch_count();

#define PV_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            cudaGetLastError(); /* clear the non-sticky error so later launches are not blamed */ \
            return pv::set_error(PV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);                 \
        }                                                                                     \
    } while (0)

#define PV_LAUNCH_CHECK()                                                                     \
    do {                                                                                      \
        pv::count_launch();                                                                   \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return pv::set_error(PV_ERR_CUDA, "kernel launch failed: %s (%s:%d)",            \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);                 \
    } while (0)

#define PV_TRY(expr)                                                                          \
    do {                                                                                      \
        int _s = (expr);                                                                      \
        if (_s != 0) return _s;                                                               \
    } while (0)

// Launch with programmatic stream serialization (see tc_common.cuh pdl_wait / pdl_trigger).  PV_NO_PDL=1 falls back to a
// plain launch (the griddepcontrol instructions are no-ops then).
bool pdl_enabled();
bool pdl_simple_enabled();     // PV_PDL_SIMPLE=1 also launches the small elementwise / reduction kernels with PDL.  Off by default:
                               // measured 6.20 vs 6.12 ms per step (their early-launched blocks wait on SM resources the
                               // persistent tensor-core kernels' tails could use)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// device side of launch_pdl for the simple (non tensor-core) kernels: first statement of the kernel.  They never trigger
// their dependents explicitly (the implicit trigger at block exit comes after this wait, which keeps the chain transitive).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_grid_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_simple(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    if (pdl_simple_enabled()) return launch_pdl(kernel, grid, block, smem, st, static_cast<KArgs>(args)...);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Per-kernel-class device timing behind pv_timing_* (bench.py's roofline): when enabled, a launch is bracketed by
// CUDA events on its own stream and tagged with its ALGORITHMIC flops / bytes (unpadded shapes: the roofline numerator) and,
// where they differ, the flops the kernel EXECUTES (recomputation, channel padding, extra compensation passes).
struct KernelTimer {
    KernelTimer(const char* name, cudaStream_t st, double flops = 0.0, double bytes = 0.0, double exec_flops = 0.0);
    ~KernelTimer();
    int slot;
    cudaStream_t st;
};
#define PV_TIMED(name, st, ...) pv::KernelTimer _pv_kt(name, st, ##__VA_ARGS__)


// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: cache the value set per device ordinal so that
// a process driving several GPUs (tests on a multi-GPU box) raises the limit on each of them.
template <typename F>
inline cudaError_t ensure_dyn_smem(F* func, size_t smem, size_t (&cache)[16]) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 15;
    if (smem > cache[dev]) {
        const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cache[dev] = smem;
    }
    return cudaSuccess;
}

}  // namespace pv
