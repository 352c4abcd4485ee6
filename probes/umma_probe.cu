// umma_probe.cu -- hardware fact-finding for the tensor-core kernels (not part of the product library).
// One CTA runs a TMA -> tcgen05.mma -> TMEM -> tcgen05.ld round trip under several descriptor recipes and compares
// with an exact CPU result (inputs are small dyadic rationals, so tf32 products and fp32 sums are exact).
//   T1  K-major SW128 operands loaded by TMA, N = 32 / 256
//   T2  A start address shifted by whole 128-byte rows (implicit-GEMM tap views), base_offset = 0 vs (addr>>7)&7
//   T3  MN-major operands (weight-gradient form: reduction over rows), incl. overlapping MN-atoms (LBO = 128 B)
//   T4  A operand from TMEM after an in-place ReLU epilogue (fused expand -> ReLU -> decay)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o probes/umma_probe probes/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>

#include "../proba-v_b200/csrc/tc_common.cuh"

using namespace pv::tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct TmaOp { int map; int c0, c1; uint32_t smem_off; uint32_t bytes; };
struct Operand { uint64_t hi; uint32_t off, inner, cnt, outer; int bo_auto; };
struct ProbeParams {
    int n_tma; TmaOp tma[14];
    Operand a, b; int nk; uint32_t idesc; int N;
    int ts; Operand b2; int nk2; uint32_t idesc2; int N2; int a2_cols_per_k;
    float* out; float* out2;
};

__device__ __forceinline__ uint64_t mk_desc(const Operand& o, uint32_t base, int ks) {
    const uint32_t addr = base + o.off + (ks / o.cnt) * o.outer + (ks % o.cnt) * o.inner;
    uint64_t d = smem_desc(o.hi, addr);
    if (o.bo_auto) d |= (uint64_t)((addr >> 7) & 7) << 49;
    return d;
}

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1,
                                                    const __grid_constant__ CUtensorMap m2, ProbeParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_tma = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    if (tid == 0) { mbar_init(bar_tma, 1); mbar_init(bar_mma, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        uint32_t total = 0;
        for (int i = 0; i < p.n_tma; ++i) total += p.tma[i].bytes;
        mbar_arrive_expect_tx(bar_tma, total);
        for (int i = 0; i < p.n_tma; ++i) {
            const CUtensorMap* m = p.tma[i].map == 0 ? &m0 : (p.tma[i].map == 1 ? &m1 : &m2);
            tma_load_2d(base + p.tma[i].smem_off, m, bar_tma, p.tma[i].c0, p.tma[i].c1);
        }
    }
    mbar_wait(bar_tma, 0);
    tc_fence_after();
    if (tid == 0) {
        for (int ks = 0; ks < p.nk; ++ks) umma_ss<true>(tmem, mk_desc(p.a, base, ks), mk_desc(p.b, base, ks), p.idesc, ks > 0);
        umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t v[32];
        const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        tmem_ld32(ta, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) p.out[(size_t)row * p.N + c0 + j] = __uint_as_float(v[j]);
        if (p.ts) {
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
            tmem_st32(ta, v);
            tmem_st_wait();
        }
    }
    if (p.ts) {
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            for (int ks = 0; ks < p.nk2; ++ks)
                umma_ts<true>(tmem + 256, tmem + ks * p.a2_cols_per_k, mk_desc(p.b2, base, ks), p.idesc2, ks > 0);
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, 1);
        tc_fence_after();
        for (int c0 = 0; c0 < p.N2; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c0, v);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) p.out2[(size_t)row * p.N2 + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled g_encode;

static CUtensorMap make_map(float* dptr, int cols, int rows, int box_cols, int box_rows, CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B) {
    CUtensorMap m;
    memset(&m, 0, sizeof m);
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(2); }
    return m;
}

static std::vector<float> rnd(size_t n, int range, float scale, unsigned seed) {
    std::vector<float> v(n);
    unsigned s = seed * 2654435761u + 12345u;
    for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; v[i] = (float)((int)((s >> 16) % (2 * range + 1)) - range) * scale; }
    return v;
}
static float* to_dev(const std::vector<float>& v) {
    float* d; CK(cudaMalloc(&d, v.size() * 4)); CK(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice)); return d;
}

static int run(const char* name, const CUtensorMap& m0, const CUtensorMap& m1, const CUtensorMap& m2, ProbeParams p,
               const std::vector<float>& exp1, const std::vector<float>* exp2 = nullptr) {
    float *o1, *o2;
    CK(cudaMalloc(&o1, 128 * 256 * 4)); CK(cudaMalloc(&o2, 128 * 256 * 4));
    CK(cudaMemset(o1, 0xFF, 128 * 256 * 4)); CK(cudaMemset(o2, 0xFF, 128 * 256 * 4));
    p.out = o1; p.out2 = o2;
    const int smem = 200 * 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem>>>(m0, m1, m2, p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s KERNEL ERROR: %s\n", name, cudaGetErrorString(e)); exit(3); }
    std::vector<float> g1(128 * p.N), g2(128 * (p.ts ? p.N2 : 1));
    CK(cudaMemcpy(g1.data(), o1, g1.size() * 4, cudaMemcpyDeviceToHost));
    double e1 = 0, e2 = 0; int bad = 0;
    for (size_t i = 0; i < g1.size(); ++i) { double d = fabs((double)g1[i] - exp1[i]); if (!(d <= 1e-5)) ++bad; if (d > e1 || d != d) e1 = d; }
    if (p.ts && exp2) {
        CK(cudaMemcpy(g2.data(), o2, g2.size() * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < g2.size(); ++i) { double d = fabs((double)g2[i] - (*exp2)[i]); if (!(d <= 1e-5)) ++bad; if (d > e2 || d != d) e2 = d; }
    }
    printf("%-44s %s  max_err1=%.4g max_err2=%.4g mismatches=%d  (got[0..3]=%g %g %g %g  exp=%g %g %g %g)\n", name,
           bad == 0 ? "PASS" : "FAIL", e1, e2, bad, g1[0], g1[1], g1[2], g1[3], exp1[0], exp1[1], exp1[2], exp1[3]);
    cudaFree(o1); cudaFree(o2);
    return bad == 0;
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    g_encode = (EncodeTiled)fn;
    if (!g_encode) { printf("no cuTensorMapEncodeTiled\n"); return 2; }

    const int XR = 512;
    std::vector<float> X = rnd((size_t)XR * 32, 2, 0.25f, 1);        // activations [rows x 32]
    std::vector<float> W32 = rnd(32 * 32, 2, 0.25f, 2);              // weights [N=32][K=32]
    std::vector<float> W256 = rnd(256 * 32, 2, 0.25f, 3);            // [N=256][K=32]
    std::vector<float> Y = rnd((size_t)128 * 128, 2, 0.25f, 4);      // [rows=128][128 ch]
    std::vector<float> G = rnd((size_t)128 * 32, 2, 0.25f, 5);       // [rows=128][32]
    std::vector<float> W64 = rnd(64 * 32, 2, 0.25f, 6);              // [N=64][K=32]
    std::vector<float> V = rnd(32 * 64, 2, 0.25f, 7);                // [N=32][K=64]
    float *dX = to_dev(X), *dW32 = to_dev(W32), *dW256 = to_dev(W256), *dY = to_dev(Y), *dG = to_dev(G), *dW64 = to_dev(W64), *dV = to_dev(V);

    CUtensorMap mX = make_map(dX, 32, XR, 32, 256);
    CUtensorMap mW32 = make_map(dW32, 32, 32, 32, 32);
    CUtensorMap mW256 = make_map(dW256, 32, 256, 32, 256);
    CUtensorMap mY = make_map(dY, 128, 128, 32, 128);
    CUtensorMap mG = make_map(dG, 32, 128, 32, 128);
    CUtensorMap mW64 = make_map(dW64, 32, 64, 32, 64);
    CUtensorMap mV = make_map(dV, 64, 32, 32, 32);

    const uint64_t HI_K = smem_desc_hi(16, 1024);          // K-major SW128: SBO = 8 rows x 128 B; LBO ignored
    int ok = 1;
    // ---- T1: plain K-major
    for (int N : {32, 256}) {
        ProbeParams p; memset(&p, 0, sizeof p);
        p.n_tma = 2;
        p.tma[0] = {0, 0, 0, 0, 256 * 128};
        p.tma[1] = {1, 0, 0, 65536, (uint32_t)N * 128};
        p.a = {HI_K, 0, 32, 4, 0, 0}; p.b = {HI_K, 65536, 32, 4, 0, 0};
        p.nk = 4; p.N = N; p.idesc = instr_desc(2, 128, N, 0, 0);
        const std::vector<float>& W = N == 32 ? W32 : W256;
        std::vector<float> e((size_t)128 * N);
        for (int i = 0; i < 128; ++i) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < 32; ++k) s += X[i * 32 + k] * W[n * 32 + k]; e[(size_t)i * N + n] = s; }
        char nm[64]; snprintf(nm, sizeof nm, "T1 K-major SW128 tf32 N=%d", N);
        ok &= run(nm, mX, N == 32 ? mW32 : mW256, mG, p, e);
    }
    // ---- T2: row-shifted A views
    for (int bo = 0; bo < 2; ++bo)
        for (int delta : {1, 5, 8, 24, 77}) {
            ProbeParams p; memset(&p, 0, sizeof p);
            p.n_tma = 2;
            p.tma[0] = {0, 0, 0, 0, 256 * 128};
            p.tma[1] = {1, 0, 0, 65536, 32 * 128};
            p.a = {HI_K, (uint32_t)delta * 128, 32, 4, 0, bo}; p.b = {HI_K, 65536, 32, 4, 0, 0};
            p.nk = 4; p.N = 32; p.idesc = instr_desc(2, 128, 32, 0, 0);
            std::vector<float> e(128 * 32);
            for (int i = 0; i < 128; ++i) for (int n = 0; n < 32; ++n) { float s = 0; for (int k = 0; k < 32; ++k) s += X[(i + delta) * 32 + k] * W32[n * 32 + k]; e[i * 32 + n] = s; }
            char nm[64]; snprintf(nm, sizeof nm, "T2 row-shift delta=%d base_offset=%s", delta, bo ? "(addr>>7)&7" : "0");
            run(nm, mX, mW32, mG, p, e);
        }
    // ---- T3: MN-major tf32 operands need the 128B-swizzle-with-32B-atom layout (descriptor layout type 1; TMA
    //      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atom = 128 B (32 tf32 along MN) x 4 K-rows, SBO = stride between K-atoms
    CUtensorMap mX32 = make_map(dX, 32, XR, 32, 256, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    CUtensorMap mY32 = make_map(dY, 128, 128, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    CUtensorMap mG32 = make_map(dG, 32, 128, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    for (int variant = 0; variant < 3; ++variant) {
        // variant 0: LBO = MN-atom stride, SBO = K-atom stride (512 B); 1: swapped; 2: layout type 2 data with type-1 desc (control)
        ProbeParams p; memset(&p, 0, sizeof p);
        p.n_tma = 5;
        for (int a = 0; a < 4; ++a) p.tma[a] = {0, a * 32, 0, (uint32_t)a * 16384, 128 * 128};
        p.tma[4] = {1, 0, 0, 65536, 128 * 128};
        const uint32_t lbo = variant == 1 ? 512 : 16384, sbo = variant == 1 ? 16384 : 512;
        p.a = {smem_desc_hi(lbo, sbo, 1), 0, 1024, 1000, 0, 0};
        p.b = {smem_desc_hi(lbo, sbo, 1), 65536, 1024, 1000, 0, 0};
        p.nk = 16; p.N = 32; p.idesc = instr_desc(2, 128, 32, 1, 1);
        std::vector<float> e(128 * 32);
        for (int j = 0; j < 128; ++j) for (int n = 0; n < 32; ++n) { float s = 0; for (int m = 0; m < 128; ++m) s += Y[m * 128 + j] * G[m * 32 + n]; e[j * 32 + n] = s; }
        char nm[96]; snprintf(nm, sizeof nm, "T3a MN-major tf32 SW128_32B variant %d", variant);
        if (variant == 2) run(nm, mY, mG, mG, p, e); else run(nm, mY32, mG32, mG32, p, e);
    }
    // A MN-major (from Y), B K-major?  not needed.  Overlapping MN atoms: D[(q,c)][n] = sum_m X[m+q+d0][c] * G[m][n]
    for (int bo = 0; bo < 2; ++bo)
        for (int d0 : {0, 4, 5}) {
            ProbeParams p; memset(&p, 0, sizeof p);
            p.n_tma = 2;
            p.tma[0] = {0, 0, 0, 0, 256 * 128};
            p.tma[1] = {1, 0, 0, 65536, 128 * 128};
            p.a = {smem_desc_hi(128, 512, 1), (uint32_t)d0 * 128, 1024, 1000, 0, bo};
            p.b = {smem_desc_hi(16384, 512, 1), 65536, 1024, 1000, 0, 0};
            p.nk = 16; p.N = 32; p.idesc = instr_desc(2, 128, 32, 1, 1);
            std::vector<float> e(128 * 32);
            for (int q = 0; q < 4; ++q) for (int c = 0; c < 32; ++c) for (int n = 0; n < 32; ++n) {
                float s = 0; for (int m = 0; m < 128; ++m) s += X[(m + q + d0) * 32 + c] * G[m * 32 + n]; e[(q * 32 + c) * 32 + n] = s; }
            char nm[96]; snprintf(nm, sizeof nm, "T3b overlapping MN atoms (LBO=128) d0=%d bo=%d", d0, bo);
            run(nm, mX32, mG32, mG32, p, e);
        }
    // mixed: A MN-major (gZ-like, 4 atoms) with N = 32 from a K-major-swizzled buffer is not needed; but check M=128,N=256:
    {
        // D[j][n] = sum_m G[m][j%32 + ...]: use A = X (overlap trick) and B = Y (N = 128 channels -> 4 atoms, LBO 16K)
        ProbeParams p; memset(&p, 0, sizeof p);
        p.n_tma = 5;
        p.tma[0] = {0, 0, 0, 0, 256 * 128};
        for (int a = 0; a < 4; ++a) p.tma[1 + a] = {1, a * 32, 0, 65536 + (uint32_t)a * 16384, 128 * 128};
        p.a = {smem_desc_hi(128, 512, 1), 0, 1024, 1000, 0, 0};
        p.b = {smem_desc_hi(16384, 512, 1), 65536, 1024, 1000, 0, 0};
        p.nk = 16; p.N = 128; p.idesc = instr_desc(2, 128, 128, 1, 1);
        std::vector<float> e(128 * 128);
        for (int q = 0; q < 4; ++q) for (int c = 0; c < 32; ++c) for (int n = 0; n < 128; ++n) {
            float s = 0; for (int m = 0; m < 128; ++m) s += X[(m + q) * 32 + c] * Y[m * 128 + n]; e[(q * 32 + c) * 128 + n] = s; }
        run("T3c MN-major A(overlap) x MN-major B N=128 (4 atoms)", mX32, mY32, mG32, p, e);
    }
    // ---- T4: TS mode after in-place ReLU
    {
        ProbeParams p; memset(&p, 0, sizeof p);
        p.n_tma = 4;
        p.tma[0] = {0, 0, 0, 0, 256 * 128};
        p.tma[1] = {1, 0, 0, 65536, 64 * 128};
        p.tma[2] = {2, 0, 0, 81920, 32 * 128};
        p.tma[3] = {2, 32, 0, 81920 + 4096, 32 * 128};
        p.a = {HI_K, 0, 32, 4, 0, 0}; p.b = {HI_K, 65536, 32, 4, 0, 0};
        p.nk = 4; p.N = 64; p.idesc = instr_desc(2, 128, 64, 0, 0);
        p.ts = 1; p.b2 = {HI_K, 81920, 32, 4, 4096, 0}; p.nk2 = 8; p.N2 = 32; p.idesc2 = instr_desc(2, 128, 32, 0, 0); p.a2_cols_per_k = 8;
        std::vector<float> e1(128 * 64), e2(128 * 32);
        for (int i = 0; i < 128; ++i) for (int n = 0; n < 64; ++n) { float s = 0; for (int k = 0; k < 32; ++k) s += X[i * 32 + k] * W64[n * 32 + k]; e1[i * 64 + n] = s; }
        for (int i = 0; i < 128; ++i) for (int n = 0; n < 32; ++n) { float s = 0; for (int k = 0; k < 64; ++k) s += fmaxf(e1[i * 64 + k], 0.f) * V[n * 64 + k]; e2[i * 32 + n] = s; }
        ok &= run("T4 TS-mode (A from TMEM after in-place ReLU)", mX, mW64, mV, p, e1, &e2);
    }
    printf("probe done, core tests %s\n", ok ? "OK" : "FAILED");
    return 0;
}
