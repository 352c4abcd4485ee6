// umma_rate.cu -- issue-rate probe: how many cycles does one tcgen05.mma (kind::tf32, M = 128) cost when it is the
// only thing the SM does?  One thread per CTA issues NMMA back-to-back MMAs on resident (garbage) shared-memory
// operands and commits; cycles are taken with clock64 around issue + completion.  Variants: N, SS vs TS (A from TMEM),
// K-major vs MN-major operands, and a rotating A start address (different 4 KB windows, like the conv taps).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o probes/umma_rate probes/umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "../proba-v_b200/csrc/tc_common.cuh"
using namespace pv::tc;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void __launch_bounds__(128) rate_kernel(int N, int ts, int mn, int rotate, int nmma, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1 && elect_one_sync()) {
        const uint64_t HI = mn ? smem_desc_hi(128, 512, 1) : smem_desc_hi(16, 1024, 2);
        const uint32_t hi32 = (uint32_t)(HI >> 32), lo_bits = (uint32_t)HI;
        const uint32_t idesc = instr_desc(2, 128, N, mn, mn);
        const uint32_t a0 = (base >> 4) | lo_bits, b0 = ((base + 65536) >> 4) | lo_bits;
        const long long t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            const uint32_t a_lo = a0 + (rotate ? (uint32_t)(i % 13) * 264u : 0u) + 2 * (i & 3);
            if (ts) umma_ts<true>(tmem + 256, tmem + (i & 15) * 8, (((uint64_t)hi32) << 32) | (b0 + 2 * (i & 3)), idesc, 1);
            else umma_ss_tf32_lohi(tmem + 256, a_lo, b0 + 2 * (i & 3), hi32, idesc, 1);
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
    long long* d; CK(cudaMalloc(&d, 148 * 8));
    const int smem = 200 * 1024, nmma = 4096;
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    struct V { const char* name; int N, ts, mn, rot; } vs[] = {
        {"SS K-major N=32", 32, 0, 0, 0}, {"SS K-major N=32 rotating A", 32, 0, 0, 1}, {"SS K-major N=64", 64, 0, 0, 0},
        {"SS K-major N=128", 128, 0, 0, 0}, {"SS K-major N=256", 256, 0, 0, 0}, {"SS MN-major N=32", 32, 0, 1, 0},
        {"TS N=32 (A from TMEM)", 32, 1, 0, 0}, {"TS N=128", 128, 1, 0, 0}, {"TS N=32, B MN-major", 32, 1, 1, 0}};
    for (auto& v : vs)
        for (int grid : {1, 148}) {
            rate_kernel<<<grid, 128, smem>>>(v.N, v.ts, v.mn, v.rot, nmma, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%-30s ERROR %s\n", v.name, cudaGetErrorString(e)); return 1; }
            long long h[148]; CK(cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
            const double cyc = (double)mx / nmma;
            printf("%-30s grid %3d: %.1f cycles / MMA  -> %.0f MAC/cycle/SM (M128 x N%d x K8)\n", v.name, grid, cyc, 128.0 * v.N * 8 / cyc, v.N);
        }
    return 0;
}
