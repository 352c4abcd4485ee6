// umma_rate.cu -- issue-rate probe: how many cycles does one tcgen05.mma (kind::tf32, M = 128) cost when it is the
// only thing the SM does?  One thread per CTA issues NMMA back-to-back MMAs on resident (garbage) shared-memory
// operands and commits; cycles are taken with clock64 around issue + completion.  Variants: N, SS vs TS (A from TMEM),
// K-major vs MN-major operands, and a rotating A start address (different 4 KB windows, like the conv taps).
// Round 2 (is "32 + N/2" a same-accumulator dependency artefact?): ALTERNATING accumulators (two or four
// independent D regions, so consecutive MMAs never chain on one accumulator) and bf16 operands (kind::f16, K = 16 per MMA).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o probes/umma_rate probes/umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "../proba-v_b200/csrc/tc_common.cuh"
using namespace pv::tc;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// NACC / BF16 / TS are compile-time so that the issuing thread's loop stays a handful of instructions per MMA (with run-time
// variants the loop itself cost ~77 cycles per iteration and hid every effect below that)
template <int NACC, int BF16, int TS>
__global__ void __launch_bounds__(128) rate_kernel(int N, int mn, int rotate, int nmma, long long* cycles) {
    constexpr int nacc = NACC, bf16 = BF16, ts = TS;
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
        const uint32_t idesc = instr_desc(bf16 ? 1 : 2, 128, N, mn, mn);
        const uint32_t acc_stride = (uint32_t)N;                   // nacc accumulators of N columns each from column 0 (TS: A lives at 384..511)
        const uint32_t a0 = (base >> 4) | lo_bits, b0 = ((base + 65536) >> 4) | lo_bits;
        const long long t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            const uint32_t a_lo = a0 + (rotate ? (uint32_t)(i % 13) * 264u : 0u) + 2 * (i & 3);
            const uint32_t d = nacc > 1 ? tmem + (uint32_t)(i & (nacc - 1)) * acc_stride : tmem + 256;
            if (ts) umma_ts<true>(nacc > 1 ? d : tmem + 256, tmem + (nacc > 1 ? 384 : 0) + (i & 15) * 8, (((uint64_t)hi32) << 32) | (b0 + 2 * (i & 3)), idesc, 1);
            else if (bf16) umma_ss<false>(d, (((uint64_t)hi32) << 32) | a_lo, (((uint64_t)hi32) << 32) | (b0 + 2 * (i & 3)), idesc, 1);
            else umma_ss_tf32_lohi(d, a_lo, b0 + 2 * (i & 3), hi32, idesc, 1);
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
    auto launch = [&](int nacc, int bf16, int ts, int grid, int N, int mn, int rot) {
#define RK(A, B, T) if (nacc == A && bf16 == B && ts == T) { CK(cudaFuncSetAttribute(rate_kernel<A, B, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); rate_kernel<A, B, T><<<grid, 128, smem>>>(N, mn, rot, nmma, d); return; }
        RK(1, 0, 0) RK(1, 0, 1) RK(2, 0, 0) RK(4, 0, 0) RK(4, 0, 1) RK(1, 1, 0) RK(2, 1, 0)
#undef RK
        printf("no such variant\n"); exit(3);
    };
    struct V { const char* name; int N, ts, mn, rot, nacc, bf16; } vs[] = {
        {"SS K-major N=32", 32, 0, 0, 0, 1, 0}, {"SS K-major N=32 rotating A", 32, 0, 0, 1, 1, 0}, {"SS K-major N=64", 64, 0, 0, 0, 1, 0},
        {"SS K-major N=96", 96, 0, 0, 0, 1, 0}, {"SS K-major N=128", 128, 0, 0, 0, 1, 0}, {"SS K-major N=256", 256, 0, 0, 0, 1, 0},
        {"SS MN-major N=32", 32, 0, 1, 0, 1, 0},
        {"TS N=32 (A from TMEM)", 32, 1, 0, 0, 1, 0}, {"TS N=128", 128, 1, 0, 0, 1, 0},
        // independent accumulators: no MMA depends on its predecessor's D
        {"SS N=32, 4 accumulators", 32, 0, 0, 0, 4, 0}, {"SS N=96, 4 accumulators", 96, 0, 0, 0, 4, 0}, {"SS N=128, 2 accumulators", 128, 0, 0, 0, 2, 0},
        {"SS N=256, 2 accumulators", 256, 0, 0, 0, 2, 0}, {"TS N=32, 4 accumulators", 32, 1, 0, 0, 4, 0}, {"TS N=64, 4 accumulators", 64, 1, 0, 0, 4, 0},
        // bf16 operands (kind::f16, K = 16 per MMA: twice the MACs per MMA)
        {"bf16 SS N=32", 32, 0, 0, 0, 1, 1}, {"bf16 SS N=96", 96, 0, 0, 0, 1, 1}, {"bf16 SS N=128", 128, 0, 0, 0, 1, 1}, {"bf16 SS N=256", 256, 0, 0, 0, 1, 1},
        {"bf16 SS N=256, 2 accumulators", 256, 0, 0, 0, 2, 1},
        {"TS N=32, B MN-major", 32, 1, 1, 0, 1, 0}};       // (illegal instruction from this probe's descriptor; last on purpose)
    for (auto& v : vs)
        for (int grid : {1, 148}) {
            launch(v.nacc, v.bf16, v.ts, grid, v.N, v.mn, v.rot);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%-30s ERROR %s\n", v.name, cudaGetErrorString(e)); return 1; }
            long long h[148]; CK(cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
            const double cyc = (double)mx / nmma;
            const int K = v.bf16 ? 16 : 8;
            printf("%-30s grid %3d: %.1f cycles / MMA  -> %.0f MAC/cycle/SM (M128 x N%d x K%d)\n", v.name, grid, cyc, 128.0 * v.N * K / cyc, v.N, K);
        }
    return 0;
}
