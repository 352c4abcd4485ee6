// conv_tc.cu -- implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma kind::tf32, fp32 accumulate
// in TMEM), operands staged in shared memory by TMA, bias / residual add / ReLU / ReLU-gradient mask / zero-padding mask
// fused into the TMEM epilogue.
//
// Replaces the cuDNN calls TensorFlow issues for Keras Conv3D inside TFA WeightNormalization (reference
// models/modelsTF.py:179-188 expConv / decConv / normConv, :159-163 convReducer / upscaleConv) and their
// Conv3DBackpropInputV2 (tape.gradient, models/trainClass.py:131).
//
// Formulation (rows.h): activations are [rows][32] fp32, one voxel per 128-byte row, ordered so that a tap is a
// constant row offset.  A CTA owns 128 consecutive output rows (GEMM M = 128).  For each group of taps (a "slab":
// the taps of one temporal plane) TMA brings rows [r0 + lo, r0 + lo + 128 + span) into shared memory ONCE in the
// 128B-swizzled K-major layout; every tap of the group is then the same tile viewed from a start address shifted by
// whole rows (verified on B200: the swizzle is a function of the absolute shared-memory address, probes/umma_probe.cu
// T2), so the 27-tap, K = 864 contraction issues 108 tcgen05.mma (M128 x N32 x K8) from 3 TMA loads and no im2col.
// Wide rows (K = 256: decConv forward, expConv data gradient) use the same mechanism with K-chunks as "taps".
//
// Warp roles (192 threads, persistent CTAs, one per SM): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2-5 = epilogue (one TMEM lane quarter each).  Pipelines: shared-memory slab ring (full/empty mbarriers, 4
// stages), TMEM accumulator double buffer (tfull/tempty), weights resident in shared memory for the CTA's lifetime.
// Roofline: tensor pipe; per output row 2*K*N flops.  With N = 32 the A operand is re-read from shared memory for
// every tap (4 KB per 16-cycle MMA), so shared-memory bandwidth, not the tensor pipe, is the expected limiter.
#include <cstdlib>
#include <vector>

#include "rows.h"
#include "tc_common.cuh"

namespace pv {

using namespace tc;

// ------------------------------------------------------------------------------------------ tensor maps (host)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_2d(CUtensorMap* m, const float* base, long long rows, int cols, int box_rows, int box_cols, int swizzle_32b_atom) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
            return set_error(PV_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle_32b_atom ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rows %lld cols %d box %dx%d", (int)r, rows, cols, box_rows, box_cols);
    return 0;
}

namespace {

constexpr int NSTAGE = 4;
constexpr int TC_THREADS = 320;     // warp 0 TMA, warp 1 MMA, warps 2-5 and 6-9: two epilogue groups (one per TMEM accumulator)

struct ConvTcArgs {
    // tiles
    int B, tiles_per_patch;
    long long in_lead, in_pstride;
    RowGeom og;
    // slabs and taps
    int nslab, slab_rows;
    int slab_lo[MAX_SLABS];            // first row of the slab relative to the tile's first input row
    int slab_c0[MAX_SLABS];            // channel coordinate of the slab (K-chunk of a wide row)
    int slab_tap0[MAX_SLABS + 1];      // taps [slab_tap0[s], slab_tap0[s+1]) belong to slab s (taps sorted by slab)
    int ntap;
    int tap_row[MAX_TAPS];             // row of the tap's view inside its slab
    int tap_wr[MAX_TAPS], tap_wc[MAX_TAPS];   // weight box coordinates (row, column) in the weight matrix
    // epilogue
    const float* bias; const float* residual; const float* relumask; float* y;
    int relu, round_tf32;
};

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

template <int NOUT>
__global__ void __launch_bounds__(TC_THREADS, 1)
rowconv_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const ConvTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * NSTAGE + 5];
    __shared__ uint32_t tmem_slot;
    constexpr uint32_t W_TAP_BYTES = NOUT * 128;                 // one tap's B operand: NOUT rows x 32 tf32
    constexpr int TMEM_COLS = 2 * NOUT < 32 ? 32 : 2 * NOUT;     // accumulator double buffer
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_smem = base;
    const uint32_t stage_bytes = (uint32_t)a.slab_rows * 128u;
    const uint32_t st_smem = base + (uint32_t)a.ntap * W_TAP_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto BAR = [&](int i) { return smem_u32(&bars[i]); };
    const int FULL = 0, EMPTY = NSTAGE, TFULL = 2 * NSTAGE, TEMPTY = 2 * NSTAGE + 2, WBAR = 2 * NSTAGE + 4;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(BAR(FULL + i), 1); mbar_init(BAR(EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(TFULL + i), 1); mbar_init(BAR(TEMPTY + i), 4); }
        mbar_init(BAR(WBAR), 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(&tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int ntiles = a.B * a.tiles_per_patch;

    if (warp == 0) {
        // ================================================================== TMA producer
        if (elect_one_sync()) {
            tma_prefetch_desc(&tm_x);
            tma_prefetch_desc(&tm_w);
            mbar_arrive_expect_tx(BAR(WBAR), (uint32_t)a.ntap * W_TAP_BYTES);
            for (int t = 0; t < a.ntap; ++t) tma_load_2d(w_smem + t * W_TAP_BYTES, &tm_w, BAR(WBAR), a.tap_wc[t], a.tap_wr[t]);
            pdl_wait();
            pdl_trigger();
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int b = tile / a.tiles_per_patch, j = tile % a.tiles_per_patch;
                const long long irow0 = a.in_lead + (long long)b * a.in_pstride + a.og.row0 + j * 128;
                for (int s = 0; s < a.nslab; ++s, ++it) {
                    const uint32_t stg = it % NSTAGE, ph = (it / NSTAGE) & 1;
                    mbar_wait(BAR(EMPTY + stg), ph ^ 1);
                    mbar_arrive_expect_tx(BAR(FULL + stg), stage_bytes);
                    tma_load_2d(st_smem + stg * stage_bytes, &tm_x, BAR(FULL + stg), a.slab_c0[s], (int)(irow0 + a.slab_lo[s]));
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer (one thread)
        if (elect_one_sync()) {
            constexpr uint64_t HI = smem_desc_hi(16, 1024, 2);          // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
            constexpr uint32_t IDESC = instr_desc(2, 128, NOUT, 0, 0);  // tf32 x tf32 -> f32, M = 128, N = NOUT
            uint32_t tap_inc[27];                     // start-address increments (16-byte units) of the 27 tap views
#pragma unroll
            for (int t = 0; t < 27; ++t) tap_inc[t] = (uint32_t)a.tap_row[t] * 8u;
            mbar_wait(BAR(WBAR), 0);
            tc_fence_after();
            uint32_t it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
                const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
                mbar_wait(BAR(TEMPTY + acc), aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem + acc * NOUT;
                uint32_t accumulate = 0;
                constexpr uint32_t HI32 = (uint32_t)(HI >> 32), LO32 = (uint32_t)HI;   // LO32: the LBO field (bits 16-29)
                if (a.nslab == 3 && a.ntap == 27) {
                    // 3x3x3 fast path: 3 slabs x 9 taps x 4 K-steps fully unrolled; per MMA one 32-bit add per descriptor
#pragma unroll
                    for (int s = 0; s < 3; ++s, ++it) {
                        const uint32_t stg = it % NSTAGE, ph = (it / NSTAGE) & 1;
                        mbar_wait(BAR(FULL + stg), ph);
                        tc_fence_after();
                        const uint32_t a_lo = ((st_smem + stg * stage_bytes) >> 4) | LO32;
                        const uint32_t b_lo = ((w_smem >> 4) | LO32) + (uint32_t)(s * 9) * (W_TAP_BYTES >> 4);
#pragma unroll
                        for (int t = 0; t < 9; ++t) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                umma_ss_tf32_lohi(d_tmem, a_lo + tap_inc[s * 9 + t] + 2 * ks, b_lo + t * (W_TAP_BYTES >> 4) + 2 * ks, HI32, IDESC, accumulate);
                                accumulate = 1;
                            }
                        }
                        umma_commit(BAR(EMPTY + stg));
                    }
                } else {
                    for (int s = 0; s < a.nslab; ++s, ++it) {
                        const uint32_t stg = it % NSTAGE, ph = (it / NSTAGE) & 1;
                        mbar_wait(BAR(FULL + stg), ph);
                        tc_fence_after();
                        const uint32_t s_addr = st_smem + stg * stage_bytes;
                        for (int t = a.slab_tap0[s]; t < a.slab_tap0[s + 1]; ++t) {
                            const uint32_t a_lo = ((s_addr + (uint32_t)a.tap_row[t] * 128u) >> 4) | LO32, b_lo = ((w_smem + (uint32_t)t * W_TAP_BYTES) >> 4) | LO32;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {             // 32 channels = 4 x K8
                                umma_ss_tf32_lohi(d_tmem, a_lo + 2 * ks, b_lo + 2 * ks, HI32, IDESC, accumulate);
                                accumulate = 1;
                            }
                        }
                        umma_commit(BAR(EMPTY + stg));  // slab may be overwritten once these MMAs have read it
                    }
                }
                umma_commit(BAR(TFULL + acc));          // accumulator complete
            }
        }
    } else {
        // ================================================================== epilogue: two groups of 4 warps (TMEM lane quarter =
        // warp % 4); group g drains accumulator g, i.e. every other tile, so one group's global-memory latency (residual /
        // mask prefetch, stores) overlaps the other group's tile
        const int q = warp & 3;
        const uint32_t grp = (uint32_t)(warp - 2) >> 2;
        uint32_t tl = 0;
        pdl_wait();
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
            const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
            if (acc != grp) continue;
            const int b = tile / a.tiles_per_patch, j = tile % a.tiles_per_patch;
            const int r = a.og.row0 + j * 128 + q * 32 + lane;            // row inside the patch
            const bool in_patch = r < a.og.row0 + a.og.nrows && r < a.og.pstride;
            const bool valid = in_patch && row_valid(a.og, r);
            const long long orow = a.og.lead + (long long)b * a.og.pstride + r;
            // N = 32: fetch this row's residual / ReLU-mask operands BEFORE waiting for the accumulator, so their
            // global-memory latency overlaps the MMAs of this tile instead of serialising behind them
            float4 pre_r[8], pre_m[8];
            if (NOUT == 32) {
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) {
                    pre_r[g4] = (a.residual && in_patch) ? __ldg(reinterpret_cast<const float4*>(a.residual + orow * NOUT) + g4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    pre_m[g4] = (a.relumask && in_patch) ? __ldg(reinterpret_cast<const float4*>(a.relumask + orow * NOUT) + g4) : make_float4(1.f, 1.f, 1.f, 1.f);
                }
            }
            mbar_wait(BAR(TFULL + acc), aph);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < NOUT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * NOUT + c0, v);
                tmem_ld_wait();
                if (c0 + 32 >= NOUT) {                  // last read of this accumulator: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(TEMPTY + acc));
                }
                if (!in_patch) continue;
                float* yp = a.y + orow * NOUT + c0;
                const float* rp = a.residual ? a.residual + orow * NOUT + c0 : nullptr;
                const float* mp = a.relumask ? a.relumask + orow * NOUT + c0 : nullptr;
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) {
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = __uint_as_float(v[g4 * 4 + e]);
                    if (a.bias) {
                        const float4 bq = __ldg(reinterpret_cast<const float4*>(a.bias + c0) + g4);
                        o[0] += bq.x; o[1] += bq.y; o[2] += bq.z; o[3] += bq.w;
                    }
                    if (rp) {
                        const float4 rq = NOUT == 32 ? pre_r[g4] : __ldg(reinterpret_cast<const float4*>(rp) + g4);
                        o[0] += rq.x; o[1] += rq.y; o[2] += rq.z; o[3] += rq.w;
                    }
                    if (a.relu) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
                    }
                    if (mp) {
                        const float4 mq = NOUT == 32 ? pre_m[g4] : __ldg(reinterpret_cast<const float4*>(mp) + g4);
                        o[0] = mq.x > 0.f ? o[0] : 0.f; o[1] = mq.y > 0.f ? o[1] : 0.f;
                        o[2] = mq.z > 0.f ? o[2] : 0.f; o[3] = mq.w > 0.f ? o[3] : 0.f;
                    }
                    if (!valid) { o[0] = o[1] = o[2] = o[3] = 0.f; }
                    if (a.round_tf32) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) o[e] = rna_tf32(o[e]);
                    }
                    reinterpret_cast<float4*>(yp)[g4] = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem);
}

}  // namespace

int launch_rowconv_tc(const RowConvP& p, cudaStream_t st) {
    static const bool force_n32 = getenv("PV_CONV3_N32") != nullptr;
    if (!force_n32 && rowconv3_tc_supported(p)) return launch_rowconv3_tc(p, st);
    if (p.residual2 || p.y_lo || p.y_pack || p.f16_pack) return set_error(PV_ERR_BAD_ARG, "rowconv_tc: extra addends / split output need the conv3 kernel");
    if (p.kc != 32 || (p.n != 32 && p.n != 256) || p.ntap < 1 || p.ntap > MAX_TAPS || !p.w_kmajor)
        return set_error(PV_ERR_BAD_ARG, "rowconv_tc: unsupported shape kc=%d n=%d ntap=%d kmajor=%d", p.kc, p.n, p.ntap, p.w_kmajor);
    if (p.og.row0 % 128) return set_error(PV_ERR_BAD_ARG, "rowconv_tc: row0 must be a multiple of 128");
    ConvTcArgs a;
    memset(&a, 0, sizeof a);
    a.B = p.B; a.tiles_per_patch = cdiv(p.og.nrows, 128);
    a.in_lead = p.in_lead; a.in_pstride = p.in_pstride; a.og = p.og;
    a.bias = p.bias; a.residual = p.residual; a.relumask = p.relumask; a.y = p.y; a.relu = p.relu; a.round_tf32 = p.round_tf32;
    // group the taps into slabs: same channel chunk, offsets within 64 rows of the slab's first tap (taps arrive sorted)
    int nslab = 0, span = 0;
    int slab_hi[MAX_SLABS];
    for (int t = 0; t < p.ntap; ++t) {
        const bool fits = nslab > 0 && p.c0[t] == a.slab_c0[nslab - 1] && p.off[t] >= a.slab_lo[nslab - 1] && p.off[t] - a.slab_lo[nslab - 1] <= 64;
        if (!fits) {
            if (nslab == MAX_SLABS) return set_error(PV_ERR_BAD_ARG, "rowconv_tc: more than %d slabs", MAX_SLABS);
            a.slab_lo[nslab] = p.off[t]; a.slab_c0[nslab] = p.c0[t]; a.slab_tap0[nslab] = t; slab_hi[nslab] = p.off[t];
            ++nslab;
        }
        if (p.off[t] > slab_hi[nslab - 1]) slab_hi[nslab - 1] = p.off[t];
        a.tap_row[t] = p.off[t] - a.slab_lo[nslab - 1];
        a.tap_wr[t] = p.wr0[t]; a.tap_wc[t] = p.wc0[t];
        if (slab_hi[nslab - 1] - a.slab_lo[nslab - 1] > span) span = slab_hi[nslab - 1] - a.slab_lo[nslab - 1];
    }
    a.slab_tap0[nslab] = p.ntap;
    a.nslab = nslab; a.ntap = p.ntap;
    a.slab_rows = ((128 + span + 7) / 8) * 8;
    if (a.slab_rows > 256) return set_error(PV_ERR_BAD_ARG, "rowconv_tc: slab of %d rows exceeds the TMA box limit", a.slab_rows);
    const size_t smem = 1024 + (size_t)p.ntap * p.n * 128 + (size_t)NSTAGE * a.slab_rows * 128;
    if (smem > 226 * 1024) return set_error(PV_ERR_BAD_ARG, "rowconv_tc: %zu bytes of shared memory needed", smem);

    const RowGeom& og = p.og;
    const long long in_rows = p.in_lead + (long long)p.B * p.in_pstride + ROW_TAIL;  // every row buffer is allocated with this zero tail
    CUtensorMap tm_x, tm_w;
    PV_TRY(make_tmap_2d(&tm_x, p.x, in_rows, p.xc, a.slab_rows, 32, 0));
    PV_TRY(make_tmap_2d(&tm_w, p.w, p.w_rows, p.w_cols, p.n, 32, 0));
    (void)og;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = a.B * a.tiles_per_patch;
    const int grid = ntiles < sms ? ntiles : sms;
    PV_TIMED(p.tag ? p.tag : "rowconv_tc", st, p.flops, 0.0);
    // opt in to the dynamic shared memory this configuration needs (static + dynamic must stay below 227 KB)
    static size_t attr32[16] = {}, attr256[16] = {};
    if (p.n == 32) {
        PV_CUDA(ensure_dyn_smem(rowconv_tc_kernel<32>, smem, attr32));
        PV_CUDA(launch_pdl(rowconv_tc_kernel<32>, grid, TC_THREADS, smem, st, tm_x, tm_w, a));
    } else {
        PV_CUDA(ensure_dyn_smem(rowconv_tc_kernel<256>, smem, attr256));
        PV_CUDA(launch_pdl(rowconv_tc_kernel<256>, grid, TC_THREADS, smem, st, tm_x, tm_w, a));
    }
    PV_LAUNCH_CHECK();
    return 0;
}

}  // namespace pv
