// resblock_tc.cu -- the 1x1x1 expand (x8) -> ReLU -> decay chain of a WDSR residual block as ONE tcgen05 kernel per
// direction; the 256-channel expanded tensor lives only in TMEM (reference models/modelsTF.py:179-183 ResConv3D:
// expConv_i + ReLU, decConv_i; the reference materialises a 571 MB tensor per block here).
//
//   forward   D = (relu(X We^T + be)) Wd^T + bd                                    resfront_fwd_kernel
//       per 128-row tile and per half h of the 256 expanded channels:
//         MMA1 (SS)  E_h[128 x 128] = X[128 x 32] . We_h^T          operands from shared memory (TMA), result in TMEM
//         epilogue   E_h <- tf32(relu(E_h + be_h))  IN PLACE in TMEM (tcgen05.ld -> registers -> tcgen05.st)
//         MMA2 (TS)  D[128 x 32] += E_h . Wd_h^T                     A operand read straight from TMEM
//       tcgen05 instructions execute in issue order, so MMA1 of the next half simply queues behind MMA2 of this one.
//       Two CTAs per SM (256 TMEM columns, ~98 KB shared memory each) overlap one CTA's epilogue with the other's MMAs.
//
// Layouts as in rows.h (PR rows, 32 channels, 128 B per row); weights are the effective (weight-normalised, tf32)
// matrices prepared by wn_prep: weT_exp [256][32], weT_dec [32][256] (both K contiguous).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "rowio.cuh"
#include "rows.h"
#include "tc_common.cuh"
#include "wgrad_reduce.cuh"

namespace pv {

using namespace tc;
int make_tmap_2d(CUtensorMap* m, const float* base, long long rows, int cols, int box_rows, int box_cols, int swizzle_32b_atom);

namespace {

constexpr int RF_THREADS = 192;

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// Round-to-nearest tf32 for a value that only feeds a kind::tf32 MMA from TMEM: add half a tf32 ulp to the bit pattern
// (sign-magnitude, so this rounds half away from zero) and leave the low 13 bits alone -- the tensor core ignores them.
// One integer add instead of the FSETP / IADD / LOP3 sequence cvt.rna.tf32.f32 compiles to; the epilogues below are
// issue-bound on exactly these per-element instructions (PV_DEBUG timing experiments: the element math was half of the
// forward kernel's run time).  Non-finite inputs are not preserved (an Inf would become a NaN pattern).
__device__ __forceinline__ uint32_t tf32_bump(uint32_t bits) { return bits + 0x1000u; }

// ----------------------------------------------------------------------------------------------------------------
// Forward and backward-data share one kernel (MODE 0 = training forward, 1 = backward-data, 2 = inference forward: the same
// as 0 without the ReLU bit mask, i.e. 3 instead of 5 instructions per expanded element in the epilogue): both are
//     MMA1 (SS)  H_h[128 x 128] = T[128 x 32] . W1_h^T      T = X (fwd) | gD (bwd);  W1 = We^T (fwd) | Wd (bwd)
//     epilogue   H_h <- f(H_h) in place in TMEM               fwd: tf32(relu(. + be)), emits the ReLU bit mask (32 B / row)
//                                                             bwd: tf32(.) where the forward's mask bit is set, else 0
//     MMA2 (TS)  O[128 x 32] += H_h . W2_h^T                 W2 = Wd^T (fwd) | We (bwd)
//     final      fwd: O + bd;  bwd: O + G (skip connection), ReLU mask of the layer below;  zero-padding mask, store
// The unit of pipelining is a half tile (128 rows x 128 expanded channels).  TMEM holds three H buffers (384 columns) and
// two O accumulators (64 columns): the MMA thread issues MMA1 of unit u+1 BEFORE MMA2 of unit u, so the tensor pipe works
// on the next unit while the epilogue warps stream unit u through tcgen05.ld/st (TMEM reads run at ~64-100 B/cycle/SM,
// profiles/r01_tmem_rate_probe.log, and are this kernel's floor).  tcgen05 instructions execute in issue order, which
// is what makes re-using an H buffer three units later safe without another barrier.  Two epilogue groups of four
// warps take alternate tiles.
constexpr int RP_THREADS = 320;
// Epilogue groups (of four warps) of the forward / backward-data kernel: template parameter G.  The backward-data kernel runs
// three groups taking every third tile (448 threads, 128 registers; measured 0.917 -> 0.857 ms per step), the forward kernels two
// (three were 2 % slower there).  With three groups a group returns to the same H buffer two barrier phases later, which a
// parity wait cannot tell from "already complete", so the "H ready" barriers are then indexed by u mod 6: every barrier has
// ONE waiting group and consecutive phases.  (The accumulator barriers are safe: tcgen05 commits arrive in issue order.)
constexpr int respipe_groups(int mode) { return mode == 1 ? 3 : 2; }
// three groups: 512 threads = {producer, MMA, two idle warps} + 12 epilogue warps.  The first warpgroup hands its registers back
// (setmaxnreg.dec 56) and the epilogue warps take 152 (4 x 32 x 56 + 12 x 32 x 152 = 65 536, and per scheduler 56 + 3 x 152 = 512
// registers per lane): three warps per scheduler without the spills a plain 448-thread launch has at its 128-register cap.  The role
// split has to happen at warpgroup level (setmaxnreg is a warpgroup-wide instruction, and ptxas budgets each side of the branch).
constexpr int respipe_threads(int g) { return g == 3 ? 512 : 64 + 128 * g; }
constexpr int respipe_first_epi_warp(int g) { return g == 3 ? 4 : 2; }
constexpr int respipe_nef(int g) { return g == 2 ? 3 : 2 * g; }       // number of "H ready" (EFULL) barriers

struct ResPipeArgs {
    int B, tiles_per_patch;
    RowGeom g;
    const float* bias1;                // fwd: be [256]
    const float* bias2;                // fwd: bd [32]
    uint32_t* mask;                    // [rows][8] ReLU bit mask: written by fwd (nullable), read by bwd
    const float* residual;             // bwd: G rows
    const float* relumask;             // bwd: rows or nullptr
    float* out;                        // rows [.. x 32]
    float* out_pack;                   // bwd, nullable: the same rows as bf16 pairs [bf16(v) | bf16(v - bf16(v))] of the un-rounded result
    int round_tf32;                    //   (what the next block's single-launch 3x3x3 data gradient reads, conv3_tc.cu MODE 2)
    int out_f16;                       // inference fwd: `out` receives fp16 pair rows [fp16(tf32(D)) x 32 | 0] instead of fp32 rows (conv3_tc.cu MODE 3)
};

// WSPLIT (backward-data of the error-compensated engine, precision 4): both weight matrices come as hi + lo (hi = tf32(w),
// lo = w - hi) and every product is issued twice, T.W_hi + T.W_lo -- the rounding of the weights is a SYSTEMATIC perturbation
// of the data-gradient chain (the same for every row, accumulating over the 36 layers), unlike the rounding of the gradients.
template <int MODE, bool WSPLIT>
__global__ void __launch_bounds__(respipe_threads(respipe_groups(MODE)), 1)
resfront_pipe_kernel(const __grid_constant__ CUtensorMap tm_t, const __grid_constant__ CUtensorMap tm_w1,
                     const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_w1l,
                     const __grid_constant__ CUtensorMap tm_w2l, const ResPipeArgs a) {
    constexpr int RPG = respipe_groups(MODE), RPP_THREADS = respipe_threads(RPG), NEF = respipe_nef(RPG), EW0 = respipe_first_epi_warp(RPG);
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[14 + NEF];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_b2[32];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr uint32_t WBYTES = WSPLIT ? 131072u : 65536u;
    const uint32_t w1_smem = base;                  // [256 rows x 128 B]: two halves of 128 rows
    const uint32_t w2_smem = base + 32768;          // 8 K-chunks x [32 rows x 128 B]
    const uint32_t w1l_smem = base + 65536, w2l_smem = base + 98304;     // WSPLIT: the lo halves, same layouts
    const uint32_t t_smem = base + WBYTES;          // 3 stages x [128 rows x 128 B]
    uint8_t* const io_scratch = smem_raw + (base - smem_u32(smem_raw)) + WBYTES + 3 * 16384;   // 8 epilogue warps x 2 KB (rowio.cuh)
    // Forward (MODE 0 / 2): the expand bias rides in MMA1 as one more K = 8 step, A = a [128 x 8] tile whose columns 0 and 1 are ones,
    // B = a [256 x 8] tile whose columns 0 / 1 are tf32(be) / be - tf32(be) -- one MMA per unit instead of an FADD per expanded element in
    // an epilogue that is bound by its instruction stream.  Both tiles are K-major SWIZZLE_128B like the operands.
    constexpr uint32_t SCRATCH_BYTES = 4 * RPG * ROWIO_SCRATCH_BYTES;
    const uint32_t ones_smem = base + WBYTES + 3 * 16384 + SCRATCH_BYTES, biasb_smem = ones_smem + 16384;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto BAR = [&](int i) { return smem_u32(&bars[i]); };
    const int FULL = 0, EMPTY = 3, WBAR = 6, EFULL = 7, EREADY = 7 + NEF, DFULL = 10 + NEF, DFREE = 12 + NEF;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(BAR(FULL + i), 1); mbar_init(BAR(EMPTY + i), 1); mbar_init(BAR(EREADY + i), 4); }
        for (int i = 0; i < NEF; ++i) mbar_init(BAR(EFULL + i), 1);
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(DFULL + i), 1); mbar_init(BAR(DFREE + i), 4); }
        mbar_init(BAR(WBAR), 1);
        fence_mbar_init();
    }
    if (MODE != 1) {
        if (threadIdx.x < 32) s_b2[threadIdx.x] = a.bias2[threadIdx.x];
        uint8_t* const ones_p = smem_raw + (ones_smem - smem_u32(smem_raw));
        for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += RPP_THREADS) reinterpret_cast<uint4*>(ones_p)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        // logical 16-byte chunk 0 of row r sits at physical chunk (r & 7) (SWIZZLE_128B: chunk index ^= row & 7)
        for (int r = threadIdx.x; r < 128; r += RPP_THREADS)
            *reinterpret_cast<float2*>(ones_p + r * 128 + ((r & 7) << 4)) = make_float2(1.f, 1.f);
        for (int n = threadIdx.x; n < 256; n += RPP_THREADS) {
            const float b = a.bias1[n], bh = rna_tf32(b);
            *reinterpret_cast<float2*>(ones_p + 16384 + n * 128 + ((n & 7) << 4)) = make_float2(bh, b - bh);
        }
        fence_proxy_async();                        // generic-proxy writes -> visible to the tensor core's async-proxy reads
    }
    if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;                // H buffers at columns 0 / 128 / 256, O accumulators at 384 / 416
    const int ntiles = a.B * a.tiles_per_patch;
    const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp < EW0) {
      if (RPG == 3) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");       // the whole first warpgroup (two of its warps are idle)
      if (warp == 0) {
        if (elect_one_sync()) {
            tma_prefetch_desc(&tm_t);
            mbar_arrive_expect_tx(BAR(WBAR), WBYTES);
            tma_load_2d(w1_smem, &tm_w1, BAR(WBAR), 0, 0);
            for (int j = 0; j < 8; ++j) tma_load_2d(w2_smem + j * 4096, &tm_w2, BAR(WBAR), 32 * j, 0);
            if (WSPLIT) {
                tma_load_2d(w1l_smem, &tm_w1l, BAR(WBAR), 0, 0);
                for (int j = 0; j < 8; ++j) tma_load_2d(w2l_smem + j * 4096, &tm_w2l, BAR(WBAR), 32 * j, 0);
            }
            pdl_wait();
            pdl_trigger();
            for (int tl = 0; tl < my_tiles; ++tl) {
                const int tile = blockIdx.x + tl * gridDim.x;
                const int b = tile / a.tiles_per_patch, j = tile % a.tiles_per_patch;
                const long long row0 = a.g.lead + (long long)b * a.g.pstride + a.g.row0 + j * 128;
                const uint32_t stg = tl % 3, ph = (tl / 3) & 1;
                mbar_wait(BAR(EMPTY + stg), ph ^ 1);
                mbar_arrive_expect_tx(BAR(FULL + stg), 16384);
                tma_load_2d(t_smem + stg * 16384, &tm_t, BAR(FULL + stg), 0, (int)row0);
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            constexpr uint64_t HI = smem_desc_hi(16, 1024, 2);
            constexpr uint32_t HI32 = (uint32_t)(HI >> 32), LO32 = (uint32_t)HI;
            constexpr uint32_t IDESC1 = instr_desc(2, 128, 128, 0, 0);
            constexpr uint32_t IDESC2 = instr_desc(2, 128, 32, 0, 0);
            mbar_wait(BAR(WBAR), 0);
            tc_fence_after();
            const int U = 2 * my_tiles;
            for (int u = 0; u <= U; ++u) {
                if (u < U) {                        // MMA1 of unit u
                    const int tl = u >> 1, h = u & 1;
                    const uint32_t stg = tl % 3, ph = (tl / 3) & 1, eb = u % 3;
                    if (h == 0) { mbar_wait(BAR(FULL + stg), ph); tc_fence_after(); }
                    const uint32_t t_lo = ((t_smem + stg * 16384) >> 4) | LO32, w_lo = ((w1_smem + h * 16384) >> 4) | LO32;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_ss_tf32_lohi(tmem + eb * 128, t_lo + 2 * ks, w_lo + 2 * ks, HI32, IDESC1, ks > 0);
                    if (MODE != 1)                  // + ones . [be_hi | be_lo]^T
                        umma_ss_tf32_lohi(tmem + eb * 128, (ones_smem >> 4) | LO32, ((biasb_smem + h * 16384) >> 4) | LO32, HI32, IDESC1, 1u);
                    if (WSPLIT) {
                        const uint32_t wl_lo = ((w1l_smem + h * 16384) >> 4) | LO32;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) umma_ss_tf32_lohi(tmem + eb * 128, t_lo + 2 * ks, wl_lo + 2 * ks, HI32, IDESC1, 1u);
                    }
                    if (h == 1) umma_commit(BAR(EMPTY + stg));
                    umma_commit(BAR(EFULL + u % NEF));
                }
                if (u >= 1) {                       // MMA2 of unit u - 1
                    const int v = u - 1, tl = v >> 1, h = v & 1;
                    const uint32_t eb = v % 3, db = tl & 1;
                    mbar_wait(BAR(EREADY + eb), (v / 3) & 1);
                    tc_fence_after();
                    if (h == 0) { mbar_wait(BAR(DFREE + db), ((tl >> 1) & 1) ^ 1); tc_fence_after(); }
#pragma unroll
                    for (int ks = 0; ks < 16; ++ks) {
                        const uint64_t bdesc = smem_desc(HI, w2_smem + (4 * h + (ks >> 2)) * 4096) + 2 * (ks & 3);
                        umma_ts<true>(tmem + 384 + 32 * db, tmem + eb * 128 + ks * 8, bdesc, IDESC2, (h > 0 || ks > 0) ? 1u : 0u);
                        if (WSPLIT) {
                            const uint64_t bl = smem_desc(HI, w2l_smem + (4 * h + (ks >> 2)) * 4096) + 2 * (ks & 3);
                            umma_ts<true>(tmem + 384 + 32 * db, tmem + eb * 128 + ks * 8, bl, IDESC2, 1u);
                        }
                    }
                    if (h == 1) umma_commit(BAR(DFULL + db));
                }
            }
        }
      }
    } else {
        if (RPG == 3) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        const int q = warp & 3;
        const int grp = (warp - EW0) >> 2;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        pdl_wait();
        // bwd: the ReLU bits are needed by the very first chunk of a tile, so they are fetched one tile ahead (the row-wide
        // residual is only needed at the end of the tile and is requested at its start)
        auto mask_row = [&](int tl2) -> const uint4* {
            const int tile2 = blockIdx.x + tl2 * gridDim.x;
            const int b2 = tile2 / a.tiles_per_patch, j2 = tile2 % a.tiles_per_patch;
            return reinterpret_cast<const uint4*>(a.mask + (a.g.lead + (long long)b2 * a.g.pstride + a.g.row0 + j2 * 128 + q * 32 + lane) * 8);
        };
        uint4 nlo = make_uint4(0u, 0u, 0u, 0u), nhi = nlo;
        if (MODE == 1 && grp < my_tiles) { const uint4* mp = mask_row(grp); nlo = __ldg(mp); nhi = __ldg(mp + 1); }
        for (int tl = grp; tl < my_tiles; tl += RPG) {
            const int tile = blockIdx.x + tl * gridDim.x;
            const int b = tile / a.tiles_per_patch, j = tile % a.tiles_per_patch;
            const int r = a.g.row0 + j * 128 + q * 32 + lane;
            const bool in_patch = r < a.g.row0 + a.g.nrows && r < a.g.pstride;
            const bool valid = in_patch && row_valid(a.g, r);
            const long long orow = a.g.lead + (long long)b * a.g.pstride + r;
            // ReLU bits of this row: word (h, c) covers expanded channels h*128 + c*32 .. +31, element e at bit 31 - e,
            // 1 = activation positive.  fwd builds them with one funnel shift per element, bwd tests them.
            uint4 mlo = make_uint4(0u, 0u, 0u, 0u), mhi = mlo;
            float4 pre_r[8];
            // the warp's 32 rows are contiguous in global memory: row-wide traffic goes through the coalescing helpers (rowio.cuh)
            const uint32_t rowmask = __ballot_sync(0xffffffffu, in_patch);
            const long long orow_w = orow - lane;
            uint8_t* const sc = io_scratch + (warp - EW0) * ROWIO_SCRATCH_BYTES;
            if (MODE == 1) {
                mlo = nlo; mhi = nhi;
                if (tl + RPG < my_tiles) { const uint4* mp = mask_row(tl + RPG); nlo = __ldg(mp); nhi = __ldg(mp + 1); }
                if (a.residual) rowio_ldg_chunks(a.residual + orow_w * 32, rowmask, pre_r);
                else {
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) pre_r[g4] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {                     // rolled on purpose: the unrolled body thrashes the instruction cache
                const int u = 2 * tl + h;
                const uint32_t eb = u % 3, hb = lane_base + eb * 128;
                uint4 w4 = h ? mhi : mlo;
                uint32_t wd[4] = {w4.x, w4.y, w4.z, w4.w};
                mbar_wait(BAR(EFULL + u % NEF), (u / NEF) & 1);
                tc_fence_after();
                uint32_t va[32], vb[32];
                tmem_ld32(hb, va);
#pragma unroll
                for (int c = 0; c < 4; ++c) {                 // chunk c lives in va (c even) or vb (c odd); the next chunk's load is
                    tmem_ld_wait();                           // issued before this one is processed
                    uint32_t (&cur)[32] = (c & 1) ? vb : va;
                    uint32_t (&nxt)[32] = (c & 1) ? va : vb;
                    if (c < 3) tmem_ld32(hb + (c + 1) * 32, nxt);
                    if (MODE == 0) {
                        // four independent 8-element sign chains (one serial 32-long funnel-shift chain would pace the warp)
                        uint32_t sg[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                        for (int e4 = 0; e4 < 8; ++e4) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                // x = relu(y) >= +0; (bits(x) - 1) has its sign bit set exactly when x == 0, i.e. when y <= 0,
                                // which is tf.nn.relu's gradient convention (0 at y == 0).  (The bias is already in the accumulator.)
                                const uint32_t x = __float_as_uint(fmaxf(__uint_as_float(cur[e4 * 4 + e]), 0.f));
                                sg[e4 >> 1] = __funnelshift_l(x - 1u, sg[e4 >> 1], 1);
                                cur[e4 * 4 + e] = tf32_bump(x);
                            }
                        }
                        wd[c] = ~((sg[0] << 24) | ((sg[1] & 0xffu) << 16) | ((sg[2] & 0xffu) << 8) | (sg[3] & 0xffu));
                    } else if (MODE == 2) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) cur[e] = tf32_bump(__float_as_uint(fmaxf(__uint_as_float(cur[e]), 0.f)));
                    } else {
                        const uint32_t bits = wd[c];
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            cur[e] = (bits & (0x80000000u >> e)) ? tf32_bump(cur[e]) : 0u;
                    }
                    tmem_st32(hb + c * 32, cur);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(EREADY + eb));
                if (MODE == 0) { w4 = make_uint4(wd[0], wd[1], wd[2], wd[3]); if (h) mhi = w4; else mlo = w4; }
            }
            if (MODE == 0 && a.mask && in_patch) {
                uint4* mp = reinterpret_cast<uint4*>(a.mask + orow * 8);
                mp[0] = mlo;
                mp[1] = mhi;
            }
            const uint32_t db = tl & 1;
            mbar_wait(BAR(DFULL + db), (tl >> 1) & 1);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld32(lane_base + 384 + 32 * db, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(DFREE + db));
            if (MODE == 1 && a.residual) {
                float4 t[8];
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) t[g4] = pre_r[g4];
                rowio_rows_from_chunks(t, pre_r, sc);
            }
            {
                float out_row[32];
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) {
                    float* o = out_row + 4 * g4;
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = __uint_as_float(v[g4 * 4 + e]);
                    if (MODE != 1) {
                        const float4 bq = reinterpret_cast<const float4*>(s_b2)[g4];
                        o[0] += bq.x; o[1] += bq.y; o[2] += bq.z; o[3] += bq.w;
                    } else {
                        o[0] += pre_r[g4].x; o[1] += pre_r[g4].y; o[2] += pre_r[g4].z; o[3] += pre_r[g4].w;
                        if (a.relumask && in_patch) {           // only the first block flows into a ReLU (mainConv1): not worth prefetch registers
                            const float4 mq = __ldg(reinterpret_cast<const float4*>(a.relumask + orow * 32) + g4);
                            if (!(mq.x > 0.f)) o[0] = 0.f;
                            if (!(mq.y > 0.f)) o[1] = 0.f;
                            if (!(mq.z > 0.f)) o[2] = 0.f;
                            if (!(mq.w > 0.f)) o[3] = 0.f;
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (!valid) o[e] = 0.f;
                }
                if (MODE == 1 && a.out_pack) {
                    float pk[32];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(out_row[2 * c], out_row[2 * c + 1]);
                        const float2 hf = __bfloat1622float2(h);
                        const __nv_bfloat162 l = __floats2bfloat162_rn(out_row[2 * c] - hf.x, out_row[2 * c + 1] - hf.y);
                        pk[c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
                        pk[16 + c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&l));
                    }
                    rowio_store_rows(a.out_pack + orow_w * 32, pk, rowmask, sc);
                }
                if (a.round_tf32) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) out_row[c] = rna_tf32(out_row[c]);
                }
                if (MODE == 2 && a.out_f16) {                 // tf32-exact values are fp16-exact inside fp16's normal range
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const __half2 h = __floats2half2_rn(out_row[2 * c], out_row[2 * c + 1]);
                        out_row[c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
                    }
#pragma unroll
                    for (int c = 16; c < 32; ++c) out_row[c] = 0.f;
                }
                rowio_store_rows(a.out + orow_w * 32, out_row, rowmask, sc);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

// ----------------------------------------------------------------------------------------------------------------
//   backward, weight path   dWd = E^T gD,  dWe = X^T gZ,  dbe = sum_rows gZ,  dbd = sum_rows gD     resfront_bwd_weight_kernel
//       The reductions run over rows, so the expanded tensors are produced TRANSPOSED -- channels on the 128 TMEM lanes,
//       rows along the columns -- which makes them legal TS-mode A operands ([M = channel lanes] x [K = rows]):
//         MMA (SS)  E_h^T  = We_h . X^T     (A = We^T rows of this half, B = X tile, both K-major)      cols [0,128)
//         MMA (SS)  gE_h^T = Wd_h . gD^T    (A = Wd rows of this half,  B = gD tile)                    cols [128,256)
//         epilogue  E_h^T <- tf32(relu(. + be)),  gZ_h^T <- tf32(gE_h^T where E > 0)  in place;  dbe += row sums of gZ^T
//         MMA (TS)  dWd_h[ch x co] += E_h^T  . gD   (B = gD tile read MN-major: 32B-atom swizzled copy)  cols [256+32h, +32)
//         MMA (TS)  dWe_h[ch x ci] += gZ_h^T . X    (B = X tile read MN-major)                            cols [320+32h, +32)
//       The four [128 x 32] accumulators stay in TMEM for the CTA's lifetime; per-CTA partials are reduced afterwards.
//       Pipelining unit = (half h of the channels, 64-row sub-tile s): E^T and gE^T of a unit take 64 + 64 TMEM columns;
//       three unit buffers (384 columns) + the four accumulators (128 columns) fill TMEM exactly.  As in the forward
//       kernel the MMA thread issues the two SS MMAs of unit u+1 before the two TS MMAs of unit u, and two epilogue
//       groups of four warps take alternate units.
// three groups: 512 threads = {producer, MMA, two idle warps} + 12 epilogue warps with setmaxnreg (see respipe_threads)
constexpr int RBW_EW0 = RESBW_GROUPS == 3 ? 4 : 2;          // first epilogue warp
constexpr int RBW_THREADS = 32 * RBW_EW0 + 128 * RESBW_GROUPS;
struct ResBwdWeightArgs {
    int B, tiles_per_patch;
    RowGeom g;
    const float* bias_e;
    const uint32_t* mask_t;            // nullable: [tile][4][256] transposed ReLU bits of the forward pass (precision 4)
    float* partials;                   // [cta][4][128][32]: dWd half 0, half 1, dWe^T half 0, half 1
    float* db_partials;                // [cta][groups][256 (dbe)] then [cta][groups x 4 warps][32 (dbd)]
};

template <bool FWD_MASK>     // FWD_MASK: ReLU decisions come from the forward pass's transposed bit mask (a.mask_t) instead of sign(E)
__global__ void __launch_bounds__(RBW_THREADS, 1)
resfront_bwd_weight_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_gd,
                           const __grid_constant__ CUtensorMap tm_x32, const __grid_constant__ CUtensorMap tm_gd32,
                           const __grid_constant__ CUtensorMap tm_weT, const __grid_constant__ CUtensorMap tm_wd,
                           const ResBwdWeightArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[12];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_be[256];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t weT_smem = base;                 // We^T [256 rows(ch) x 32 ci]   A operand of E^T
    const uint32_t wd_smem = base + 32768;          // Wd   [256 rows(ch) x 32 co]   A operand of gE^T
    const uint32_t st_smem = base + 65536;          // 2 stages x { X (K-major), gD (K-major), X (MN), gD (MN) } x 16 KB
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto BAR = [&](int i) { return smem_u32(&bars[i]); };
    const int FULL = 0, EMPTY = 2, WBAR = 4, EFULL = 5, EREADY = 8, DONE = 11;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(FULL + i), 1); mbar_init(BAR(EMPTY + i), 1 + 8); }
        for (int i = 0; i < 3; ++i) { mbar_init(BAR(EFULL + i), 1); mbar_init(BAR(EREADY + i), 4); }
        mbar_init(BAR(WBAR), 1); mbar_init(BAR(DONE), 1);
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < 256; i += RBW_THREADS) s_be[i] = a.bias_e[i];
    if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;                // unit buffers at columns 0 / 128 / 256 (E^T | gE^T), accumulators at 384 + 32 k
    const int ntiles = a.B * a.tiles_per_patch;
    const int t_lo = (int)((long long)ntiles * blockIdx.x / gridDim.x), t_hi = (int)((long long)ntiles * (blockIdx.x + 1) / gridDim.x);
    const int my_tiles = t_hi - t_lo;

    if (warp < RBW_EW0) {
      if (RESBW_GROUPS == 3) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");       // the whole first warpgroup (two of its warps are idle)
      if (warp == 0) {
        if (elect_one_sync()) {
            mbar_arrive_expect_tx(BAR(WBAR), 65536);
            tma_load_2d(weT_smem, &tm_weT, BAR(WBAR), 0, 0);
            tma_load_2d(wd_smem, &tm_wd, BAR(WBAR), 0, 0);
            pdl_wait();
            pdl_trigger();
            for (int tl = 0; tl < my_tiles; ++tl) {
                const int tile = t_lo + tl;
                const int b = tile / a.tiles_per_patch, j = tile % a.tiles_per_patch;
                const int row0 = (int)(a.g.lead + (long long)b * a.g.pstride + a.g.row0 + j * 128);
                const uint32_t stg = tl & 1, ph = (tl >> 1) & 1, sa = st_smem + stg * 65536;
                mbar_wait(BAR(EMPTY + stg), ph ^ 1);
                mbar_arrive_expect_tx(BAR(FULL + stg), 65536);
                tma_load_2d(sa, &tm_x, BAR(FULL + stg), 0, row0);
                tma_load_2d(sa + 16384, &tm_gd, BAR(FULL + stg), 0, row0);
                tma_load_2d(sa + 32768, &tm_x32, BAR(FULL + stg), 0, row0);
                tma_load_2d(sa + 49152, &tm_gd32, BAR(FULL + stg), 0, row0);
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            constexpr uint64_t HI = smem_desc_hi(16, 1024, 2);            // K-major SW128
            constexpr uint32_t HI32 = (uint32_t)(HI >> 32), LO32 = (uint32_t)HI;
            constexpr uint64_t HI_MN = smem_desc_hi(128, 512, 1);         // MN-major, 128B swizzle / 32B atom
            constexpr uint32_t IDESC1 = instr_desc(2, 128, 64, 0, 0);     // [128 ch] x [N = 64 rows]
            constexpr uint32_t IDESC2 = instr_desc(2, 128, 32, 0, 1);     // TS: A from TMEM, B MN-major, N = 32
            mbar_wait(BAR(WBAR), 0);
            tc_fence_after();
            // the SS MMAs run TWO units ahead of the TS MMAs: the units alternate between the epilogue groups, so while group A's unit
            // u - 2 and group B's unit u - 1 are in the epilogue, A's next unit u is already computed in the third buffer (one ahead, each
            // group waited for its next unit's SS MMAs after every unit)
            const int U = 4 * my_tiles;
            for (int u = 0; u < U + 2; ++u) {
                if (u < U) {                        // the two SS MMAs of unit u = (tile, h, s)
                    const int tl = u >> 2, h = (u >> 1) & 1, sub = u & 1;
                    const uint32_t stg = tl & 1, ph = (tl >> 1) & 1, sa = st_smem + stg * 65536, eb = u % 3;
                    if ((u & 3) == 0) { mbar_wait(BAR(FULL + stg), ph); tc_fence_after(); }
                    const uint32_t x_lo = ((sa + sub * 8192) >> 4) | LO32, g_lo = ((sa + 16384 + sub * 8192) >> 4) | LO32;
                    const uint32_t w1 = ((weT_smem + h * 16384) >> 4) | LO32, w2 = ((wd_smem + h * 16384) >> 4) | LO32;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_ss_tf32_lohi(tmem + eb * 128, w1 + 2 * ks, x_lo + 2 * ks, HI32, IDESC1, ks > 0);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_ss_tf32_lohi(tmem + eb * 128 + 64, w2 + 2 * ks, g_lo + 2 * ks, HI32, IDESC1, ks > 0);
                    umma_commit(BAR(EFULL + eb));
                }
                if (u >= 2) {                       // the two TS MMAs of unit u - 2
                    const int v = u - 2, tl = v >> 2, h = (v >> 1) & 1, sub = v & 1;
                    const uint32_t stg = tl & 1, sa = st_smem + stg * 65536, eb = v % 3;
                    mbar_wait(BAR(EREADY + eb), (v / 3) & 1);
                    tc_fence_after();
                    const uint64_t x32 = smem_desc(HI_MN, sa + 32768 + sub * 8192), g32 = smem_desc(HI_MN, sa + 49152 + sub * 8192);
                    const uint32_t first = (tl == 0 && sub == 0) ? 0u : 1u;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)        // K = 8 rows per step: 1024 B further into the MN-major tiles
                        umma_ts<true>(tmem + 384 + 32 * h, tmem + eb * 128 + ks * 8, g32 + 64 * ks, IDESC2, (first | (ks > 0)) ? 1u : 0u);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_ts<true>(tmem + 448 + 32 * h, tmem + eb * 128 + 64 + ks * 8, x32 + 64 * ks, IDESC2, (first | (ks > 0)) ? 1u : 0u);
                    if ((v & 3) == 3) umma_commit(BAR(EMPTY + stg));      // last unit of the tile: the stage may be refilled
                }
            }
            umma_commit(BAR(DONE));
        }
      }
    } else {
        if (RESBW_GROUPS == 3) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        const int q = warp & 3;
        const int grp = (warp - RBW_EW0) >> 2;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        float dbe0 = 0.f, dbe1 = 0.f, dbd = 0.f;
        const int U = 4 * my_tiles;
        pdl_wait();                                          // the partial buffers may still be read by the previous reduction
        // FWD_MASK: the unit's two mask words (32 rows each) for this thread's channel, fetched one unit ahead (the load's L2 / HBM
        // latency is of the order of a unit's whole processing time)
        auto mask_words = [&](int u2, uint32_t (&w)[2]) {
            const int tl2 = u2 >> 2, h2 = (u2 >> 1) & 1, sub2 = u2 & 1;
            const uint32_t* mp = a.mask_t + ((size_t)(t_lo + tl2) * 4 + sub2 * 2) * 256 + h2 * 128 + q * 32 + lane;
            w[0] = __ldg(mp); w[1] = __ldg(mp + 256);
        };
        uint32_t nmt[2] = {0u, 0u};
        if (FWD_MASK && grp < U) mask_words(grp, nmt);
#pragma unroll 1
        for (int u = grp; u < U; u += RESBW_GROUPS) {      // with three groups unit u always lives in H buffer u % 3 == grp
            const int tl = u >> 2, h = (u >> 1) & 1, sub = u & 1;
            const uint32_t stg = tl & 1, ph = (tl >> 1) & 1, eb = u % 3;
            if (h == 0) {   // dbd: column sums of this group's 64 rows of the gD tile (K-major SW128 copy: 16-byte chunk ^= row & 7)
                mbar_wait(BAR(FULL + stg), ph);
                const uint8_t* gp = smem_raw + (st_smem - smem_u32(smem_raw)) + stg * 65536 + 16384;
                float sacc = 0.f;
                for (int r = sub * 64 + q * 16; r < sub * 64 + q * 16 + 16; ++r)
                    sacc += *reinterpret_cast<const float*>(gp + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
                dbd += sacc;
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(EMPTY + stg));
            }
            mbar_wait(BAR(EFULL + eb), (u / 3) & 1);
            tc_fence_after();
            const float be = s_be[h * 128 + q * 32 + lane];           // this thread's channel
            float zs[4] = {0.f, 0.f, 0.f, 0.f};               // four independent chains for the bias-gradient row sum
            uint32_t mt[2] = {nmt[0], nmt[1]};
            if (FWD_MASK && u + RESBW_GROUPS < U) mask_words(u + RESBW_GROUPS, nmt);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t e[32], v[32];
                tmem_ld32(lane_base + eb * 128 + c * 32, e);
                tmem_ld32(lane_base + eb * 128 + 64 + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const float x = fmaxf(__uint_as_float(e[k]) + be, 0.f);
                    // tf.nn.relu's gradient convention: 0 at E == 0.  With the forward's mask the value and the mask may disagree
                    // on elements within rounding distance of zero; the mask is what the data path used, so it decides gZ.
                    e[k] = tf32_bump(__float_as_uint(x));
                    if (FWD_MASK) {
                        // one predicate per element from the mask word (LOP3 with an immediate), shared by both selects; left to itself the
                        // compiler builds an all-ones / zero word per element instead (shift left, arithmetic shift right, AND: +2 instructions)
                        uint32_t vb;
                        float zt;
                        asm("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %2, %3;\nsetp.ne.u32 p, t, 0;\nselp.b32 %0, %4, 0, p;\nselp.f32 %1, %5, 0f00000000, p;\n}\n"
                            : "=r"(vb), "=f"(zt) : "r"(mt[c]), "r"(1u << k), "r"(tf32_bump(v[k])), "f"(__uint_as_float(v[k])));
                        zs[k & 3] += zt;                                   // bias gradient from the unrounded value
                        v[k] = vb;
                    } else {
                        const bool pos = x > 0.f;
                        zs[k & 3] += pos ? __uint_as_float(v[k]) : 0.f;
                        v[k] = pos ? tf32_bump(v[k]) : 0u;
                    }
                }
                tmem_st32(lane_base + eb * 128 + c * 32, e);
                tmem_st32(lane_base + eb * 128 + 64 + c * 32, v);
            }
            const float zsum = (zs[0] + zs[1]) + (zs[2] + zs[3]);
            if (h) dbe1 += zsum; else dbe0 += zsum;
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(EREADY + eb));
        }
        float* dbp = a.db_partials + (size_t)blockIdx.x * RESBW_DBP;
        dbp[grp * 256 + q * 32 + lane] = dbe0;
        dbp[grp * 256 + 128 + q * 32 + lane] = dbe1;
        dbp[RESBW_GROUPS * 256 + (grp * 4 + q) * 32 + lane] = dbd;
        mbar_wait(BAR(DONE), 0);
        tc_fence_after();
        float* out = a.partials + (size_t)blockIdx.x * 4 * 4096;
#pragma unroll 1
        for (int g2 = 0; g2 < 2 && grp < 2; ++g2) { // group 0 drains the dWd accumulators, group 1 the dWe^T ones
            const int g = grp * 2 + g2;
            uint32_t v[32];
            tmem_ld32(lane_base + 384 + g * 32, v);
            tmem_ld_wait();
            float4* o = reinterpret_cast<float4*>(out + ((size_t)g * 128 + q * 32 + lane) * 32);
#pragma unroll
            for (int e = 0; e < 8; ++e)
                o[e] = my_tiles > 0 ? make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

// immediate (non-deferred) reduction of the per-CTA partials (wgrad_reduce.cuh has the body)
__global__ void __launch_bounds__(256)
resfront_reduce_kernel(const float* __restrict__ partials, const float* __restrict__ dbp, int ncta,
                       float* __restrict__ dwd, float* __restrict__ dwe, float* __restrict__ dbe, float* __restrict__ dbd) {
    __shared__ float4 sm[256];
    tc::pdl_wait();
    resfront_reduce_body(blockIdx.x, partials, dbp, ncta, dwd, dwe, dbe, dbd, sm);
}

template <int MODE, bool WSPLIT = false>
static int launch_respipe(const float* t, const float* w1, const float* w2, const ResPipeArgs& a0, const char* tag, double flops, cudaStream_t st,
                          const float* w1_lo = nullptr, const float* w2_lo = nullptr) {
    ResPipeArgs a = a0;
    a.tiles_per_patch = cdiv(a.g.nrows, 128);
    const long long rows = a.g.lead + (long long)a.B * a.g.pstride + ROW_TAIL;
    CUtensorMap tm_t, tm_w1, tm_w2;
    PV_TRY(make_tmap_2d(&tm_t, t, rows, 32, 128, 32, 0));
    PV_TRY(make_tmap_2d(&tm_w1, w1, 256, 32, 256, 32, 0));      // [256 rows][32]: We^T (fwd) | Wd (bwd)
    PV_TRY(make_tmap_2d(&tm_w2, w2, 32, 256, 32, 32, 0));       // [32 rows][256]: Wd^T (fwd) | We (bwd)
    CUtensorMap tm_w1l = tm_w1, tm_w2l = tm_w2;
    if (WSPLIT) {
        if (!w1_lo || !w2_lo) return set_error(PV_ERR_BAD_ARG, "resfront: split weights requested without the lo matrices");
        PV_TRY(make_tmap_2d(&tm_w1l, w1_lo, 256, 32, 256, 32, 0));
        PV_TRY(make_tmap_2d(&tm_w2l, w2_lo, 32, 256, 32, 32, 0));
    }
    const size_t smem = 1024 + (WSPLIT ? 131072 : 65536) + 3 * 16384 + 4 * respipe_groups(MODE) * ROWIO_SCRATCH_BYTES +
                        (MODE != 1 ? 16384 + 32768 : 0);        // forward: the ones / bias tiles of the bias MMA
    static size_t attr[16] = {};
    PV_CUDA(ensure_dyn_smem(resfront_pipe_kernel<MODE, WSPLIT>, smem, attr));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = a.B * a.tiles_per_patch;
    const int grid = ntiles < sms ? ntiles : sms;
    PV_TIMED(tag, st, flops, 0.0, (WSPLIT ? 2.0 : 1.0) * 2.0 * 2.0 * (double)ntiles * 128.0 * 32.0 * 256.0);
    PV_CUDA(launch_pdl(resfront_pipe_kernel<MODE, WSPLIT>, grid, respipe_threads(respipe_groups(MODE)), smem, st, tm_t, tm_w1, tm_w2, tm_w1l, tm_w2l, a));
    PV_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// D = decConv(relu(expConv(X))) on PR rows.  weT_exp [256][32], weT_dec [32][256], biases padded to 256 / 32.
// relu_bits (nullable): [rows][8] uint32, bit (c % 32) of word c / 32 = (E[row][c] > 0), consumed by the backward-data kernel.
int launch_resfront_fwd_tc(const float* x, const float* weT_exp, const float* weT_dec, const float* bias_e, const float* bias_d,
                           float* d, uint32_t* relu_bits, const RowGeom& g, int B, int round_tf32, double flops, cudaStream_t st, int out_f16) {
    ResPipeArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.g = g; a.bias1 = bias_e; a.bias2 = bias_d; a.mask = relu_bits; a.out = d; a.round_tf32 = round_tf32;
    if (out_f16 && (relu_bits || !round_tf32)) return set_error(PV_ERR_BAD_ARG, "resfront_fwd: fp16 rows are an inference output of tf32-rounded values");
    a.out_f16 = out_f16;
    if (!relu_bits) return launch_respipe<2>(x, weT_exp, weT_dec, a, "resfront_fwd_infer", flops, st);
    return launch_respipe<0>(x, weT_exp, weT_dec, a, "resfront_fwd", flops, st);
}

// gA = ((gD Wd^T) .* relu_bits) We + G  (.* relumask).  w_dec [256][32] (= weff of decConv), w_exp [32][256] (= weff of expConv)
int launch_resfront_bwd_data_tc(const float* gd, const float* w_dec, const float* w_exp, const uint32_t* relu_bits,
                                const float* residual, const float* relumask, float* ga, const RowGeom& g,
                                int B, int round_tf32, double flops, cudaStream_t st, const float* w_dec_lo, const float* w_exp_lo, float* ga_pack) {
    ResPipeArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.g = g; a.mask = const_cast<uint32_t*>(relu_bits); a.residual = residual; a.relumask = relumask; a.out = ga; a.round_tf32 = round_tf32;
    a.out_pack = ga_pack;
    if (w_dec_lo || w_exp_lo) return launch_respipe<1, true>(gd, w_dec, w_exp, a, "resfront_bwd_data_w2", flops, st, w_dec_lo, w_exp_lo);
    return launch_respipe<1>(gd, w_dec, w_exp, a, "resfront_bwd_data", flops, st);
}

}  // namespace pv

namespace pv {
// weight / bias gradients of expConv and decConv of one block, E and gZ recomputed on chip
int launch_resfront_bwd_weight_tc(const float* x, const float* gd, const float* weT_exp, const float* w_dec, const float* bias_e,
                                  float* dw_dec, float* dw_exp, float* db_exp, float* db_dec, const RowGeom& g, int B,
                                  float* partials, size_t partial_floats, double flops, cudaStream_t st, ReduceQueue* rq,
                                  const uint32_t* relu_bits_t) {
    ResBwdWeightArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.tiles_per_patch = cdiv(g.nrows, 128); a.g = g; a.bias_e = bias_e; a.mask_t = relu_bits_t;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = a.B * a.tiles_per_patch;
    const int grid = ntiles < sms ? ntiles : sms;
    const size_t need = (size_t)grid * (4 * 4096 + RESBW_DBP);
    float* deferred = rq ? rq->take(need) : nullptr;
    if (deferred) { partials = deferred; partial_floats = need; }
    if (!partials || partial_floats < need) return set_error(PV_ERR_BAD_ARG, "resfront_bwd_weight: partial buffer too small");
    a.partials = partials; a.db_partials = partials + (size_t)grid * 4 * 4096;
    const long long rows = g.lead + (long long)B * g.pstride + ROW_TAIL;
    CUtensorMap tm_x, tm_gd, tm_x32, tm_gd32, tm_weT, tm_wd;
    PV_TRY(make_tmap_2d(&tm_x, x, rows, 32, 128, 32, 0));
    PV_TRY(make_tmap_2d(&tm_gd, gd, rows, 32, 128, 32, 0));
    PV_TRY(make_tmap_2d(&tm_x32, x, rows, 32, 128, 32, 1));
    PV_TRY(make_tmap_2d(&tm_gd32, gd, rows, 32, 128, 32, 1));
    PV_TRY(make_tmap_2d(&tm_weT, weT_exp, 256, 32, 256, 32, 0));
    PV_TRY(make_tmap_2d(&tm_wd, w_dec, 256, 32, 256, 32, 0));
    const size_t smem = 1024 + 65536 + 2 * 65536;
    static size_t attr[16] = {}, attr_m[16] = {};
    if (relu_bits_t) PV_CUDA(ensure_dyn_smem(resfront_bwd_weight_kernel<true>, smem, attr_m));
    else PV_CUDA(ensure_dyn_smem(resfront_bwd_weight_kernel<false>, smem, attr));
    {
        // algorithmic: dWe + dWd.  Executed: E^T and gE^T are recomputed on chip, and every GEMM runs on the padded 32 x 256 shapes
        // over whole 128-row tiles: 4 GEMMs of 2 * rows * 32 * 256 flops.
        PV_TIMED("resfront_bwd_weight", st, flops, 0.0, 4.0 * 2.0 * (double)ntiles * 128.0 * 32.0 * 256.0);
        if (relu_bits_t) PV_CUDA(launch_pdl(resfront_bwd_weight_kernel<true>, grid, RBW_THREADS, smem, st, tm_x, tm_gd, tm_x32, tm_gd32, tm_weT, tm_wd, a));
        else PV_CUDA(launch_pdl(resfront_bwd_weight_kernel<false>, grid, RBW_THREADS, smem, st, tm_x, tm_gd, tm_x32, tm_gd32, tm_weT, tm_wd, a));
        PV_LAUNCH_CHECK();
    }
    if (deferred) {
        ReduceJob j;
        memset(&j, 0, sizeof j);
        j.kind = 1; j.nblocks = RESFRONT_REDUCE_BLOCKS; j.partials = a.partials; j.dbp = a.db_partials; j.ncta = grid;
        j.dwd = dw_dec; j.dwe = dw_exp; j.dbe = db_exp; j.dbd = db_dec;
        rq->push(j);
    } else {
        PV_TIMED("wgrad_reduce", st);
        PV_CUDA(launch_pdl(resfront_reduce_kernel, RESFRONT_REDUCE_BLOCKS, 256, 0, st, (const float*)a.partials, (const float*)a.db_partials, grid, dw_dec, dw_exp, db_exp, db_dec));
        PV_LAUNCH_CHECK();
    }
    return 0;
}
}  // namespace pv
