// conv3_tc.cu -- the 3x3x3, 32 -> 32 channel convolution (normConv forward / data gradient, convReducer, reference
// models/modelsTF.py:159-163,185-188 and their Conv3DBackpropInputV2) as an implicit GEMM whose MMA N dimension is
// widened from 32 to 96 by folding the three dw taps into N.
//
// Why: on sm_100a an M128 x N x K8 kind::tf32 MMA with both operands in shared memory costs max(N/2, 32 + N/4) cycles
// (probes/umma_rate.cu, profiles/r02_umma_rate_probe.log: the operands stream from shared memory at 128 B/cycle, 4 KB of A per
// MMA whatever N is), so with N = 32 the A fetch alone paces the tensor pipe.  Taps that differ only in dw read the SAME activation
// rows shifted by one row, so
//     Q_j[rho] = sum_{g=(dt,dh)} X[rho + base_g + 1] . W[g, dw = j]          j = 0,1,2      (one N = 96 MMA chain, 9 groups)
//     out[r]   = Q_0[r - 1] + Q_1[r] + Q_2[r + 1]
// i.e. 36 MMAs of 56 cycles per 128-row tile instead of 108 MMAs of 40 cycles.  The row shift of the two outer thirds
// is done by the epilogue: TMEM lanes are rows, so it is a lane shift -- warp shuffles inside a warp, a 1 KB shared
// memory exchange across the four warps of a TMEM lane quarter group; the tile's first/last lanes are halo.
//
// Tiling.  PR layout (rows.h; 'same' convolutions): a plane is 22 lines of 23 rows (column 22 = zero padding).  Tiles
// start at plane rows 0, 126, 252, 378: lane 0 of the first tile needs no left halo because row -1 is a padding
// column (Q_0 there is a sum over zero rows), lane 127 of the last tile is row 505, a padding column itself -> exactly
// four tiles per plane, 36 per patch (the N = 32 kernel needed 38).  Tiles are ordered (chunk, patch, t) and every CTA
// owns a contiguous range, so going from plane t to t + 1 re-uses two of the three temporal slabs already in shared
// memory, and going from the last plane of a patch to the first plane of the next re-uses one (the shared zero plane):
// ~1.2 TMA slab loads per tile instead of 3, and the four-stage slab ring never drains inside a CTA's range.  Other
// layouts (G: valid convolutions of the reducers) use flat tiles of 126 output rows with a halo lane on both sides and
// three fresh slabs per tile.
//
// Warp roles as in conv_tc.cu: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer (one elected thread), warps 2-5 and
// 6-9 two epilogue groups draining alternate tiles (TMEM accumulator double buffer, 2 x 96 columns; 2 x 192 in the pair-row modes).
//
// Modes (template parameter, RowConvP::f16_pack): 0 = kind::tf32 on fp32 rows (single-pass engine: forward, data gradient);
// 1 = the error-compensated forward in one launch on fp16 pair rows (main + correction accumulators); 2 = the split-weight data
// gradient in one launch on bf16 pair rows; 3 = kind::f16 on the hi halves of fp16 pair rows (inference).  See the kernel's comment.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "rowio.cuh"
#include "rows.h"
#include "tc_common.cuh"

namespace pv {

using namespace tc;
int make_tmap_2d(CUtensorMap* m, const float* base, long long rows, int cols, int box_rows, int box_cols, int swizzle_32b_atom);

namespace {

constexpr int C3_THREADS = 320;
constexpr int C3_STAGES = 4;
constexpr int C3_TILE = 126;           // output rows per tile (128 TMEM lanes minus the two halo lanes)

struct Conv3Args {
    int B;
    int plane_mode;                    // 1: PR tiling (patch, chunk, t); 0: flat tiles over [row0, row0 + nrows)
    int chunks, nt;                    // plane mode: tiles per plane, planes per patch; flat: tiles per patch, 1
    int plane_rows;                    // plane mode: rows per plane (input and output geometry agree)
    int plane_out_rows;                // plane mode: rows of a plane that may be written (nh * pw)
    int chain_patches;                 // plane mode: pstride == (nt + 1) * plane, so patch b+1 continues patch b's slab chain
    long long in_lead, in_pstride;
    RowGeom og;
    int slab_rows;                     // 128 + 2 * pw rounded up to 8
    int slab_lo[3];                    // first row of temporal slab dtI relative to the tile's lane-0 row
    int pw;                            // rows per image line: dh view stride inside a slab
    int tap_wr[MAX_TAPS], tap_wc[MAX_TAPS];
    const float* bias; const float* residual; const float* residual2; const float* relumask; float* y; float* y_lo;
    float* y_pack;                     // MODE 1: packed fp16 pair rows of (hi, lo) (rows.h PACK_SCALE); MODE 2: bf16 pair rows; nullable
    int relu, round_tf32;
};

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

struct TileInfo { int b, c, t; int r0; int nnew; bool lane0_out; };      // nnew: temporal slabs this tile has to load (the others are its predecessor's)

__device__ __forceinline__ TileInfo tile_info(const Conv3Args& a, int tile, int t_lo) {
    TileInfo ti;
    if (a.plane_mode) {                                      // tile = (c * B + b) * nt + t
        const int per_chunk = a.B * a.nt;
        ti.c = tile / per_chunk;
        const int rem = tile - ti.c * per_chunk;
        ti.b = rem / a.nt; ti.t = rem - ti.b * a.nt;
        ti.r0 = a.og.row0 + ti.t * a.plane_rows + ti.c * C3_TILE;
        ti.nnew = (tile == t_lo || rem == 0) ? 3 : (ti.t == 0 ? (a.chain_patches ? 2 : 3) : 1);
        ti.lane0_out = ti.c == 0;
    } else {
        ti.b = tile / a.chunks; ti.c = tile - ti.b * a.chunks; ti.t = 0;
        ti.r0 = a.og.row0 - 1 + ti.c * C3_TILE;
        ti.nnew = 3;
        ti.lane0_out = false;
    }
    return ti;
}

// F16 (the error-compensated convolution in ONE launch): the operands are packed fp16 pair rows (rows.h), activations
// [hi | 2^12 lo], weights [2^12 w_lo | w_hi].  Per tap the MMA thread issues kind::f16 MMAs (K = 16 per step over the same 128-byte rows)
// into TWO accumulators: the MAIN product x_hi w_hi = the first half of the activation row against the second half of the weight row
// (2 steps), and both CORRECTIONS x_hi w_lo + x_lo w_hi = the whole rows against each other (4 steps, 2^12-scaled).  hi = tf32(v) has 11
// significant bits and is exact in fp16, so the main product loses nothing against kind::tf32 while taking half the accumulation steps
// (tcgen05 truncates the accumulator once per step), and the corrections never touch the full-size sum.  The epilogue adds
// main + 2^-12 corrections (+ bias + skip connection) in fp32, round to nearest.
//
// MODE 2 (the split-weight data gradient in ONE launch): gradients are not O(1), so both operands are bf16 PAIRS (fp32's exponent range,
// 16 significant bits, no scaling): gradient rows [bf16(g) | bf16(g - bf16(g))], weight rows [bf16(w) | bf16(w - bf16(w))] in the data
// gradient's layout.  Per tap: main accumulator g_a w_a (2 steps), second accumulator g_b w_a + g_a w_b (4 steps); same epilogue with
// scale 1.  What is dropped is 2^-16 of each weight (the single-pass engine drops 2^-12, which is the systematic error the split
// removes) and 2^-16 of each gradient (tf32 operands keep 2^-12).  (kind::f16 does NOT take bf16 gradients against fp16 weights:
// mixed 16-bit operand formats raise an illegal-instruction fault on sm_100a, tried in round 2.)
//
// MODE 3 (inference of the single-pass engine): the main product alone from the hi halves of fp16 pair rows -- the values are tf32-exact, so
// the result is the kind::tf32 kernel's, with 2 instead of 4 MMAs per tap row (the kernel is bound by the operand fetch from shared memory).
template <int MODE>
__global__ void __launch_bounds__(C3_THREADS, 1)
rowconv3_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const Conv3Args a) {
    constexpr bool F16 = MODE == 1 || MODE == 2;                  // pair-row modes with two accumulators
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * C3_STAGES + 5];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_bias[32];
    __shared__ __align__(16) float xch[2][2][4][2][32];       // [epilogue group][tile parity][warp][0: last lane's Q_0, 1: first lane's Q_2][channel]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_smem = base;                              // 27 taps x [32 co rows x 128 B], sorted-tap order
    const uint32_t stage_bytes = (uint32_t)a.slab_rows * 128u;
    const uint32_t st_smem = base + 27u * 4096u;
    uint8_t* const io_scratch = smem_raw + (base - smem_u32(smem_raw)) + 27u * 4096u + C3_STAGES * stage_bytes;   // 8 epilogue warps x 2 KB (rowio.cuh)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto BAR = [&](int i) { return smem_u32(&bars[i]); };
    const int FULL = 0, EMPTY = C3_STAGES, TFULL = 2 * C3_STAGES, TEMPTY = 2 * C3_STAGES + 2, WBAR = 2 * C3_STAGES + 4;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C3_STAGES; ++i) { mbar_init(BAR(FULL + i), 1); mbar_init(BAR(EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(TFULL + i), 1); mbar_init(BAR(TEMPTY + i), 4); }
        mbar_init(BAR(WBAR), 1);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) s_bias[threadIdx.x] = a.bias ? a.bias[threadIdx.x] : 0.f;
    constexpr uint32_t ACC_COLS = F16 ? 192u : 96u;               // F16: main accumulator | correction accumulator
    if (warp == 1) tmem_alloc<(F16 ? 512 : 256)>(smem_u32(&tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int ntiles = a.B * a.chunks * a.nt;
    const int t_lo = (int)((long long)ntiles * blockIdx.x / gridDim.x), t_hi = (int)((long long)ntiles * (blockIdx.x + 1) / gridDim.x);

    if (warp == 0) {
        // ================================================================== TMA producer
        if (elect_one_sync()) {
            tma_prefetch_desc(&tm_x);
            tma_prefetch_desc(&tm_w);
            mbar_arrive_expect_tx(BAR(WBAR), 27u * 4096u);
            for (int t = 0; t < 27; ++t) tma_load_2d(w_smem + t * 4096, &tm_w, BAR(WBAR), a.tap_wc[t], a.tap_wr[t]);
            pdl_wait();                                      // activations come from the previous kernel
            pdl_trigger();
            uint32_t n = 0;                                  // slabs loaded so far: slab k lives in stage k % 4
            for (int tile = t_lo; tile < t_hi; ++tile) {
                const TileInfo ti = tile_info(a, tile, t_lo);
                const long long irow0 = a.in_lead + (long long)ti.b * a.in_pstride + ti.r0;
                for (int s = 3 - ti.nnew; s < 3; ++s, ++n) {
                    const uint32_t stg = n % C3_STAGES, ph = (n / C3_STAGES) & 1;
                    mbar_wait(BAR(EMPTY + stg), ph ^ 1);
                    mbar_arrive_expect_tx(BAR(FULL + stg), stage_bytes);
                    tma_load_2d(st_smem + stg * stage_bytes, &tm_x, BAR(FULL + stg), 0, (int)(irow0 + a.slab_lo[s]));
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer (one thread)
        if (elect_one_sync()) {
            constexpr uint64_t HI = smem_desc_hi(16, 1024, 2);          // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
            constexpr uint32_t HI32 = (uint32_t)(HI >> 32), LO32 = (uint32_t)HI;
            // tf32 (MODE 1: fp16, MODE 2: bf16) operands -> f32, M = 128, N = 96 (three dw taps)
            constexpr uint32_t IDESC = instr_desc(MODE == 0 ? 2 : (MODE == 2 ? 1 : 0), 128, 96, 0, 0);   // MODE 1, 3: fp16
            const uint32_t dh_inc = (uint32_t)a.pw * 8u;                // one image line further into the slab (16-byte units)
            mbar_wait(BAR(WBAR), 0);
            tc_fence_after();
            // The issuing thread must stay ahead of the tensor pipe (an N = 96 MMA retires every ~56 cycles), so the per-tile
            // bookkeeping is incremental: (c, b, t) counters instead of divisions, ring stages and barrier parities in registers.
            TileInfo ti = tile_info(a, t_lo, t_lo);
            int tb = ti.b, tt = ti.t;
            uint32_t nnew = 3;
            uint32_t s0 = 0, s1 = 0, s2 = 0, nxt = 0;        // stages of the slabs dtI = 0,1,2 in use; next stage the producer fills
            uint32_t fullph = 0;                             // bit s: parity FULL[s] completes with next
            uint32_t tl = 0;
            for (int tile = t_lo; tile < t_hi; ++tile, ++tl) {
                const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
                for (uint32_t i = 0; i < nnew; ++i) { s0 = s1; s1 = s2; s2 = nxt; nxt = nxt == C3_STAGES - 1 ? 0 : nxt + 1; }
                // what the next tile inherits decides which slabs are released after this one
                uint32_t nnext = 3;
                if (a.plane_mode) {
                    if (++tt == a.nt) { tt = 0; if (++tb == a.B) tb = 0; }
                    nnext = tt != 0 ? 1u : ((tb != 0 && a.chain_patches) ? 2u : 3u);
                }
                if (tile + 1 == t_hi) nnext = 3;
                mbar_wait(BAR(TEMPTY + acc), aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem + acc * ACC_COLS;
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const uint32_t stg = s == 0 ? s0 : (s == 1 ? s1 : s2);
                    if (s >= 3 - (int)nnew) { mbar_wait(BAR(FULL + stg), (fullph >> stg) & 1); fullph ^= 1u << stg; tc_fence_after(); }
                    const uint32_t a_lo = ((st_smem + stg * stage_bytes) >> 4) | LO32;
                    const uint32_t b_lo = ((w_smem >> 4) | LO32) + (uint32_t)(s * 9) * 256u;
#pragma unroll
                    for (int dh = 0; dh < 3; ++dh) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            if (MODE == 2) {
                                // K-steps 0, 1 of a row = the first halves (g_a, w_a), 2, 3 = the remainders: main g_a w_a; second g_b w_a + g_a w_b
                                if (ks < 2) umma_ss_f16_lohi(d_tmem, a_lo + dh * dh_inc + 2 * ks, b_lo + (uint32_t)(dh * 3) * 256u + 2 * ks, HI32, IDESC, (s | dh | ks) ? 1u : 0u);
                                umma_ss_f16_lohi(d_tmem + 96, a_lo + dh * dh_inc + 2 * ((ks + 2) & 3), b_lo + (uint32_t)(dh * 3) * 256u + 2 * ks, HI32, IDESC, (s | dh | ks) ? 1u : 0u);
                            } else if (MODE == 3) {
                                if (ks < 2) umma_ss_f16_lohi(d_tmem, a_lo + dh * dh_inc + 2 * ks, b_lo + (uint32_t)(dh * 3) * 256u + 4 + 2 * ks, HI32, IDESC, (s | dh | ks) ? 1u : 0u);
                            } else if (F16) {
                                // corrections: whole rows; main: activation K-steps 0, 1 (hi) against weight K-steps 2, 3 (w_hi)
                                umma_ss_f16_lohi(d_tmem + 96, a_lo + dh * dh_inc + 2 * ks, b_lo + (uint32_t)(dh * 3) * 256u + 2 * ks, HI32, IDESC, (s | dh | ks) ? 1u : 0u);
                                if (ks < 2) umma_ss_f16_lohi(d_tmem, a_lo + dh * dh_inc + 2 * ks, b_lo + (uint32_t)(dh * 3) * 256u + 4 + 2 * ks, HI32, IDESC, (s | dh | ks) ? 1u : 0u);
                            } else {
                                umma_ss_tf32_lohi(d_tmem, a_lo + dh * dh_inc + 2 * ks, b_lo + (uint32_t)(dh * 3) * 256u + 2 * ks, HI32, IDESC, (s | dh | ks) ? 1u : 0u);
                            }
                        }
                    }
                    // release the slabs the next tile does not inherit: as many (oldest first) as it loads itself
                    if (s < (int)nnext) umma_commit(BAR(EMPTY + stg));
                }
                umma_commit(BAR(TFULL + acc));
                nnew = nnext;
            }
        }
    } else {
        // ================================================================== epilogue: two groups of 4 warps (TMEM lane quarter = warp % 4)
        const int q = warp & 3;
        const uint32_t grp = (uint32_t)(warp - 2) >> 2;
        uint32_t tl = 0;
        pdl_wait();
        for (int tile = t_lo; tile < t_hi; ++tile, ++tl) {
            const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
            if (acc != grp) continue;
            const TileInfo ti = tile_info(a, tile, t_lo);
            const int L = q * 32 + lane;
            const int r = ti.r0 + L;                                      // row inside the patch
            bool in_patch = (L >= 1 || ti.lane0_out) && L <= C3_TILE && r < a.og.row0 + a.og.nrows && r < a.og.pstride;
            if (a.plane_mode) in_patch = in_patch && (ti.c * C3_TILE + L) < a.plane_out_rows;
            const bool valid = in_patch && row_valid(a.og, r);
            const long long orow = a.og.lead + (long long)ti.b * a.og.pstride + r;
            // this warp's 32 rows are contiguous in global memory: all traffic goes through the coalescing helpers of rowio.cuh
            const uint32_t rowmask = __ballot_sync(0xffffffffu, in_patch);
            const long long orow_w = orow - lane;                         // row of this warp's lane 0
            uint8_t* const sc = io_scratch + (warp - 2) * ROWIO_SCRATCH_BYTES;
            // the rows' residual (forward) or ReLU-mask (data gradient) operand, requested before waiting for the accumulator
            const float* pre_src = a.residual ? a.residual : a.relumask;
            float4 pre[8];
            if (pre_src) rowio_ldg_chunks(pre_src + orow_w * 32, rowmask, pre);
            if (MODE == 1 && a.residual2) {                               // the lo half of the skip connection (compensated forward)
                float4 t2[8];
                rowio_ldg_chunks(a.residual2 + orow_w * 32, rowmask, t2);
#pragma unroll
                for (int i = 0; i < 8; ++i) { pre[i].x += t2[i].x; pre[i].y += t2[i].y; pre[i].z += t2[i].z; pre[i].w += t2[i].w; }
            }
            mbar_wait(BAR(TFULL + acc), aph);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS;
            // cross-warp halo: the last lane's Q_0 row goes to the next warp's lane 0, the first lane's Q_2 row to the previous warp's lane 31
            float (*xb)[2][32] = xch[grp][(tl >> 1) & 1];
            float o[32];
            if (F16) {
                // main + 2^-12 x correction accumulator, one dw third at a time (middle, left, right) so that only one third of each
                // accumulator is live next to the running sum: the kernel sits at its 168-register cap
                constexpr float CS = MODE == 2 ? 1.0f : 1.0f / PACK_SCALE;
                uint32_t vm[32], vc[32];
                tmem_ld32(taddr + 32, vm);
                tmem_ld32(taddr + 128, vc);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) o[c] = fmaf(__uint_as_float(vc[c]), CS, __uint_as_float(vm[c]));
                tmem_ld32(taddr, vm);
                tmem_ld32(taddr + 96, vc);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float t = fmaf(__uint_as_float(vc[c]), CS, __uint_as_float(vm[c]));
                    vm[c] = __float_as_uint(t);
                    const float left = __shfl_up_sync(0xffffffffu, t, 1);
                    o[c] += lane > 0 ? left : 0.f;
                }
                if (lane == 31) {
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4)
                        reinterpret_cast<float4*>(xb[q][0])[g4] = make_float4(__uint_as_float(vm[4 * g4]), __uint_as_float(vm[4 * g4 + 1]), __uint_as_float(vm[4 * g4 + 2]), __uint_as_float(vm[4 * g4 + 3]));
                }
                tmem_ld32(taddr + 64, vm);
                tmem_ld32(taddr + 160, vc);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(TEMPTY + acc));
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float t = fmaf(__uint_as_float(vc[c]), CS, __uint_as_float(vm[c]));
                    vm[c] = __float_as_uint(t);
                    const float right = __shfl_down_sync(0xffffffffu, t, 1);
                    o[c] += lane < 31 ? right : 0.f;
                }
                if (lane == 0) {
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4)
                        reinterpret_cast<float4*>(xb[q][1])[g4] = make_float4(__uint_as_float(vm[4 * g4]), __uint_as_float(vm[4 * g4 + 1]), __uint_as_float(vm[4 * g4 + 2]), __uint_as_float(vm[4 * g4 + 3]));
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            } else {
                uint32_t v0[32], v1[32], v2[32];
                tmem_ld32(taddr, v0);
                tmem_ld32(taddr + 32, v1);
                tmem_ld32(taddr + 64, v2);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(TEMPTY + acc));
                if (lane == 31) {
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4)
                        reinterpret_cast<float4*>(xb[q][0])[g4] = make_float4(__uint_as_float(v0[4 * g4]), __uint_as_float(v0[4 * g4 + 1]), __uint_as_float(v0[4 * g4 + 2]), __uint_as_float(v0[4 * g4 + 3]));
                }
                if (lane == 0) {
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4)
                        reinterpret_cast<float4*>(xb[q][1])[g4] = make_float4(__uint_as_float(v2[4 * g4]), __uint_as_float(v2[4 * g4 + 1]), __uint_as_float(v2[4 * g4 + 2]), __uint_as_float(v2[4 * g4 + 3]));
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[c]), 1);
                    const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[c]), 1);
                    o[c] = __uint_as_float(v1[c]) + (lane > 0 ? left : 0.f) + (lane < 31 ? right : 0.f);
                }
            }

            if (lane == 0 && q > 0) {
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) {
                    const float4 h = reinterpret_cast<const float4*>(xb[q - 1][0])[g4];
                    o[4 * g4] += h.x; o[4 * g4 + 1] += h.y; o[4 * g4 + 2] += h.z; o[4 * g4 + 3] += h.w;
                }
            }
            if (lane == 31 && q < 3) {
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) {
                    const float4 h = reinterpret_cast<const float4*>(xb[q + 1][1])[g4];
                    o[4 * g4] += h.x; o[4 * g4 + 1] += h.y; o[4 * g4 + 2] += h.z; o[4 * g4 + 3] += h.w;
                }
            }
            if (pre_src) { float4 t[8]; 
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) t[g4] = pre[g4];
                rowio_rows_from_chunks(t, pre, sc); }
#pragma unroll
            for (int g4 = 0; g4 < 8; ++g4) {
                float* e = o + 4 * g4;
                const float4 bq = reinterpret_cast<const float4*>(s_bias)[g4];
                e[0] += bq.x; e[1] += bq.y; e[2] += bq.z; e[3] += bq.w;
                if (a.residual) { e[0] += pre[g4].x; e[1] += pre[g4].y; e[2] += pre[g4].z; e[3] += pre[g4].w; }
                if (a.relu) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) e[k] = fmaxf(e[k], 0.f);
                }
                if (a.relumask) {
                    const float4 mq = (a.residual && in_patch) ? __ldg(reinterpret_cast<const float4*>(a.relumask + orow * 32) + g4) : pre[g4];
                    e[0] = mq.x > 0.f ? e[0] : 0.f; e[1] = mq.y > 0.f ? e[1] : 0.f;
                    e[2] = mq.z > 0.f ? e[2] : 0.f; e[3] = mq.w > 0.f ? e[3] : 0.f;
                }
                if (!valid) { e[0] = e[1] = e[2] = e[3] = 0.f; }
                if (a.round_tf32 && !(MODE == 2 && a.y_pack)) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) e[k] = rna_tf32(e[k]);
                }
            }
            if (MODE == 2) {
                if (a.y_pack) {                                           // the un-rounded gradient as a bf16 pair row for the next data gradient
                    float pk[32];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * c], o[2 * c + 1]);
                        const float2 hf = __bfloat1622float2(h);
                        const __nv_bfloat162 l = __floats2bfloat162_rn(o[2 * c] - hf.x, o[2 * c + 1] - hf.y);
                        pk[c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
                        pk[16 + c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&l));
                    }
                    rowio_store_rows(a.y_pack + orow_w * 32, pk, rowmask, sc);
                    if (a.round_tf32) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) o[c] = rna_tf32(o[c]);
                    }
                }
                rowio_store_rows(a.y + orow_w * 32, o, rowmask, sc);
            } else if (MODE == 1 && (a.y_lo || a.y_pack)) {               // hi / lo split of the fp32 result (hi is tf32-exact)
                float lo[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) { const float hi = rna_tf32(o[c]); lo[c] = o[c] - hi; o[c] = hi; }
                rowio_store_rows(a.y + orow_w * 32, o, rowmask, sc);
                if (a.y_lo) rowio_store_rows(a.y_lo + orow_w * 32, lo, rowmask, sc);
                if (a.y_pack) {                                           // [ fp16(hi) x 32 | fp16(PACK_SCALE * lo) x 32 ] as 32 words
                    float pk[32];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        // the packed hi half is fp16(hi); where that differs from hi (|v| below fp16's normal range) the difference moves into lo
                        const __half2 h = __floats2half2_rn(o[2 * c], o[2 * c + 1]);
                        const float2 hf = __half22float2(h);
                        const __half2 l = __floats2half2_rn((lo[2 * c] + (o[2 * c] - hf.x)) * PACK_SCALE, (lo[2 * c + 1] + (o[2 * c + 1] - hf.y)) * PACK_SCALE);
                        pk[c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
                        pk[16 + c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&l));
                    }
                    rowio_store_rows(a.y_pack + orow_w * 32, pk, rowmask, sc);
                }
            } else {
                rowio_store_rows(a.y + orow_w * 32, o, rowmask, sc);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<(F16 ? 512 : 256)>(tmem);
}

}  // namespace

// true when `p` is a 27-tap (dt, dh, dw) lattice convolution 32 -> 32 this kernel handles
bool rowconv3_tc_supported(const RowConvP& p) {
    if (p.ntap != 27 || p.n != 32 || p.xc != 32 || p.kc != 32 || !p.w_kmajor) return false;
    const int pw = p.off[3] - p.off[0], plane = p.off[9] - p.off[0];
    if (pw < 4 || pw > 60 || plane < 3 * pw) return false;
    for (int t = 0; t < 27; ++t)
        if (p.c0[t] != 0 || p.off[t] != p.off[0] + (t / 9) * plane + ((t / 3) % 3) * pw + t % 3) return false;
    return true;
}

int launch_rowconv3_tc(const RowConvP& p, cudaStream_t st) {
    if (!rowconv3_tc_supported(p)) return set_error(PV_ERR_BAD_ARG, "rowconv3_tc: not a 3x3x3 lattice convolution on 32-channel rows");
    Conv3Args a;
    memset(&a, 0, sizeof a);
    const RowGeom& og = p.og;
    const int pw = p.off[3] - p.off[0];
    a.B = p.B; a.in_lead = p.in_lead; a.in_pstride = p.in_pstride; a.og = og; a.pw = pw;
    a.bias = p.bias; a.residual = p.residual; a.relumask = p.relumask; a.y = p.y; a.relu = p.relu; a.round_tf32 = p.round_tf32;
    a.residual2 = p.residual2; a.y_lo = p.y_lo; a.y_pack = p.y_pack;
    if (p.residual2 && !p.residual) return set_error(PV_ERR_BAD_ARG, "rowconv3_tc: residual2 needs residual");
    if ((p.residual2 || p.y_lo) && p.f16_pack != 1) return set_error(PV_ERR_BAD_ARG, "rowconv3_tc: residual2 / y_lo belong to the compensated forward (f16_pack = 1)");
    if (p.y_pack && p.f16_pack != 1 && p.f16_pack != 2) return set_error(PV_ERR_BAD_ARG, "rowconv3_tc: y_pack needs one of the pair-row modes");
    a.slab_rows = ((128 + 2 * pw + 7) / 8) * 8;
    // lane l of a tile accumulates Q[rho = r0 + l]; its A rows for group g are rho + base_g + 1 with base_g = off[3g]
    for (int s = 0; s < 3; ++s) a.slab_lo[s] = p.off[9 * s] + 1;
    for (int t = 0; t < 27; ++t) { a.tap_wr[t] = p.wr0[t]; a.tap_wc[t] = p.wc0[t]; }
    // PR tiling needs: centred taps, same geometry in and out, one zero padding column per line, whole planes
    const bool centred = p.off[13] == 0;
    const bool plane_mode = centred && p.in_lead == og.lead && p.in_pstride == og.pstride && og.pw == pw && og.nw == og.pw - 1 &&
                            og.row0 == 0 && og.t0 == 0 && og.nrows == og.nt * og.plane && p.off[9] - p.off[0] == og.plane &&
                            og.nh * og.pw <= og.plane;
    if (plane_mode) {
        a.plane_mode = 1; a.nt = og.nt; a.plane_rows = og.plane; a.plane_out_rows = og.nh * og.pw;
        a.chunks = cdiv(og.nh * og.pw - 2, C3_TILE);      // first tile yields 127 rows, the others 126; the last row is a padding column
        a.chain_patches = og.pstride == (long long)(og.nt + 1) * og.plane ? 1 : 0;
    } else {
        a.plane_mode = 0; a.nt = 1; a.chunks = cdiv(og.nrows, C3_TILE);
    }
    const size_t smem = 1024 + 27 * 4096 + (size_t)C3_STAGES * a.slab_rows * 128 + 8 * ROWIO_SCRATCH_BYTES;
    if (smem > 222 * 1024) return set_error(PV_ERR_BAD_ARG, "rowconv3_tc: %zu bytes of shared memory needed", smem);
    const long long in_rows = p.in_lead + (long long)p.B * p.in_pstride + ROW_TAIL;
    CUtensorMap tm_x, tm_w;
    PV_TRY(make_tmap_2d(&tm_x, p.x, in_rows, 32, a.slab_rows, 32, 0));
    PV_TRY(make_tmap_2d(&tm_w, p.w, p.w_rows, p.w_cols, 32, 32, 0));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = a.B * a.chunks * a.nt;
    const int grid = ntiles < sms ? ntiles : sms;
    // executed flops: the compensated launch runs the main product (K = 32 per tap) and both corrections (K = 64 per tap)
    PV_TIMED(p.tag ? p.tag : "rowconv3_tc", st, p.flops, 0.0, ((p.f16_pack == 1 || p.f16_pack == 2) ? 3.0 : 1.0) * 2.0 * (double)ntiles * 128.0 * 96.0 * 288.0);
    static size_t attr[16] = {}, attr_h[16] = {}, attr_g[16] = {}, attr_s[16] = {};
    if (p.f16_pack == 3) {
        PV_CUDA(ensure_dyn_smem(rowconv3_tc_kernel<3>, smem, attr_s));
        PV_CUDA(launch_pdl(rowconv3_tc_kernel<3>, grid, C3_THREADS, smem, st, tm_x, tm_w, a));
    } else if (p.f16_pack == 2) {
        PV_CUDA(ensure_dyn_smem(rowconv3_tc_kernel<2>, smem, attr_g));
        PV_CUDA(launch_pdl(rowconv3_tc_kernel<2>, grid, C3_THREADS, smem, st, tm_x, tm_w, a));
    } else if (p.f16_pack) {
        PV_CUDA(ensure_dyn_smem(rowconv3_tc_kernel<1>, smem, attr_h));
        PV_CUDA(launch_pdl(rowconv3_tc_kernel<1>, grid, C3_THREADS, smem, st, tm_x, tm_w, a));
    } else {
        PV_CUDA(ensure_dyn_smem(rowconv3_tc_kernel<0>, smem, attr));
        PV_CUDA(launch_pdl(rowconv3_tc_kernel<0>, grid, C3_THREADS, smem, st, tm_x, tm_w, a));
    }
    PV_LAUNCH_CHECK();
    return 0;
}

}  // namespace pv
