// resblock_x3_tc.cu -- the fused expand (x8) -> ReLU -> decay forward of a WDSR residual block (reference models/modelsTF.py:
// 179-183) with ERROR-COMPENSATED tf32 products (pv_cfg.precision = 4, "tf32x3").
//
// Why (profiles/r02_tf32_numerics_study.md): the L1 shift loss differentiates to sign(residual), so a forward SR error of 3e-4
// (single-pass tf32, 10-bit operands through 40 layers) flips the sign of ~0.1 % of the pixels and puts a ~1e-2 error on every
// gradient tensor, whatever the precision of the backward pass.  The north_star's 1e-3 gradient bar therefore needs an
// fp32-grade FORWARD.  Every operand is carried as hi + lo with hi = tf32(v) (exact in the MMA) and lo = v - hi (|lo| <= 2^-11 |v|,
// truncated to tf32 by the tensor core: 2^-21 overall), and every product is
//        x w  ~=  x_hi w_hi + x_lo w_hi + x_hi w_lo                     (three kind::tf32 MMAs into one fp32 accumulator).
//
// Per 128-row tile and per QUARTER q of the 256 expanded channels (64 channels; three unit buffers + 2 x 2 output accumulators):
//     MMA1 (SS, kind::tf32, N = 64)  E_q = X_hi We_hi,q^T + X_lo We_hi,q^T + X_hi We_lo,q^T              12 MMAs
//     epilogue   v = relu(E_q + be);  ReLU bit mask;  fp16 pair [fp16(v) | fp16(2^12 (v - fp16(v)))] in place, 2 values per TMEM column
//     MMA2 (TS, kind::f16, N = 32)   D += E16 Wd_hi,q^T;   Dc += Elo16 Wd_hi,q^T + E16 (2^12 Wd_lo,q)^T     12 MMAs
//     final      D + 2^-12 Dc + bd -> (hi, lo) -> two row arrays (the lo array as packed fp16 pair rows when a 3x3x3 conv reads it)
// Pipelining: the MMA thread issues MMA1 ahead of MMA2 (see the issue order there); tcgen05 operations execute in issue order,
// which is what makes re-using a unit buffer three units later safe.  Two epilogue groups of four warps take alternate tiles.  mbarrier rule (profiles/r01_next_round_notes.md): a parity wait is only correct if the waiter sees EVERY
// phase of its barrier, so the "E ready" barriers are indexed by (group, quarter): one waiting group, consecutive phases.
#include <cuda_fp16.h>

#include "rowio.cuh"
#include "rows.h"
#include "tc_common.cuh"

namespace pv {

using namespace tc;
int make_tmap_2d(CUtensorMap* m, const float* base, long long rows, int cols, int box_rows, int box_cols, int swizzle_32b_atom);

namespace {

constexpr int RX_THREADS = 320;
constexpr int RX_STAGES = 3;

struct ResX3Args {
    int B, tiles_per_patch;
    RowGeom g;
    const float* bias1;                // be [256]
    const float* bias2;                // bd [32]
    uint32_t* mask;                    // [rows][8] ReLU bit mask (training) or nullptr
    uint32_t* mask_t;                  // [tile][4 row blocks][256 channels]: bit k = row k of the 32-row block (for the weight-gradient kernel)
    float* out_hi;                     // D rows: tf32(D)
    float* out_lo;                     //         D - tf32(D), or (pack_out) the packed fp16 pair rows of (hi, lo): what the
    int pack_out;                      //         compensated 3x3x3 convolution reads (rows.h PACK_SCALE)
    int bias_mma;                      // the expand bias rides in MMA1 (ones tile x bias tile, one more K = 8 step) instead of an FADD per element
    uint32_t bias_lbo, bias_sbo;       // byte strides of the two un-swizzled tiles (K-adjacent core matrices, 8-row groups)
};

__device__ __forceinline__ uint32_t tf32_rn_bits(uint32_t bits) { return (bits + 0x1000u) & 0xffffe000u; }

// The decay GEMM runs in kind::f16 on fp16 PAIRS: the epilogue packs the activated quarter as [fp16(v) x 64 | fp16(2^12 (v - fp16(v))) x 64]
// in place over its accumulator (two 16-bit values per TMEM column, the even channel in the low half: 64 columns), the decay weights
// arrive as the fp16 pair rows of rows.h, and MMA2 is 12 K = 16 MMAs per quarter: E16 Wd_hi into the main accumulator,
// Elo16 Wd_hi + E16 2^12 Wd_lo into the correction accumulator (scaled back by 2^-12 in the final epilogue).
// ILV: issue order of the quarter units (see the MMA thread).
template <int TRAIN, bool ILV, bool BMMA>
__global__ void __launch_bounds__(RX_THREADS, 1)
resfront_fwd_x3_kernel(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
                       const __grid_constant__ CUtensorMap tm_w1h, const __grid_constant__ CUtensorMap tm_w1l,
                       const __grid_constant__ CUtensorMap tm_w2p, const ResX3Args a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[22];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_b1[256];        // (only without bias_mma)
    __shared__ __align__(16) float s_b2[32];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w1h_smem = base, w1l_smem = base + 32768;            // [256 rows x 128 B] each
    const uint32_t w2p_smem = base + 65536;                             // 8 K-chunks x [32 rows x 128 B]: fp16 pair rows [2^12 w_lo x 32 | w_hi x 32]
    const uint32_t x_smem = base + 98304;                               // RX_STAGES x { X_hi, X_lo } x [128 rows x 128 B]
    uint8_t* const io_scratch = smem_raw + (base - smem_u32(smem_raw)) + 98304 + RX_STAGES * 32768;   // 8 epilogue warps x 2 KB (rowio.cuh)
    // bias_mma: A = [128 x 8] tile with ones in columns 0, 1; B = [256 x 8] tile with tf32(be) / be - tf32(be) in columns 0 / 1.  Shared memory
    // is nearly full, so both are UN-SWIZZLED K-major tiles (core matrix = 8 rows x 16 bytes): 4 KB + 8 KB instead of 16 KB + 32 KB.
    const uint32_t ones_smem = base + 98304 + RX_STAGES * 32768 + 8 * ROWIO_SCRATCH_BYTES, biasb_smem = ones_smem + 4096;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto BAR = [&](int i) { return smem_u32(&bars[i]); };
    const int FULL = 0, EMPTY = 3, WBAR = 6, EFULL = 7, EREADY = 15, DFULL = 18, DFREE = 20;
    if (threadIdx.x == 0) {
        for (int i = 0; i < RX_STAGES; ++i) { mbar_init(BAR(FULL + i), 1); mbar_init(BAR(EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(DFULL + i), 1); mbar_init(BAR(DFREE + i), 4); }
        for (int i = 0; i < 8; ++i) mbar_init(BAR(EFULL + i), 1);
        for (int i = 0; i < 3; ++i) mbar_init(BAR(EREADY + i), 4);
        mbar_init(BAR(WBAR), 1);
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < 256; i += RX_THREADS) s_b1[i] = a.bias1[i];
    if (BMMA) {
        uint8_t* const tp = smem_raw + (ones_smem - smem_u32(smem_raw));
        // element (row r, k) of a tile: (r / 8) * sbo + (k / 4) * lbo + (r % 8) * 16 + (k % 4) * 4
        for (int i = threadIdx.x; i < (4096 + 8192) / 16; i += RX_THREADS) reinterpret_cast<uint4*>(tp)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        for (int r = threadIdx.x; r < 128; r += RX_THREADS)
            *reinterpret_cast<float2*>(tp + (r >> 3) * a.bias_sbo + (r & 7) * 16) = make_float2(1.f, 1.f);
        for (int n = threadIdx.x; n < 256; n += RX_THREADS) {
            const float b = a.bias1[n], bh = __uint_as_float(tf32_rn_bits(__float_as_uint(b)));
            *reinterpret_cast<float2*>(tp + 4096 + (n >> 3) * a.bias_sbo + (n & 7) * 16) = make_float2(bh, b - bh);
        }
        fence_proxy_async();                        // generic-proxy writes -> visible to the tensor core's async-proxy reads
    }
    if (threadIdx.x < 32) s_b2[threadIdx.x] = a.bias2[threadIdx.x];
    if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;                // unit buffers (E_hi 64 | E_lo 64) at columns 0 / 128 / 256, D accumulators at 384 / 416 (main), 448 / 480 (corrections)
    const int ntiles = a.B * a.tiles_per_patch;
    const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0) {
        if (elect_one_sync()) {
            tma_prefetch_desc(&tm_xh);
            tma_prefetch_desc(&tm_xl);
            mbar_arrive_expect_tx(BAR(WBAR), 98304);
            tma_load_2d(w1h_smem, &tm_w1h, BAR(WBAR), 0, 0);
            tma_load_2d(w1l_smem, &tm_w1l, BAR(WBAR), 0, 0);
            for (int j = 0; j < 8; ++j) tma_load_2d(w2p_smem + j * 4096, &tm_w2p, BAR(WBAR), 32 * j, 0);
            pdl_wait();
            pdl_trigger();
            for (int tl = 0; tl < my_tiles; ++tl) {
                const int tile = blockIdx.x + tl * gridDim.x;
                const int b = tile / a.tiles_per_patch, j = tile % a.tiles_per_patch;
                const long long row0 = a.g.lead + (long long)b * a.g.pstride + a.g.row0 + j * 128;
                const uint32_t stg = tl % RX_STAGES, ph = (tl / RX_STAGES) & 1;
                mbar_wait(BAR(EMPTY + stg), ph ^ 1);
                mbar_arrive_expect_tx(BAR(FULL + stg), 32768);
                tma_load_2d(x_smem + stg * 32768, &tm_xh, BAR(FULL + stg), 0, (int)row0);
                tma_load_2d(x_smem + stg * 32768 + 16384, &tm_xl, BAR(FULL + stg), 0, (int)row0);
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            constexpr uint64_t HI = smem_desc_hi(16, 1024, 2);
            constexpr uint32_t HI32 = (uint32_t)(HI >> 32), LO32 = (uint32_t)HI;
            constexpr uint32_t IDESC1 = instr_desc(2, 128, 64, 0, 0);
            constexpr uint32_t IDESC2 = instr_desc(0, 128, 32, 0, 0);       // fp16 operands (A from TMEM), N = 32
            mbar_wait(BAR(WBAR), 0);
            tc_fence_after();
            // Issue order of the quarter units.  Sequential (ILV = false): tile after tile, MMA1 one unit ahead of MMA2; the epilogue group of
            // the NEXT tile then gets its first unit only while this tile's last quarter is in the epilogue (ncu: 20 % of the stall samples sit
            // on that wait).  Interleaved (ILV = true): the units of a PAIR of tiles alternate (A0 B0 A1 B1 ... A3 B3, A = even tile = epilogue
            // group 0, B = odd tile = group 1) with MMA1 two units ahead: each group finds its next unit computed while its own current unit
            // and the other group's occupy the other two buffers.  Both tiles of a pair are then in flight at once, hence three X stages.
            const int U = 4 * my_tiles;
            constexpr int LA = ILV ? 2 : 1;
            auto unit_of = [&](int n, int& tl, int& q) {
                if (ILV) {
                    const int p = n >> 3, r = n & 7;
                    if (2 * p + 1 < my_tiles) { tl = 2 * p + (r & 1); q = r >> 1; }
                    else { tl = 2 * p; q = r; }                 // the last, unpaired tile: its four units in a row
                } else { tl = n >> 2; q = n & 3; }
            };
            for (int n = 0; n < U + LA; ++n) {
                if (n < U) {                        // MMA1 of unit n
                    int tl, q;
                    unit_of(n, tl, q);
                    const uint32_t stg = tl % RX_STAGES, ph = (tl / RX_STAGES) & 1, eb = n % 3;
                    if (q == 0) { mbar_wait(BAR(FULL + stg), ph); tc_fence_after(); }
                    const uint32_t xh = ((x_smem + stg * 32768) >> 4) | LO32, xl = ((x_smem + stg * 32768 + 16384) >> 4) | LO32;
                    const uint32_t wh = ((w1h_smem + q * 8192) >> 4) | LO32, wl = ((w1l_smem + q * 8192) >> 4) | LO32;
                    const uint32_t d = tmem + eb * 128;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_ss_tf32_lohi(d, xh + 2 * ks, wh + 2 * ks, HI32, IDESC1, ks > 0);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_ss_tf32_lohi(d, xl + 2 * ks, wh + 2 * ks, HI32, IDESC1, 1u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_ss_tf32_lohi(d, xh + 2 * ks, wl + 2 * ks, HI32, IDESC1, 1u);
                    if (BMMA) {                     // + ones . [be_hi | be_lo]^T of this quarter's 64 channels (un-swizzled descriptors)
                        const uint64_t hn = smem_desc_hi(a.bias_lbo, a.bias_sbo, 0);
                        const uint32_t hn32 = (uint32_t)(hn >> 32), ln32 = (uint32_t)hn;
                        umma_ss_tf32_lohi(d, (ones_smem >> 4) | ln32, ((biasb_smem + (uint32_t)q * 8u * a.bias_sbo) >> 4) | ln32, hn32, IDESC1, 1u);
                    }
                    if (q == 3) umma_commit(BAR(EMPTY + stg));
                    umma_commit(BAR(EFULL + (tl & 1) * 4 + q));
                }
                if (n >= LA) {                      // MMA2 of unit n - LA
                    const int v = n - LA;
                    int tl, q;
                    unit_of(v, tl, q);
                    const uint32_t eb = v % 3, db = tl & 1;
                    mbar_wait(BAR(EREADY + eb), (v / 3) & 1);
                    tc_fence_after();
                    if (q == 0) { mbar_wait(BAR(DFREE + db), ((tl >> 1) & 1) ^ 1); tc_fence_after(); }
                    // Two accumulators per tile: the main chain E16 Wd_hi and the correction chain.  tcgen05 truncates the fp32 accumulator at
                    // every accumulation step (one ulp of the ACCUMULATOR, whatever the addend's size), so chaining the tiny corrections behind
                    // the main sum would cost a truncation of the full-size value per correction step (profiles/r02_tf32_numerics_study.md);
                    // the correction accumulator is 2^-11 of the size (before its 2^12 scale) and its truncations are harmless.
                    const uint32_t d = tmem + 384 + 32 * db, dc = tmem + 448 + 32 * db, eh = tmem + eb * 128, el = eh + 32;
#pragma unroll
                    for (int jl = 0; jl < 2; ++jl) {                // the quarter's two 32-channel chunks: 16 TMEM columns each
                        const uint64_t bp = smem_desc(HI, w2p_smem + (uint32_t)(2 * q + jl) * 4096);     // row = [2^12 w_lo x 32 | w_hi x 32]
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            const uint32_t ac = (uint32_t)(jl * 16 + t * 8);
                            umma_ts<false>(d, eh + ac, bp + 2 * (2 + t), IDESC2, (q > 0 || jl > 0 || t > 0) ? 1u : 0u);
                            umma_ts<false>(dc, el + ac, bp + 2 * (2 + t), IDESC2, (q > 0 || jl > 0 || t > 0) ? 1u : 0u);
                            umma_ts<false>(dc, eh + ac, bp + 2 * t, IDESC2, 1u);
                        }
                    }
                    if (q == 3) umma_commit(BAR(DFULL + db));
                }
            }
        }
    } else {
        const int q4 = warp & 3;
        const int grp = (warp - 2) >> 2;
        const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
        pdl_wait();
        for (int tl = grp; tl < my_tiles; tl += 2) {
            const int tile = blockIdx.x + tl * gridDim.x;
            const int b = tile / a.tiles_per_patch, j = tile % a.tiles_per_patch;
            const int r = a.g.row0 + j * 128 + q4 * 32 + lane;
            const bool in_patch = r < a.g.row0 + a.g.nrows && r < a.g.pstride;
            const bool valid = in_patch && row_valid(a.g, r);
            const long long orow = a.g.lead + (long long)b * a.g.pstride + r;
            const uint32_t rowmask = __ballot_sync(0xffffffffu, in_patch);
            const long long orow_w = orow - lane;
            uint8_t* const sc = io_scratch + (warp - 2) * ROWIO_SCRATCH_BYTES;
            const uint32_t tph = (tl >> 1) & 1;               // this group's (tl >> 1)-th tile: every barrier below sees consecutive phases
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {                     // rolled: the unrolled body thrashes the instruction cache
                // position of this unit in the MMA thread's issue order (unit_of there)
                const int u = !ILV ? 4 * tl + q : ((tl | 1) < my_tiles ? 8 * (tl >> 1) + 2 * q + (tl & 1) : 8 * (tl >> 1) + q);
                const uint32_t eb = u % 3, hb = lane_base + eb * 128;
                mbar_wait(BAR(EFULL + grp * 4 + q), tph);
                tc_fence_after();
                uint32_t va[32], vb[32];
                uint32_t ph[32], pl[32];                      // the quarter as fp16 pairs, two channels per word
                uint32_t wq[2] = {0u, 0u};                    // TRAIN: this quarter's two words of the row's ReLU mask
                tmem_ld32(hb, va);
                tmem_ld32(hb + 32, vb);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c == 0) tmem_ld_wait();               // both loads are complete after the first wait
                    uint32_t (&cur)[32] = c ? vb : va;
                    uint32_t sg[4] = {0u, 0u, 0u, 0u};
                    uint32_t colbits = 0u;                    // TRAIN: lane e keeps the row bit-vector of channel e of this chunk
                    const float4* be4 = reinterpret_cast<const float4*>(s_b1 + q * 64 + c * 32);
#pragma unroll
                    for (int e4 = 0; e4 < 8; ++e4) {
                        const float4 bq = be4[e4];
                        const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float v = fmaxf(BMMA ? __uint_as_float(cur[e4 * 4 + e]) : __uint_as_float(cur[e4 * 4 + e]) + bb[e], 0.f);
                            const uint32_t x = __float_as_uint(v);
                            // (bits(v) - 1) has its sign bit set exactly when v == 0, i.e. when the pre-activation is <= 0
                            // (tf.nn.relu's gradient convention)
                            if (TRAIN) sg[e4 >> 1] = __funnelshift_l(x - 1u, sg[e4 >> 1], 1);
                            cur[e4 * 4 + e] = x;
                        }
                    }
                    {
#pragma unroll
                        for (int k = 0; k < 16; ++k) {        // channels 2k, 2k + 1 of this chunk -> one word each of the hi and lo halves
                            const float v0 = __uint_as_float(cur[2 * k]), v1 = __uint_as_float(cur[2 * k + 1]);
                            const __half2 h2 = __floats2half2_rn(v0, v1);
                            const float2 hf = __half22float2(h2);
                            const __half2 l2 = __floats2half2_rn((v0 - hf.x) * PACK_SCALE, (v1 - hf.y) * PACK_SCALE);
                            ph[c * 16 + k] = *reinterpret_cast<const uint32_t*>(&h2);
                            pl[c * 16 + k] = *reinterpret_cast<const uint32_t*>(&l2);
                        }
                    }
                    if (TRAIN) {                              // word q * 2 + c of the row's mask (static register indexing under the rolled loop)
                        const uint32_t word = ~((sg[0] << 24) | ((sg[1] & 0xffu) << 16) | ((sg[2] & 0xffu) << 8) | (sg[3] & 0xffu));
                        // 32 x 32 bit-matrix transpose across the warp (lane = row, bit 31 - e = channel e  ->  lane = channel, bit k = row k):
                        // reverse the bits so that bit e = channel e, then five butterfly stages
                        colbits = __brev(word);
#pragma unroll
                        for (int j = 16; j >= 1; j >>= 1) {
                            const uint32_t msk = j == 16 ? 0x0000ffffu : (j == 8 ? 0x00ff00ffu : (j == 4 ? 0x0f0f0f0fu : (j == 2 ? 0x33333333u : 0x55555555u)));
                            const uint32_t other = __shfl_xor_sync(0xffffffffu, colbits, j);
                            // lanes with bit j clear keep their low halves and take the partner's low halves shifted up; the others mirror it
                            colbits = (lane & j) ? ((colbits & ~msk) | ((other >> j) & msk)) : ((colbits & msk) | ((other << j) & ~msk));
                        }
                        wq[c] = word;
                        // transposed copy: one coalesced 128-byte store per (32 rows x 32 channels)
                        if (a.mask_t) a.mask_t[((size_t)tile * 4 + q4) * 256 + q * 64 + c * 32 + lane] = colbits;
                    }
                }
                // words 2q, 2q + 1 of the row-major mask (one 8-byte store per quarter: collecting the tile's eight words in registers
                // under the rolled q loop costs a 16-compare select chain per quarter)
                if (TRAIN && a.mask && in_patch) reinterpret_cast<uint2*>(a.mask + orow * 8)[q] = make_uint2(wq[0], wq[1]);
                tmem_st32(hb, ph);                            // in place over the accumulator (both halves of it are in registers)
                tmem_st32(hb + 32, pl);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(EREADY + eb));
            }
            const uint32_t db = tl & 1;                       // == grp
            mbar_wait(BAR(DFULL + db), tph);
            tc_fence_after();
            uint32_t v[32], vc[32];
            tmem_ld32(lane_base + 384 + 32 * db, v);
            tmem_ld32(lane_base + 448 + 32 * db, vc);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(DFREE + db));
#pragma unroll
            for (int c = 0; c < 32; ++c)              // main + 2^-12 corrections, rounded to nearest
                v[c] = __float_as_uint(fmaf(__uint_as_float(vc[c]), 1.0f / PACK_SCALE, __uint_as_float(v[c])));
            float hi[32], lo[32];
#pragma unroll
            for (int g4 = 0; g4 < 8; ++g4) {
                const float4 bq = reinterpret_cast<const float4*>(s_b2)[g4];
                const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float d = valid ? __uint_as_float(v[g4 * 4 + e]) + bb[e] : 0.f;
                    const float h = __uint_as_float(tf32_rn_bits(__float_as_uint(d)));
                    hi[g4 * 4 + e] = h;
                    lo[g4 * 4 + e] = d - h;
                }
            }
            rowio_store_rows(a.out_hi + orow_w * 32, hi, rowmask, sc);
            if (a.pack_out) {
                float pk[32];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const __half2 h2 = __floats2half2_rn(hi[2 * c], hi[2 * c + 1]);
                    const float2 hf = __half22float2(h2);                 // lo' = lo + (hi - fp16(hi)), rows.h
                    const __half2 l2 = __floats2half2_rn((lo[2 * c] + (hi[2 * c] - hf.x)) * PACK_SCALE, (lo[2 * c + 1] + (hi[2 * c + 1] - hf.y)) * PACK_SCALE);
                    pk[c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h2));
                    pk[16 + c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&l2));
                }
                rowio_store_rows(a.out_lo + orow_w * 32, pk, rowmask, sc);
            } else {
                rowio_store_rows(a.out_lo + orow_w * 32, lo, rowmask, sc);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace

// (D_hi, D_lo) = split(decConv(relu(expConv(X_hi + X_lo)))) on PR rows with compensated products.
//   weT_exp_* [256][32] (K contiguous; hi = tf32(w), lo = w - hi), weT_dec_pack [32][256] as fp16 pair rows per 32-channel chunk
//   (rows.h: [2^12 w_lo x 32 | w_hi x 32]), biases padded to 256 / 32.
//   relu_bits (nullable): [rows][8] uint32 in resblock_tc.cu's format (consumed by the backward-data kernel).
//   relu_bits_t (nullable): the same bits transposed, [tile][4][256] uint32 with bit k = row 32 * block + k of the tile, for the
//   weight-gradient kernel (whose threads own channels): it must use the FORWARD's mask -- its own single-pass recomputation of E
//   disagrees with the compensated forward on ~1e-4 of the elements, which alone is a ~1e-2 error on dWe (sqrt law).
int launch_resfront_fwd_x3_tc(const float* x_hi, const float* x_lo, const float* weT_exp_hi, const float* weT_exp_lo,
                              const float* weT_dec_pack, const float* bias_e, const float* bias_d,
                              float* d_hi, float* d_lo, uint32_t* relu_bits, uint32_t* relu_bits_t, const RowGeom& g, int B, double flops,
                              cudaStream_t st, int pack_out) {
    ResX3Args a;
    memset(&a, 0, sizeof a);
    a.B = B; a.g = g; a.bias1 = bias_e; a.bias2 = bias_d; a.mask = relu_bits; a.mask_t = relu_bits_t; a.out_hi = d_hi; a.out_lo = d_lo; a.pack_out = pack_out;
    a.tiles_per_patch = cdiv(g.nrows, 128);
    const long long rows = g.lead + (long long)B * g.pstride + ROW_TAIL;
    CUtensorMap tm_xh, tm_xl, tm_w1h, tm_w1l, tm_w2p;
    PV_TRY(make_tmap_2d(&tm_xh, x_hi, rows, 32, 128, 32, 0));
    PV_TRY(make_tmap_2d(&tm_xl, x_lo, rows, 32, 128, 32, 0));
    PV_TRY(make_tmap_2d(&tm_w1h, weT_exp_hi, 256, 32, 256, 32, 0));
    PV_TRY(make_tmap_2d(&tm_w1l, weT_exp_lo, 256, 32, 256, 32, 0));
    PV_TRY(make_tmap_2d(&tm_w2p, weT_dec_pack, 32, 256, 32, 32, 0));
    const size_t smem = 1024 + 98304 + RX_STAGES * 32768 + 8 * ROWIO_SCRATCH_BYTES + 4096 + 8192;
    {   // PV_X3_BIAS_MMA=0 keeps the per-element bias add (A/B); PV_X3_BIAS_DESC=lbo,sbo overrides the tile strides (bring-up)
        static const char* off = getenv("PV_X3_BIAS_MMA");
        static const char* ds = getenv("PV_X3_BIAS_DESC");
        a.bias_mma = !(off && off[0] == '0');
        a.bias_lbo = 128; a.bias_sbo = 256;
        if (ds) { unsigned l = 0, s2 = 0; if (sscanf(ds, "%u,%u", &l, &s2) == 2) { a.bias_lbo = l; a.bias_sbo = s2; } }
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = a.B * a.tiles_per_patch;
    const int grid = ntiles < sms ? ntiles : sms;
    // executed: three products per GEMM on the padded 32 x 256 shapes
    PV_TIMED(relu_bits ? "resfront_fwd_x3" : "resfront_fwd_x3_infer", st, flops, 0.0, 3.0 * 2.0 * 2.0 * (double)ntiles * 128.0 * 32.0 * 256.0);
    static const bool sequential = getenv("PV_X3_SEQUENTIAL") != nullptr;     // A/B: the previous issue order
    static size_t attr[8][16] = {};
    auto go = [&](auto kern, size_t (&at)[16]) -> int {
        PV_CUDA(ensure_dyn_smem(kern, smem, at));
        PV_CUDA(launch_pdl(kern, grid, RX_THREADS, smem, st, tm_xh, tm_xl, tm_w1h, tm_w1l, tm_w2p, a));
        return 0;
    };
    const int variant = (relu_bits ? 4 : 0) | (sequential ? 0 : 2) | (a.bias_mma ? 1 : 0);
    switch (variant) {
        case 0: PV_TRY(go(resfront_fwd_x3_kernel<0, false, false>, attr[0])); break;
        case 1: PV_TRY(go(resfront_fwd_x3_kernel<0, false, true>, attr[1])); break;
        case 2: PV_TRY(go(resfront_fwd_x3_kernel<0, true, false>, attr[2])); break;
        case 3: PV_TRY(go(resfront_fwd_x3_kernel<0, true, true>, attr[3])); break;
        case 4: PV_TRY(go(resfront_fwd_x3_kernel<1, false, false>, attr[4])); break;
        case 5: PV_TRY(go(resfront_fwd_x3_kernel<1, false, true>, attr[5])); break;
        case 6: PV_TRY(go(resfront_fwd_x3_kernel<1, true, false>, attr[6])); break;
        default: PV_TRY(go(resfront_fwd_x3_kernel<1, true, true>, attr[7])); break;
    }
    PV_LAUNCH_CHECK();
    return 0;
}

}  // namespace pv
