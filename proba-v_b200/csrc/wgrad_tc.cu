// wgrad_tc.cu -- weight gradients on the tcgen05 tensor cores (kind::tf32, fp32 accumulate in TMEM).
//
// Replaces TensorFlow's Conv3DBackpropFilterV2 + BiasAddGrad for the trunk layers (tape.gradient,
// reference models/trainClass.py:131; layers models/modelsTF.py:159-163,179-188):
//     dW[tap][ci][co] = sum over rows r of  x[r + off(tap)][ci] * gz[r][co]          db[co] = sum_r gz[r][co]
// The reduction runs over voxels (rows), so both MMA operands are read "MN-major": a [rows][32] fp32 tile is
// presented as a 32 x rows matrix.  For tf32 the hardware requires the 128B-swizzle-with-32B-atom shared-memory layout
// (descriptor layout type 1; TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) -- probes/umma_probe.cu T3.
// M = 128 of the MMA is filled three ways:
//   conv3  (normConv / convReducer / upscaleConv): four ROW-ADJACENT taps (dw = -1..2, the last one discarded) are one
//          A operand whose four 32-channel atoms are 128 B apart, i.e. overlapping views of the same rows (LBO = 128).
//          The three dh taps ride on the OTHER operand: dW[dt,dh,dw] = sum_r' x[r' + off(dt,0,dw)] * gz[r' - dh*pw], so
//          B is three 32-channel atoms of the gz tile one image line (pw rows) apart (LBO = pw * 128 B) and N = 96.
//          An M128 x N x K8 tf32 MMA costs max(N/2, 32 + N/4) cycles (probes/umma_rate.cu, profiles/r02_umma_rate_probe.log: the 4 KB
//          A fetch paces small N), so one N = 96 MMA (56 cycles) replaces three N = 32 MMAs (120 cycles).  Three dt groups -> three [128 x 96]
//          accumulators = 288 TMEM columns.  Tiles run 2*pw rows past the patch's row range so that every gz row meets
//          every dh (gz is zero outside its valid extent, rows.h invariant).
//   wide x (decConv: x = E, 256 channels): 2 groups of 4 channel atoms (LBO = box stride), N = 32 (gz = gD)
//   wide gz (expConv: gz = gZ, 256 channels): the transpose -- M = gz channels (2 groups), N = 32 = x channels
// Every CTA owns a contiguous range of row tiles and keeps its accumulators in TMEM for its whole lifetime; partial
// sums go to a [cta][group][128][32] scratch and a second small kernel reduces them in a fixed order (deterministic,
// no atomics).  The epilogue warps meanwhile form the bias gradient from the gz tiles already sitting in shared memory.
#include "rows.h"
#include "tc_common.cuh"
#include "wgrad_reduce.cuh"

namespace pv {

using namespace tc;
int make_tmap_2d(CUtensorMap* m, const float* base, long long rows, int cols, int box_rows, int box_cols, int swizzle_32b_atom);

namespace {

constexpr int WG_THREADS = 192;
constexpr int MAX_GROUPS = 9;
constexpr int MAX_BOXES = 9;

struct WgradTcArgs {
    int B, tiles_per_patch, TR;        // TR rows per tile (128 or 64)
    long long a_lead, a_pstride, b_lead, b_pstride;
    int row0;
    int nstage; uint32_t stage_bytes;
    // TMA boxes of one stage: boxes [0, nabox) from the A tensor map, box nabox.. from the B tensor map
    int nabox, nbbox, abox_rows, bbox_rows;
    int box_lo[MAX_BOXES], box_c0[MAX_BOXES]; uint32_t box_off[MAX_BOXES];
    // MMA groups
    int ngroup; uint32_t grp_off[MAX_GROUPS]; uint32_t grp_lbo;
    int nn;                            // N of one MMA: 32, or 96 (conv3: three dh atoms of the gz tile, b_lbo bytes apart)
    uint32_t b_off, b_lbo;
    uint32_t bias_row0;                // first row of the gz box that belongs to this tile (conv3: 2*pw)
    // bias gradient: column sums of these boxes (32 channels each)
    int nbias; uint32_t bias_off[8];
    float* partials;                   // [cta][ngroup][128][32]
    float* db_partials;                // [cta][4 warps][nbias][32]
};

template <int DUMMY>
__global__ void __launch_bounds__(WG_THREADS, 1)
rowwgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const WgradTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * 4 + 1];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto BAR = [&](int i) { return smem_u32(&bars[i]); };
    const int FULL = 0, EMPTY = 4, DONE = 8;
    if (threadIdx.x == 0) {
        for (int i = 0; i < a.nstage; ++i) { mbar_init(BAR(FULL + i), 1); mbar_init(BAR(EMPTY + i), 1 + 4); }
        mbar_init(BAR(DONE), 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int ntiles = a.B * a.tiles_per_patch;
    const int t_lo = (int)((long long)ntiles * blockIdx.x / gridDim.x), t_hi = (int)((long long)ntiles * (blockIdx.x + 1) / gridDim.x);

    if (warp == 0) {
        if (elect_one_sync()) {
            tma_prefetch_desc(&tm_a);
            tma_prefetch_desc(&tm_b);
            pdl_wait();
            pdl_trigger();
            uint32_t it = 0;
            for (int tile = t_lo; tile < t_hi; ++tile, ++it) {
                const int b = tile / a.tiles_per_patch, j = tile % a.tiles_per_patch;
                const long long arow = a.a_lead + (long long)b * a.a_pstride + a.row0 + (long long)j * a.TR;
                const long long brow = a.b_lead + (long long)b * a.b_pstride + a.row0 + (long long)j * a.TR;
                const uint32_t stg = it % a.nstage, ph = (it / a.nstage) & 1;
                mbar_wait(BAR(EMPTY + stg), ph ^ 1);
                mbar_arrive_expect_tx(BAR(FULL + stg), (uint32_t)(a.nabox * a.abox_rows + a.nbbox * a.bbox_rows) * 128u);
                const uint32_t s_addr = base + stg * a.stage_bytes;
                for (int i = 0; i < a.nabox + a.nbbox; ++i)
                    tma_load_2d(s_addr + a.box_off[i], i < a.nabox ? &tm_a : &tm_b, BAR(FULL + stg), a.box_c0[i],
                                (int)((i < a.nabox ? arow : brow) + a.box_lo[i]));
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            const uint64_t HI_A = smem_desc_hi(a.grp_lbo, 512, 1);     // MN-major, 128B swizzle / 32B atom; K atoms (4 rows) 512 B apart
            const uint64_t HI_B = smem_desc_hi(a.b_lbo, 512, 1);
            const uint32_t HI32 = (uint32_t)(HI_A >> 32), LO_A = (uint32_t)HI_A, LO_B = (uint32_t)HI_B;
            const uint32_t IDESC = instr_desc(2, 128, a.nn, 1, 1);
            uint32_t grp_inc[MAX_GROUPS];
#pragma unroll
            for (int g = 0; g < MAX_GROUPS; ++g) grp_inc[g] = a.grp_off[g] >> 4;
            uint32_t it = 0;
            for (int tile = t_lo; tile < t_hi; ++tile, ++it) {
                const uint32_t stg = it % a.nstage, ph = (it / a.nstage) & 1;
                mbar_wait(BAR(FULL + stg), ph);
                tc_fence_after();
                const uint32_t s_lo = (base + stg * a.stage_bytes) >> 4;
                const uint32_t a_lo = s_lo | LO_A, b_lo = (s_lo + (a.b_off >> 4)) | LO_B;
                const uint32_t first = tile > t_lo ? 1u : 0u;
                if (a.nn == 96) {
                    for (int ks = 0; ks < a.TR / 8; ++ks) {
#pragma unroll
                        for (int g = 0; g < 3; ++g)
                            umma_ss_tf32_lohi(tmem + g * 96, a_lo + grp_inc[g] + ks * 64, b_lo + ks * 64, HI32, IDESC, first | (ks > 0 ? 1u : 0u));
                    }
                } else {
                    for (int ks = 0; ks < a.TR / 8; ++ks)
                        for (int g = 0; g < a.ngroup; ++g)
                            umma_ss_tf32_lohi(tmem + g * 32, a_lo + (a.grp_off[g] >> 4) + ks * 64, b_lo + ks * 64, HI32, IDESC, first | (ks > 0 ? 1u : 0u));
                }
                umma_commit(BAR(EMPTY + stg));
            }
            umma_commit(BAR(DONE));
        }
    } else {
        // bias gradient while the tensor core works, then the accumulator drain
        pdl_wait();                                          // the partial buffers may still be read by the previous reduction
        const int q = warp & 3;
        float bsum[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bsum[i] = 0.f;
        uint32_t it = 0;
        const int rows_per_warp = a.TR / 4;
        for (int tile = t_lo; tile < t_hi; ++tile, ++it) {
            const uint32_t stg = it % a.nstage, ph = (it / a.nstage) & 1;
            mbar_wait(BAR(FULL + stg), ph);
            const uint8_t* sp = smem_raw + (base - smem_u32(smem_raw)) + stg * a.stage_bytes;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < a.nbias) {
                    const uint8_t* bx = sp + a.bias_off[i];
                    float s = 0.f;
                    for (int r = a.bias_row0 + q * rows_per_warp; r < (int)a.bias_row0 + (q + 1) * rows_per_warp; ++r)     // 32B-atom swizzle: chunk ^= row & 3
                        s += *reinterpret_cast<const float*>(bx + r * 128 + ((((lane >> 3) ^ (r & 3)) << 5) | ((lane & 7) << 2)));
                    bsum[i] += s;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(EMPTY + stg));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < a.nbias) a.db_partials[(((size_t)blockIdx.x * 4 + q) * a.nbias + i) * 32 + lane] = bsum[i];
        mbar_wait(BAR(DONE), 0);
        tc_fence_after();
        float* out = a.partials + (size_t)blockIdx.x * a.ngroup * 4096;
        for (int g = 0; g < a.ngroup; ++g) {
            if (a.nn == 96 && q == 3) break;                  // conv3: lanes 96..127 are the unused fourth dw slot (never reduced)
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + g * 32, v);
            tmem_ld_wait();
            // conv3: columns g*32.. of the TMEM block are (dt = g/3, atom j = g%3) with dh index 2 - j; stored as group dt*3 + dh
            const int gs = a.nn == 96 ? (g / 3) * 3 + (2 - g % 3) : g;
            float4* o = reinterpret_cast<float4*>(out + ((size_t)gs * 128 + q * 32 + lane) * 32);
            if (t_hi > t_lo) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    o[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]));
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

// immediate (non-deferred) reduction: one launch right after the producer (wgrad_reduce.cuh has the body)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partials, const float* __restrict__ dbp, int ncta, int ngroup, int mode,
                    const __grid_constant__ WgradScatter sc, int nbias) {
    __shared__ float4 sm[256];
    tc::pdl_wait();
    wgrad_reduce_body(blockIdx.x, partials, dbp, ncta, ngroup, mode, sc, nbias, sm);
}

}  // namespace

int launch_rowwgrad_tc(const RowWgradP& p, cudaStream_t st, float* partials, size_t partial_floats, ReduceQueue* rq) {
    int mode;
    if (p.xc == 32 && p.n == 32 && p.ntap == 27) mode = 0;
    else if (p.xc == 256 && p.n == 32 && p.ntap == 8) mode = 1;
    else if (p.xc == 32 && p.n == 256 && p.ntap == 1) mode = 2;
    else return set_error(PV_ERR_BAD_ARG, "rowwgrad_tc: unsupported shape xc=%d n=%d ntap=%d", p.xc, p.n, p.ntap);
    WgradTcArgs a;
    memset(&a, 0, sizeof a);
    a.B = p.B; a.row0 = p.og.row0;
    const float *amat, *bmat;          // A: the wide / tap-shifted operand (M side), B: the 32-channel operand (N side)
    int acols;
    if (mode == 2) { amat = p.gz; acols = 256; bmat = p.x; a.a_lead = p.og.lead; a.a_pstride = p.og.pstride; a.b_lead = p.in_lead; a.b_pstride = p.in_pstride; }
    else { amat = p.x; acols = p.xc; bmat = p.gz; a.a_lead = p.in_lead; a.a_pstride = p.in_pstride; a.b_lead = p.og.lead; a.b_pstride = p.og.pstride; }
    a.nn = 32; a.b_lbo = 128; a.bias_row0 = 0;
    int extra_rows = 0;
    if (mode == 0) {
        a.TR = 128; a.nstage = 3;
        // taps sorted ascending: index 9*dt + 3*dh + dw.  A: one slab per dt holding rows [off(dt,0,0), +TR+3] (four dw views);
        // B: the gz tile plus the 2*pw rows before it, whose three line-shifted views are the dh atoms (atom j <-> dh = 2 - j)
        const int pw = p.off[3] - p.off[0];
        for (int g = 0; g < 9; ++g) {
            if (p.off[3 * g + 1] != p.off[3 * g] + 1 || p.off[3 * g + 2] != p.off[3 * g] + 2 || p.off[3 * g] != p.off[9 * (g / 3)] + (g % 3) * pw)
                return set_error(PV_ERR_BAD_ARG, "rowwgrad_tc: taps must form a (dt, dh, dw) lattice with unit dw stride");
        }
        if (pw < 4 || pw > 60) return set_error(PV_ERR_BAD_ARG, "rowwgrad_tc: unsupported line stride %d", pw);
        a.abox_rows = a.TR + 8;
        a.nabox = 3;
        for (int s = 0; s < 3; ++s) { a.box_lo[s] = p.off[9 * s]; a.box_c0[s] = 0; a.box_off[s] = (uint32_t)s * a.abox_rows * 128; a.grp_off[s] = a.box_off[s]; }
        a.ngroup = 9; a.grp_lbo = 128;             // ngroup counts the [128 x 32] partial blocks (dt, dh); the MMAs run as 3 x N96
        a.nn = 96; a.b_lbo = (uint32_t)pw * 128; a.bias_row0 = 2 * pw;
        a.nbbox = 1; a.bbox_rows = ((a.TR + 2 * pw + 7) / 8) * 8;
        a.box_lo[3] = -2 * pw; a.box_c0[3] = 0; a.box_off[3] = 3u * a.abox_rows * 128;
        a.b_off = a.box_off[3];
        a.nbias = 1; a.bias_off[0] = a.b_off;
        extra_rows = 2 * pw;
    } else {
        a.TR = 64; a.nstage = 3;
        a.nabox = 8; a.abox_rows = a.TR;
        for (int i = 0; i < 8; ++i) { a.box_lo[i] = 0; a.box_c0[i] = 32 * i; a.box_off[i] = (uint32_t)i * a.TR * 128; }
        a.ngroup = 2; a.grp_lbo = (uint32_t)a.TR * 128;
        a.grp_off[0] = 0; a.grp_off[1] = 4u * a.TR * 128;
        a.nbbox = 1; a.bbox_rows = a.TR;
        a.box_lo[8] = 0; a.box_c0[8] = 0; a.box_off[8] = 8u * a.TR * 128;
        a.b_off = a.box_off[8];
        if (mode == 1) { a.nbias = 1; a.bias_off[0] = a.b_off; }
        else { a.nbias = 8; for (int i = 0; i < 8; ++i) a.bias_off[i] = a.box_off[i]; }
    }
    a.tiles_per_patch = cdiv(p.og.nrows + extra_rows, a.TR);
    a.stage_bytes = (uint32_t)(a.nabox * a.abox_rows + a.nbbox * a.bbox_rows) * 128u;
    const size_t smem = 1024 + (size_t)a.nstage * a.stage_bytes;
    if (smem > 226 * 1024) return set_error(PV_ERR_BAD_ARG, "rowwgrad_tc: %zu bytes of shared memory needed", smem);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = a.B * a.tiles_per_patch;
    const int grid = ntiles < sms ? ntiles : sms;
    const size_t need = (size_t)grid * a.ngroup * 4096 + (size_t)grid * 4 * a.nbias * 32;
    float* deferred = rq ? rq->take(need) : nullptr;       // deferred reduction: a private region of the trainer's arena
    if (deferred) { partials = deferred; partial_floats = need; }
    if (!partials || partial_floats < need) return set_error(PV_ERR_BAD_ARG, "rowwgrad_tc: partial buffer too small (%zu < %zu)", partial_floats, need);
    a.partials = partials; a.db_partials = partials + (size_t)grid * a.ngroup * 4096;
    const long long a_rows = a.a_lead + (long long)p.B * a.a_pstride + ROW_TAIL, b_rows = a.b_lead + (long long)p.B * a.b_pstride + ROW_TAIL;
    CUtensorMap tm_a, tm_b;
    PV_TRY(make_tmap_2d(&tm_a, amat, a_rows, acols, a.abox_rows, 32, 1));
    PV_TRY(make_tmap_2d(&tm_b, bmat, b_rows, 32, a.bbox_rows, 32, 1));
    static size_t attr[16] = {};
    PV_CUDA(ensure_dyn_smem(rowwgrad_tc_kernel<0>, smem, attr));
    {
        PV_TIMED(p.tag ? p.tag : "rowwgrad_tc", st, p.flops, 0.0);
        PV_CUDA(launch_pdl(rowwgrad_tc_kernel<0>, grid, WG_THREADS, smem, st, tm_a, tm_b, a));
        PV_LAUNCH_CHECK();
    }
    WgradScatter sc;
    sc.dw = p.dw; sc.dw_cols = p.dw_cols; sc.db = p.db;
    for (int i = 0; i < MAX_TAPS; ++i) { sc.dwr0[i] = p.dwr0[i]; sc.dwc0[i] = p.dwc0[i]; }
    if (deferred) {
        ReduceJob j;
        memset(&j, 0, sizeof j);
        j.kind = 0; j.nblocks = wgrad_reduce_blocks(a.ngroup, a.nbias); j.partials = a.partials; j.dbp = a.db_partials; j.ncta = grid;
        j.ngroup = a.ngroup; j.mode = mode; j.nbias = a.nbias; j.sc = sc;
        rq->push(j);
    } else {
        PV_TIMED("wgrad_reduce", st);
        PV_CUDA(launch_pdl(wgrad_reduce_kernel, wgrad_reduce_blocks(a.ngroup, a.nbias), 256, 0, st, (const float*)a.partials, (const float*)a.db_partials, grid, a.ngroup, mode, sc, a.nbias));
        PV_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace pv
