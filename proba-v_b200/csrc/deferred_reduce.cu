// deferred_reduce.cu -- ONE launch that performs every weight-gradient partial reduction of a backward pass
// (wgrad_reduce.cuh).  The job table travels as a __grid_constant__ kernel parameter (a few KB), each block finds its
// job with a short scan over the block-start prefix and runs that job's fixed-order body.
#include "wgrad_reduce.cuh"

namespace pv {
namespace {

__global__ void __launch_bounds__(256) deferred_reduce_kernel(const __grid_constant__ ReduceJobs J) {
    pdl_grid_wait();
    __shared__ float4 sm[256];
    int j = 0;
    while (j + 1 < J.njobs && (int)blockIdx.x >= J.block_start[j + 1]) ++j;
    const ReduceJob& job = J.jobs[j];
    const int blk = blockIdx.x - J.block_start[j];
    if (job.kind == 0) wgrad_reduce_body(blk, job.partials, job.dbp, job.ncta, job.ngroup, job.mode, job.sc, job.nbias, sm);
    else if (job.kind == 1) resfront_reduce_body(blk, job.partials, job.dbp, job.ncta, job.dwd, job.dwe, job.dbe, job.dbd, sm);
    else if (job.kind == 2) first_conv_reduce_body(blk, job.partials, job.ncta, job.sc.dw, job.sc.db, sm);
    else skip2d_reduce_body(blk, job.partials, job.ncta, job.C, job.o0, job.o1, job.o2, job.o3, job.o4, job.o5, sm);
}

}  // namespace

int launch_deferred_reduce(ReduceQueue& q, cudaStream_t st) {
    if (q.jobs.njobs == 0) return 0;
    const int blocks = q.jobs.block_start[q.jobs.njobs];
    PV_TIMED("wgrad_reduce", st);
    PV_CUDA(launch_pdl_simple(deferred_reduce_kernel, blocks, 256, 0, st, q.jobs));
    PV_LAUNCH_CHECK();
    q.jobs.njobs = 0;
    return 0;
}

}  // namespace pv
