// dec_staged.cu -- the four launches of the staged order-1 decode (dec_staged.cuh): head, table, chain, post.
// Every launch covers the whole job list of a batch and takes the jobs routed to it (route 2 with a DecPrep
// record); a stream the head stage hands back gets route 1 and is decoded by dec_kernel<true> (kernels.cu),
// which the host launches behind the head stage.
#include "dec_staged.cuh"
#include "kernels.h"

namespace b200 {

constexpr int DH_WARPS = 2;             // head: two streams per CTA, 8 KiB (one DecO0Smem) each
constexpr int DC_WARPS = 2;             // chain: two streams per CTA

__global__ void __launch_bounds__(DH_WARPS * 32)
dec_head_kernel(DecJob *jobs, uint32_t njobs, Pool pool) {
    // (16-byte alignment is all that is promised: the hardware places dynamic shared memory behind its own 1 KiB, and
    // a larger claim here lets the compiler fold dec_o0's run-time alignment test to "true")
    extern __shared__ __align__(16) uint8_t smem_head[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t j = blockIdx.x * DH_WARPS + wid;
    if (j >= njobs || jobs[j].route != 2 || !jobs[j].prep) return;
    dec_head(jobs[j], *(DecPrep *)jobs[j].prep, smem_head + (size_t)wid * DEC_SMEM_O0, pool, lane);
}

__global__ void __launch_bounds__(DT_THREADS, 4)
dec_table_kernel(DecJob *jobs, uint32_t njobs) {
    __shared__ DecTabSmem S;
    const uint32_t j = blockIdx.x;
    if (j >= njobs || !jobs[j].prep) return;
    DecPrep &P = *(DecPrep *)jobs[j].prep;
    if (P.state != 1) return;
    dec_table_stage(P, S);
}

// 28 warps per SM (<= 72 registers): the 3815 streams of a 1 GB block are one wave
__global__ void __launch_bounds__(DC_WARPS * 32, 14)
dec_chain_kernel(DecJob *jobs, uint32_t njobs) {
    __shared__ DecChainSmem S[DC_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t j = blockIdx.x * DC_WARPS + wid;
    if (j >= njobs || !jobs[j].prep) return;
    DecPrep &P = *(DecPrep *)jobs[j].prep;
    if (P.state != 1) return;
    const int e = P.N == 32 ? dec_chain_big<32>(P, S[wid], lane) : dec_chain_big<4>(P, S[wid], lane);
    if (e && lane == 0) { P.state = 2; P.status = ST_FAIL; }
}

__global__ void __launch_bounds__(DP_THREADS, 4)
dec_post_kernel(DecJob *jobs, uint32_t njobs) {
    __shared__ DecPostSmem S;
    const uint32_t j = blockIdx.x;
    if (j >= njobs || !jobs[j].prep) return;
    DecPrep &P = *(DecPrep *)jobs[j].prep;
    if (P.state == 0) return;                       // the general kernel decoded (or failed) this stream
    if (P.state == 2) {
        if (threadIdx.x == 0) { jobs[j].status = P.status; jobs[j].out_size = 0; }
        return;
    }
    dec_post_stage(jobs[j], P, S);
}

cudaError_t dec_staged_stats(unsigned long long *out16, bool reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out16, g_dec_stats, sizeof(g_dec_stats));
    if (e == cudaSuccess && reset) {
        unsigned long long z[16] = {0};
        e = cudaMemcpyToSymbol(g_dec_stats, z, sizeof(z));
    }
    return e;
}

size_t dec_prep_bytes() { return (sizeof(DecPrep) + 255) & ~(size_t)255; }

static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

cudaError_t launch_dec_head(DecJob *d_jobs, uint32_t n, Pool pool, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const size_t sm = (size_t)DEC_SMEM_O0 * DH_WARPS;
    dec_head_kernel<<<cdiv(n, DH_WARPS), DH_WARPS * 32, sm, st>>>(d_jobs, n, pool);
    return cudaGetLastError();
}

cudaError_t launch_dec_staged_rest(DecJob *d_jobs, uint32_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    dec_table_kernel<<<n, DT_THREADS, 0, st>>>(d_jobs, n);
    dec_chain_kernel<<<cdiv(n, DC_WARPS), DC_WARPS * 32, 0, st>>>(d_jobs, n);
    dec_post_kernel<<<n, DP_THREADS, 0, st>>>(d_jobs, n);
    return cudaGetLastError();
}

}  // namespace b200
