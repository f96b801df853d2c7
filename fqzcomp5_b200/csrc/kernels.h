// kernels.h -- host-visible launchers for the kernels in kernels.cu.
#pragma once
#include "common.cuh"
#include "rans_decode.cuh"

namespace b200 {

// one warp per stream, several streams per CTA
constexpr int ENC_WARPS = 4;
constexpr int ENC_WARPS_O1 = 2;
constexpr int DEC_WARPS = 2;
constexpr int DEC_WARPS_O1 = 2;
// shared memory per warp (bytes)
constexpr uint32_t ENC_SMEM_O0 = 6144;     // EncO0Smem: ring + 256 encoder symbols + histogram
constexpr uint32_t ENC_SMEM_O1 = 9856;     // EncO1Smem header + 4-byte encoder symbols for <= 41 symbols (22 warps per SM)
// order-1 streams that are PACKed / RLEd first (their alphabet is usually all 256 byte values):
// room for the 8 x 256 counters of the partitioned pair count behind the EncO1Smem header
constexpr uint32_t ENC_SMEM_O1_WIDE = 11520;
// order-1 streams prepared by prep_kernel: ring + rank map (1280 bytes) + 4-byte encoder symbols for <= 40 symbols;
// the order-0 coder of the run-length meta-data and of the table (6 KiB) uses the same bytes first.  28 warps per SM.
constexpr uint32_t ENC_SMEM_O1_PREP = 7680;
constexpr uint32_t DEC_SMEM_O0 = 8192;     // DecO0Smem, 8 KiB aligned
constexpr uint32_t DEC_SMEM_O1 = 8192;     // DecO1Smem header + 16-bit cumulative rows + 64-bucket index for <= 41 symbols (26 warps per SM: measured 1.35x over 15 KiB / 256 buckets)

cudaError_t launch_hist(EncJob *d_jobs, uint32_t n, cudaStream_t st);
// PACK / RLE streams with a prep area: transforms, counts and the order-1 model by one CTA per stream
cudaError_t launch_prep(EncJob *d_jobs, uint32_t n, cudaStream_t st);
size_t prep_area_bytes(uint32_t in_size);
cudaError_t launch_enc(EncJob *d_jobs, uint32_t n, uint32_t route, Pool pool, cudaStream_t st, bool inslot = false);
cudaError_t launch_dec(DecJob *d_jobs, uint32_t n, bool o1, Pool pool, cudaStream_t st);
// staged decode of order-1 streams behind PACK / RLE (dec_staged.cu): the head stage first -- it hands streams it
// does not take back to route 1, so launch_dec(o1) for this batch must be ordered behind it -- then table, chain, post
size_t dec_prep_bytes();
cudaError_t dec_staged_stats(unsigned long long *out16, bool reset);   // diagnostics: see g_dec_stats
cudaError_t launch_dec_head(DecJob *d_jobs, uint32_t n, Pool pool, cudaStream_t st);
cudaError_t launch_dec_staged_rest(DecJob *d_jobs, uint32_t n, cudaStream_t st);
// method trial: items first[k]..first[k+1]-1 are the candidates of input k -> sizes of all, first smallest kept
cudaError_t launch_trial_select(EncJob *d_jobs, uint32_t njobs, uint32_t ninputs, const uint32_t *d_first,
                                uint32_t *d_csize, uint32_t *d_jobidx, int32_t *d_best, cudaStream_t st);
cudaError_t launch_pack(const EncJob *d_jobs, uint32_t n, uint64_t *d_off, uint32_t *d_size,
                        uint64_t *d_total, uint8_t *d_out, uint64_t out_cap, uint32_t align, cudaStream_t st);
// in-slot output: STRIPE parents are assembled inside their own slots, then every item's place is reported
cudaError_t launch_inslot_results(EncJob *d_jobs, uint32_t njobs, const uint32_t *d_parents, uint32_t nparents,
                                  const uint8_t *base, uint64_t *d_off, uint32_t *d_size, cudaStream_t st);

}  // namespace b200
