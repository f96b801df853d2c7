// stripe.h -- RANS_ORDER_STRIPE support: host-side planning (only header bytes are
// read on the host) and launchers for the transpose / select kernels.
// Reference: rANS_static4x16pr.c:1266-1393 (encode), :1615-1694 (decode),
// utils.h:79-138 (unstripe).
#pragma once
#include <vector>
#include "common.cuh"

namespace b200 {

struct StripeSub { uint64_t off; uint32_t len; int order; };

struct StripePlan {
    int item;               // index in the caller's batch
    uint32_t in_size;
    uint32_t N;             // number of stripes actually used
    uint32_t nmeth;         // candidate methods per stripe
    uint32_t nsub;          // N * nmeth
    uint32_t first_job;     // parent record; sub-stream (i, j) is first_job + 1 + i*nmeth + j
    size_t o_transposed;    // offset of the transposed copy in the work arena
    std::vector<StripeSub> sub;
};

void stripe_plan_encode(StripePlan &sp, int item, uint32_t in_size, int order, uint32_t cap);

struct DecItem {
    bool fail = false, stripe = false;
    uint32_t first_job = 0, njobs = 0, ulen = 0, N = 0;
    size_t o_tmp = 0;
    std::vector<uint32_t> sub_off, sub_clen, sub_ulen, sub_idx;
};
bool stripe_plan_decode(DecItem &it, const unsigned char *in, uint32_t in_size, uint32_t out_size);

cudaError_t launch_stripe_split(const uint8_t *d_in, uint8_t *d_tr, uint32_t n, uint32_t N, cudaStream_t st);
cudaError_t launch_stripe_join(const uint8_t *d_parts, uint8_t *d_out, uint32_t n, uint32_t N, cudaStream_t st);
// every STRIPE parent of a batch at once: d_parents lists their job indices
cudaError_t launch_stripe_split_batch(const EncJob *d_jobs, const uint32_t *d_parents, uint32_t nparents,
                                      uint32_t max_in_size, cudaStream_t st);
cudaError_t launch_stripe_select(EncJob *d_jobs, const uint32_t *d_parents, uint32_t nparents, cudaStream_t st);
cudaError_t launch_dec_results(const DecJob *d_jobs, uint32_t n, uint32_t *d_osz, int *d_status, cudaStream_t st);

}  // namespace b200
