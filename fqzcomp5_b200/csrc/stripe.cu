// stripe.cu -- RANS_ORDER_STRIPE: N-way byte transpose, per-stripe choice of the
// smallest candidate method, and the host-side planning around them.
// Reference: rANS_static4x16pr.c:1266-1393 (encode), :1615-1694 (decode),
// utils.h:79-138 (unstripe).
#include "stripe.h"

namespace b200 {

// part j holds bytes j, j+N, j+2N, ...; parts are stored back to back
__device__ __forceinline__ uint32_t part_start(uint32_t j, uint32_t n, uint32_t N) {
    uint32_t q = n / N, r = n % N;
    return j * q + (j < r ? j : r);
}

__global__ void stripe_split_kernel(const uint8_t *in, uint8_t *tr, uint32_t n, uint32_t N) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        tr[part_start(i % N, n, N) + i / N] = in[i];
}

// All STRIPE parents of a batch in one launch: blockIdx.y walks the parents, the transposed copy of
// parent p is the input of its first sub-stream (the parts are stored back to back from there).
__global__ void stripe_split_batch_kernel(const EncJob *jobs, const uint32_t *parents, uint32_t nparents) {
    for (uint32_t p = blockIdx.y; p < nparents; p += gridDim.y) {
        const EncJob &P = jobs[parents[p]];
        const uint32_t n = P.in_size, N = P.stripe_n;
        const uint8_t *in = P.in;
        uint8_t *tr = const_cast<uint8_t *>(jobs[parents[p] + 1].in);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
            tr[part_start(i % N, n, N) + i / N] = in[i];
    }
}

__global__ void stripe_join_kernel(const uint8_t *parts, uint8_t *out, uint32_t n, uint32_t N) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = parts[part_start(i % N, n, N) + i / N];
}

// One thread walks the stripes in order, exactly as the reference's loop does:
// a candidate counts only if the space left in the caller's buffer would have
// let it succeed (need_cap), and the first smallest one wins.
__global__ void stripe_select_kernel(EncJob *jobs, const uint32_t *parents, uint32_t nparents) {
    const uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= nparents) return;
    const uint32_t parent = parents[pi];
    EncJob &P = jobs[parent];
    const uint32_t N = P.stripe_n, nmeth = P.stripe_nmeth;
    uint8_t *out = P.slot;
    uint32_t *list = (uint32_t *)(P.slot + P.slot_cap);
    const uint32_t cap = P.cap;
    int order = P.order;
    // the caller's order after the size fix-ups; the flag byte drops NOSZ (:1314)
    if ((order & ORDER_SIMD_AUTO) && P.in_size >= 50000 && !(order & X_STRIPE)) order |= X_32;
    if (P.in_size <= 1000) order &= ~X_32;
    out[0] = (uint8_t)(order & ~X_NOSZ);
    uint32_t meta = 1 + var_put_u32(out + 1, P.in_size);
    uint32_t status = ST_OK;
    if (cap == 0 || meta >= cap) status = ST_FAIL;
    out[meta++] = (uint8_t)N;
    uint64_t used = 7 + 5 * N, total = 0;
    for (uint32_t i = 0; i < N && status == ST_OK; i++) {
        uint32_t best = 0xffffffffu, best_sz = 0x7fffffffu;
        for (uint32_t j = 0; j < nmeth; j++) {
            uint32_t q = parent + 1 + i * nmeth + j;
            const EncJob &S = jobs[q];
            if (used > cap) continue;
            uint64_t room = cap - used;
            uint32_t sz = S.head_len + S.tail_len;
            if (S.status != ST_OK || S.need_cap > room || !sz) continue;
            if (best_sz > sz) { best_sz = sz; best = q; }
        }
        if (best == 0xffffffffu) { status = ST_FAIL; break; }
        list[i] = best;
        used += best_sz;
        total += best_sz;
        meta += var_put_u32(out + meta, best_sz);
    }
    P.status = status;
    P.head_len = meta;
    P.tail_len = (uint32_t)total;
    P.tail = nullptr;
    P.stripe_n = N;
}

__global__ void dec_results_kernel(const DecJob *jobs, uint32_t n, uint32_t *osz, int *status) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { osz[k] = jobs[k].out_size; status[k] = jobs[k].status; }
}

cudaError_t launch_stripe_split(const uint8_t *d_in, uint8_t *d_tr, uint32_t n, uint32_t N, cudaStream_t st) {
    if (!n) return cudaSuccess;
    uint32_t blocks = (n + 1023) / 1024;
    if (blocks > 148 * 16) blocks = 148 * 16;
    stripe_split_kernel<<<blocks, 256, 0, st>>>(d_in, d_tr, n, N);
    return cudaGetLastError();
}
cudaError_t launch_stripe_join(const uint8_t *d_parts, uint8_t *d_out, uint32_t n, uint32_t N, cudaStream_t st) {
    if (!n) return cudaSuccess;
    uint32_t blocks = (n + 1023) / 1024;
    if (blocks > 148 * 16) blocks = 148 * 16;
    stripe_join_kernel<<<blocks, 256, 0, st>>>(d_parts, d_out, n, N);
    return cudaGetLastError();
}
cudaError_t launch_stripe_split_batch(const EncJob *d_jobs, const uint32_t *d_parents, uint32_t nparents,
                                      uint32_t max_in_size, cudaStream_t st) {
    if (!nparents) return cudaSuccess;
    uint32_t bx = (max_in_size + 4095) / 4096;         // 16 bytes per thread and trip at most
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    uint32_t by = nparents < 32768 ? nparents : 32768;
    stripe_split_batch_kernel<<<dim3(bx, by), 256, 0, st>>>(d_jobs, d_parents, nparents);
    return cudaGetLastError();
}
cudaError_t launch_stripe_select(EncJob *d_jobs, const uint32_t *d_parents, uint32_t nparents, cudaStream_t st) {
    if (!nparents) return cudaSuccess;
    stripe_select_kernel<<<(nparents + 31) / 32, 32, 0, st>>>(d_jobs, d_parents, nparents);
    return cudaGetLastError();
}
cudaError_t launch_dec_results(const DecJob *d_jobs, uint32_t n, uint32_t *d_osz, int *d_status, cudaStream_t st) {
    if (!n) return cudaSuccess;
    dec_results_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_jobs, n, d_osz, d_status);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- host planning
void stripe_plan_encode(StripePlan &sp, int item, uint32_t in_size, int order, uint32_t cap) {
    (void)cap;
    static const int methods[4] = {1, 64, 128, 0};
    if ((order & ORDER_SIMD_AUTO) && in_size >= 50000 && !(order & X_STRIPE)) order |= X_32;
    if (in_size <= 1000) order &= ~X_32;
    uint32_t N = (order >> 8) & 0xff;
    if (N == 0) N = 4;
    if (N > in_size) N = in_size;
    sp.item = item;
    sp.in_size = in_size;
    sp.N = N;
    sp.nmeth = 0;
    int meth[4];
    for (int j = 0; j < 4; j++) {
        if ((order & methods[j]) != methods[j]) continue;
        if ((order & ORDER_STRIPE_NO0) && !(methods[j] & 1)) continue;
        meth[sp.nmeth++] = methods[j];
    }
    sp.nsub = N * sp.nmeth;
    sp.sub.resize(sp.nsub);
    uint32_t q = in_size / N, r = in_size % N;
    for (uint32_t i = 0; i < N; i++) {
        uint32_t start = i * q + (i < r ? i : r), len = q + (r > i);
        for (uint32_t j = 0; j < sp.nmeth; j++)
            sp.sub[i * sp.nmeth + j] = StripeSub{start, len, meth[j] | X_NOSZ | (order & X_32)};
    }
}

static int host_var_get(const unsigned char *p, const unsigned char *end, uint32_t *v) {
    const unsigned char *s = p;
    uint32_t x = 0;
    int cnt = 0;
    unsigned char c = 0x80;
    while ((c & 0x80) && p < end && cnt < 6) { c = *p++; x = (x << 7) | (c & 0x7f); cnt++; }
    *v = x;
    return (int)(p - s);
}

bool stripe_plan_decode(DecItem &it, const unsigned char *in, uint32_t in_size, uint32_t out_size) {
    const unsigned char *end = in + in_size;
    uint32_t ulen, meta = 1;
    meta += host_var_get(in + meta, end, &ulen);
    if (meta >= in_size) return false;
    uint32_t N = in[meta++];
    if (N < 1) return false;
    if (ulen != out_size) return false;              // :1640
    it.stripe = true;
    it.N = N;
    it.ulen = ulen;
    it.sub_off.resize(N); it.sub_clen.resize(N); it.sub_ulen.resize(N); it.sub_idx.resize(N);
    std::vector<uint32_t> clen(N);
    uint64_t ctot = 0;
    for (uint32_t i = 0; i < N; i++) {
        it.sub_ulen[i] = ulen / N + ((ulen % N) > i);
        it.sub_idx[i] = i ? it.sub_idx[i - 1] + it.sub_ulen[i - 1] : 0;
        meta += host_var_get(in + meta, end, &clen[i]);
        ctot += clen[i];
        if (meta > in_size || clen[i] > in_size || clen[i] < 1) return false;
    }
    if (meta + ctot > in_size) return false;
    for (uint32_t i = 0; i < N; i++) {
        it.sub_off[i] = meta;
        // the reference hands over everything that is left (:1680); a valid sub-stream
        // never reads past its own clen bytes, so that is all we stage
        it.sub_clen[i] = clen[i];
        meta += clen[i];
    }
    return true;
}

}  // namespace b200
