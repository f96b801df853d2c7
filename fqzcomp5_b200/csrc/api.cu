// api.cu -- the C ABI of libb200rans.so (include/b200rans.h): per-thread CUDA
// contexts, batch planning, host<->device staging, and the reference's seven
// entry points expressed as one-stream batches.  No codec arithmetic happens on
// the host: every byte of every stream is produced or consumed by the kernels.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <vector>
#include <thread>
#include <mutex>
#include <atomic>

#include "runtime.h"
#include "fastq.h"
#include "crc32.h"

using namespace b200;
using namespace b200rt;

#define API extern "C" __attribute__((visibility("default")))

namespace b200rt {

thread_local Ctx *tls_ctx = nullptr;
thread_local int tls_device = -1;
struct CtxOwner { ~CtxOwner() { delete tls_ctx; tls_ctx = nullptr; } };
thread_local CtxOwner tls_owner;

void set_thread_device(int device) { tls_device = device; }
std::atomic<uint64_t> g_worker_launches{0};

Ctx *get_ctx(int *err) {
    (void)&tls_owner;
    int want = tls_device;
    if (want < 0) {
        const char *e = getenv("B200RANS_DEVICE");
        want = e ? atoi(e) : 0;
    }
    if (tls_ctx && tls_ctx->dev != want) { delete tls_ctx; tls_ctx = nullptr; }
    if (!tls_ctx) {
        Ctx *c = new Ctx();
        int r = c->init(want);
        if (r) { delete c; if (err) *err = r; return nullptr; }
        tls_ctx = c;
    }
    cudaSetDevice(tls_ctx->dev);
    return tls_ctx;
}

// ---------------------------------------------------------------- planning
// Effective order after the size-dependent fix-ups the reference applies first
// (rANS_static4x16pr.c:1256-1265).  Only used to route streams to kernels and to
// size scratch; the kernels redo the fix-ups themselves.
inline int effective_order(uint32_t in_size, int order) {
    if ((order & ORDER_SIMD_AUTO) && in_size >= 50000 && !(order & X_STRIPE)) order |= X_32;
    if (in_size <= 20) order &= ~X_STRIPE;
    if (in_size <= 1000) order &= ~X_32;
    return order;
}

// Build and run the encode of a batch whose inputs are already on the device.
// On return (asynchronously on st): d_out holds the streams, d_out_off / d_out_size (and d_total,
// packed mode only) describe them.
//   packed mode (inslot == false): streams are copied back to back into d_out, each starting on a
//       multiple of pack_align; slots live in the lane's work arena.
//   in-slot mode: every caller item's slot is carved out of d_out (rans_compress_bound_4x16 bytes,
//       256-byte aligned -- the reference's "one bound-sized buffer per call") and the stream stays
//       where the encoder built it; no scan, no gather.  out_cap >= enc_slots_bound().
// Bytes a stream's private slot needs.  rans_compress_bound_4x16 reserves 257*257*3 bytes for an order-1
// table whatever the input size; a table never exceeds the alphabet list (<= 513 bytes) plus two bytes
// per distinct (context, symbol) pair and two per zero run (at most one more run than pairs per row),
// i.e. 4 * (n + 32) + 2 * 256 + 513 bytes, so small streams (STRIPE sub-streams above all) get small slots.
inline uint32_t slot_cap_for(uint32_t isz, int ord) {
    uint32_t b = compress_bound(isz, ord);
    if (ord & 0xff) {       // the bound takes the order-1 branch for ANY non-zero flag byte (rANS_static4x16pr.c:97-100)
        const uint64_t table_worst = 257 * 257 * 3 + 4;
        const uint64_t table_real = (ord & 1) ? std::min<uint64_t>(table_worst, 4ull * isz + 2048) : 0;
        b -= (uint32_t)(table_worst - table_real);
    }
    return (uint32_t)al(b + 16, 16);
}

size_t enc_slots_bound(int n, const uint32_t *in_size, const int *order) {
    size_t t = 0;
    for (int k = 0; k < n; k++) {
        size_t b = al(compress_bound(in_size[k], order[k]) + 16, 16);
        if (effective_order(in_size[k], order[k]) & X_STRIPE) b += STRIPE_LIST_BYTES + 16;
        t = al(t, 256) + b;
    }
    return al(t, 256) + 256;
}

int enc_core(Ctx &C, Lane &Ln, cudaStream_t st, int n, const uint8_t *d_in, const uint64_t *in_off,
             const uint32_t *in_size, const int *order, const uint32_t *caps,
             uint8_t *d_out, size_t out_cap, uint64_t *d_out_off, uint32_t *d_out_size,
             uint64_t *d_total, const Trial *trial, uint32_t pack_align, bool inslot) {
    if (n <= 0) return 0;
    if (pack_align == 0 || (pack_align & (pack_align - 1))) return B200RANS_EINVAL;
    // ---- pass 1: count jobs and size scratch
    std::vector<StripePlan> stripes;
    size_t njobs = 0;
    Layout L;                   // the lane's work arena
    Layout LO;                  // in-slot mode: the caller's d_out, from its first 256-byte aligned address
    const size_t out_adj = (size_t)((256 - ((uintptr_t)d_out & 255)) & 255);
    std::vector<uint32_t> first(n);
    for (int k = 0; k < n; k++) {
        int eo = effective_order(in_size[k], order[k]);
        first[k] = (uint32_t)njobs;
        if (eo & X_STRIPE) {
            StripePlan sp;
            stripe_plan_encode(sp, k, in_size[k], order[k], caps ? caps[k] : compress_bound(in_size[k], order[k]));
            sp.first_job = (uint32_t)njobs;
            njobs += 1 + sp.nsub;             // parent (assembly record) + sub-streams
            stripes.push_back(sp);
        } else njobs++;
    }
    std::vector<EncJob> jobs(njobs);
    std::vector<uint8_t> in_out(njobs, 0);    // slot offset is relative to d_out (in-slot mode, caller items)
    size_t o_jobs = L.take(njobs * sizeof(EncJob));
    size_t o_ctr = L.take(256);
    size_t o_par = L.take(stripes.size() * 4 + 4);
    uint32_t n_o0 = 0, n_o1 = 0, n_o1w = 0, n_o1p = 0, n_model = 0, n_prep = 0;
    size_t pool_bytes = 0;
    const bool prep = use_prep();
    // ---- pass 2: place slots / work buffers
    auto place = [&](size_t j, const uint8_t *in, uint32_t isz, int ord, uint32_t cap, uint32_t item, bool parent) {
        EncJob &J = jobs[j];
        memset(&J, 0, sizeof(J));
        J.in = in; J.in_size = isz; J.order = ord; J.cap = cap; J.item = item;
        // caller items keep the reference's bound in in-slot mode (it is what the header promises)
        uint32_t slot_cap = (inslot && item != 0xffffffffu) ? (uint32_t)al(compress_bound(isz, ord) + 16, 16)
                                                            : slot_cap_for(isz, ord);
        J.slot_cap = slot_cap;
        const size_t slot_bytes = (size_t)slot_cap + (parent ? STRIPE_LIST_BYTES + 16 : 0);
        if (inslot && item != 0xffffffffu) { J.slot = (uint8_t *)LO.take(slot_bytes, 256); in_out[j] = 1; }
        else J.slot = (uint8_t *)L.take(slot_bytes, 256);
        if (parent) return;
        if (ord & (X_PACK | X_RLE)) {
            J.work = (uint8_t *)L.take((size_t)isz * 4 + isz / 4 + 8192, 256);
            // transforms, counts and the order-1 model by one CTA per stream, in front of the coder warp
            if (prep && isz && !(ord & X_CAT)) { J.prep = (uint8_t *)(L.take(prep_area_bytes(isz), 256) + 1); n_prep++; }
        } else if (isz >= hist_min_bytes((ord & 1) != 0)) {
            // big plain streams: counts come from hist_kernel (order-1: up to (isz+1)^2 or 256^2 pairs)
            size_t pairs = ((ord & 1) && isz >= 8) ? std::min<size_t>(65536, ((size_t)isz + 1) * (isz + 1)) : 0;
            J.model = (uint32_t *)(L.take((MODEL_HDR_WORDS + pairs) * 4, 256) + 1);   // +1: null stays null
            n_model++;
        }
        if ((ord & 1) && isz >= 8) {
            // the order-1 kernel (larger shared memory per warp); transformed streams get the launch
            // with room for the partitioned pair count
            J.route = J.prep ? ROUTE_O1_PREP : (ord & (X_PACK | X_RLE)) ? ROUTE_O1_WIDE : ROUTE_O1;
            if (J.route == ROUTE_O1_WIDE) n_o1w++;
            if (J.route == ROUTE_O1_PREP) n_o1p++;
            n_o1++;
            // symbol table (16 B/pair), pair counts (4 B/pair), coded table scratch
            // + 16-bit pair keys of the partitioned pair count (large alphabets without a model);
            // prepared streams bring their own areas
            const size_t m = std::min<size_t>(256, (size_t)isz + 1);       // alphabet of a short stream
            if (!J.prep)
                pool_bytes += m * m * 12 + std::min<size_t>(300 * 1024, 5 * (size_t)isz + 8192) + 2048 +
                              (J.model ? 0 : 2 * (size_t)isz + 2048);
        } else n_o0++;
    };
    size_t si = 0;
    uint32_t max_stripe_in = 0;
    for (int k = 0; k < n; k++) {
        uint32_t j = first[k];
        uint32_t cap = caps ? caps[k] : compress_bound(in_size[k], order[k]);
        if (si < stripes.size() && stripes[si].item == k) {
            StripePlan &sp = stripes[si++];
            sp.o_transposed = L.take(in_size[k], 256);
            place(j, d_in + in_off[k], in_size[k], order[k], cap, (uint32_t)k, true);
            jobs[j].route = ROUTE_NONE;      // assembled by stripe_select, not coded
            jobs[j].stripe_n = sp.N;
            jobs[j].stripe_nmeth = sp.nmeth;
            max_stripe_in = std::max(max_stripe_in, in_size[k]);
            for (uint32_t s2 = 0; s2 < sp.nsub; s2++) {
                const StripeSub &ss = sp.sub[s2];
                // sub-streams get generous private slots; the capacity rule of the
                // reference is re-applied at selection time (need_cap)
                place(j + 1 + s2, (const uint8_t *)(sp.o_transposed + ss.off), ss.len, ss.order,
                      0x7ffffff0u, 0xffffffffu, false);
            }
        } else {
            place(j, d_in + in_off[k], in_size[k], order[k], cap, (uint32_t)k, false);
        }
    }
    if (inslot && out_adj + LO.off > out_cap) return B200RANS_ESPACE;
    // the reference's worst case is one table per stream; real tables are tiny.
    pool_bytes = std::min<size_t>(pool_bytes, (size_t)8 << 30);
    pool_bytes = std::max<size_t>(pool_bytes, (size_t)8 << 20);
    size_t o_pool = L.take(pool_bytes, 256);
    size_t o_soff = L.take(njobs * 8), o_ssz = L.take(njobs * 4), o_tot = L.take(8);
    int r = Ln.work.ensure(L.off + 256);
    if (r) return r;
    uint8_t *W = Ln.work.p;
    // relocate offsets into pointers
    for (size_t j = 0; j < njobs; j++) {
        EncJob &J = jobs[j];
        J.slot = (in_out[j] ? d_out + out_adj : W) + (size_t)J.slot;
        if (J.work) J.work = W + (size_t)J.work;
        if (J.model) J.model = (uint32_t *)(W + ((size_t)J.model - 1));
        if (J.prep) J.prep = W + ((size_t)J.prep - 1);
    }
    for (auto &sp : stripes)
        for (uint32_t s2 = 0; s2 < sp.nsub; s2++) {
            EncJob &J = jobs[sp.first_job + 1 + s2];
            J.in = W + (size_t)J.in;
        }
    Stage *S;
    const size_t par_bytes = stripes.size() * 4;
    size_t stage_bytes = njobs * sizeof(EncJob) + par_bytes + 256;
    if ((r = Ln.get_stage(stage_bytes, &S))) return r;
    memcpy(S->h.p, jobs.data(), njobs * sizeof(EncJob));
    uint32_t *h_par = (uint32_t *)(S->h.p + njobs * sizeof(EncJob));
    for (size_t i = 0; i < stripes.size(); i++) h_par[i] = stripes[i].first_job;
    EncJob *d_jobs = (EncJob *)(W + o_jobs);
    uint32_t *d_par = (uint32_t *)(W + o_par);
    CK(cudaMemcpyAsync(d_jobs, S->h.p, njobs * sizeof(EncJob), cudaMemcpyHostToDevice, st));
    if (par_bytes) CK(cudaMemcpyAsync(d_par, h_par, par_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(W + o_ctr, 0, 256, st));
    Pool pool{W + o_pool, pool_bytes, (unsigned long long *)(W + o_ctr)};
    const uint32_t npar = (uint32_t)stripes.size();

    // ---- STRIPE: transpose parents into their sub-stream inputs (one launch for all of them)
    if (npar) { CK(launch_stripe_split_batch(d_jobs, d_par, npar, max_stripe_in, st)); C.launches++; }
    // ---- PACK / RLE streams: transforms, counts, order-1 models (one CTA per stream)
    if (n_prep) { CK(launch_prep(d_jobs, (uint32_t)njobs, st)); C.launches++; }
    // ---- histograms of the plain streams at full occupancy
    if (n_model) { CK(launch_hist(d_jobs, (uint32_t)njobs, st)); C.launches++; }
    // ---- encode: order-0 streams on the lean kernel, the rest on the order-1 kernel
    if (C.prof) CK(cudaEventRecord(C.pe[0], st));
    {
        // streams of different routes are independent: their launches go to side streams so that the
        // few long PACK / RLE order-1 streams of a mixed batch (a method trial) overlap the rest
        const uint32_t routes[4] = {ROUTE_O1_PREP, ROUTE_O1_WIDE, ROUTE_O1, ROUTE_O0};        // longest first
        const bool have[4] = {n_o1p != 0, n_o1w != 0, n_o1 > n_o1w + n_o1p, n_o0 != 0};
        const int nr = (int)have[0] + have[1] + have[2] + have[3];
        if (nr > 1) CK(cudaEventRecord(Ln.fork, st));
        int used = 0;
        for (int i = 0; i < 4; i++) {
            if (!have[i]) continue;
            cudaStream_t s2 = used == 0 ? st : Ln.aux[used - 1];
            if (used) CK(cudaStreamWaitEvent(s2, Ln.fork, 0));
            CK(launch_enc(d_jobs, (uint32_t)njobs, routes[i], pool, s2, inslot));
            C.launches++;
            if (used) { CK(cudaEventRecord(Ln.join[used - 1], s2)); CK(cudaStreamWaitEvent(st, Ln.join[used - 1], 0)); }
            used++;
        }
    }
    if (C.prof) { CK(cudaEventRecord(C.pe[1], st)); C.pe_valid[0] = true; }
    CK(cudaEventRecord(S->ev, st)); S->busy = true;
    // ---- STRIPE: choose the smallest method per sub-stream and write the parents' headers
    if (npar) { CK(launch_stripe_select(d_jobs, d_par, npar, st)); C.launches++; }
    // ---- method trial: keep the first smallest candidate of each input
    if (trial) {
        CK(launch_trial_select(d_jobs, (uint32_t)njobs, trial->inputs, trial->d_first, trial->d_csize,
                               trial->d_jobidx, trial->d_best, st));
        C.launches += 2;
    }
    uint64_t *d_off = d_out_off ? d_out_off : (uint64_t *)(W + o_soff);
    uint32_t *d_sz = d_out_size ? d_out_size : (uint32_t *)(W + o_ssz);
    if (inslot) {
        // ---- streams stay in their slots; STRIPE parents collect their sub-streams into theirs
        CK(launch_inslot_results(d_jobs, (uint32_t)njobs, d_par, npar, d_out, d_off, d_sz, st));
        C.launches += npar ? 2 : 1;
        return 0;
    }
    // ---- pack the finished streams of the caller's items (sub-streams carry no item)
    CK(launch_pack(d_jobs, (uint32_t)njobs, d_off, d_sz, d_total ? d_total : (uint64_t *)(W + o_tot), d_out, out_cap,
                   pack_align, st));
    C.launches += 2;
    return 0;
}

// Decode core: inputs and outputs on the device.  d_status/d_osz are device arrays.
int dec_core(Ctx &C, Lane &Ln, cudaStream_t st, int n, const uint8_t *d_in, const uint64_t *in_off,
             const uint32_t *in_size, const uint8_t *flags /* first byte of each stream, or null */,
             uint8_t *d_out, const uint64_t *out_off, const uint32_t *out_cap,
             uint32_t *d_osz, int *d_status) {
    if (n <= 0) return 0;
    Layout L;
    size_t o_jobs = L.take((size_t)n * sizeof(DecJob));
    size_t o_ctr = L.take(256);
    std::vector<DecJob> jobs(n);
    uint32_t n_o0 = 0, n_o1 = 0, n_st = 0;
    size_t pool_bytes = 0;
    const bool staged = use_dec_staged();
    for (int k = 0; k < n; k++) {
        DecJob &J = jobs[k];
        memset(&J, 0, sizeof(J));
        J.in = d_in + in_off[k]; J.in_size = in_size[k];
        J.out = d_out + out_off[k]; J.out_cap = out_cap[k];
        // flag byte unknown: transform scratch is provided and the stream goes to the general
        // (order-1) kernel, which also decodes order-0 and raw streams
        const bool known = flags != nullptr;
        const int f = known ? flags[k] : (X_PACK | X_RLE | 1);
        const bool o1 = !known || ((f & 1) && !(f & X_CAT));
        // order-1 behind PACK / RLE with a known flag byte: staged decode (the head stage hands small alphabets back
        // to the general kernel); its post stage keeps a bitmap of the run-length bytes behind the usual scratch
        const bool st2 = staged && known && o1 && (f & (X_PACK | X_RLE)) && !(f & X_STRIPE);
        if (f & (X_PACK | X_RLE))
            J.tmp = (uint8_t *)L.take((size_t)out_cap[k] * 2 + 4096 + (st2 ? (size_t)out_cap[k] / 4 + 4096 : 0), 256);
        if (st2) {
            J.route = 2; n_st++;
            J.prep = (uint8_t *)L.take(dec_prep_bytes(), 256);
            pool_bytes += 257 * 257 * 3 + 257 * 256 * 4 + 256 * 2048 + 16 * 1024;
        } else if (o1) {
            J.route = 1; n_o1++;
            pool_bytes += 257 * 257 * 3 + 257 * 256 * 4 + 256 * 2048 + 16 * 1024;   // table text + DecO1Big
        } else n_o0++;
    }
    pool_bytes = std::min<size_t>(pool_bytes, (size_t)16 << 30);
    pool_bytes = std::max<size_t>(pool_bytes, (size_t)8 << 20);
    size_t o_pool = L.take(pool_bytes, 256);
    int r = Ln.work.ensure(L.off + 256);
    if (r) return r;
    uint8_t *W = Ln.work.p;
    for (auto &J : jobs) {
        if (J.tmp) J.tmp = W + (size_t)J.tmp;
        if (J.route == 2) J.prep = W + (size_t)J.prep;
    }
    Stage *S;
    if ((r = Ln.get_stage((size_t)n * sizeof(DecJob), &S))) return r;
    memcpy(S->h.p, jobs.data(), (size_t)n * sizeof(DecJob));
    DecJob *d_jobs = (DecJob *)(W + o_jobs);
    CK(cudaMemcpyAsync(d_jobs, S->h.p, (size_t)n * sizeof(DecJob), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(W + o_ctr, 0, 256, st));
    Pool pool{W + o_pool, pool_bytes, (unsigned long long *)(W + o_ctr)};
    if (C.prof) CK(cudaEventRecord(C.pe[2], st));
    // independent streams: the launches run side by side -- order-0 on a side stream from the start, the
    // general order-1 kernel on another one behind the head stage (which may route streams to it), the
    // staged launches on `st`
    int njoin = 0;
    if (n_o0 && (n_o1 || n_st)) {
        CK(cudaEventRecord(Ln.fork, st));
        CK(cudaStreamWaitEvent(Ln.aux[0], Ln.fork, 0));
        CK(launch_dec(d_jobs, (uint32_t)n, false, pool, Ln.aux[0]));
        CK(cudaEventRecord(Ln.join[njoin], Ln.aux[0]));
        njoin++;
        C.launches++;
    } else if (n_o0) { CK(launch_dec(d_jobs, (uint32_t)n, false, pool, st)); C.launches++; }
    if (n_st) {
        CK(launch_dec_head(d_jobs, (uint32_t)n, pool, st));
        CK(cudaEventRecord(Ln.fork2, st));
        CK(cudaStreamWaitEvent(Ln.aux[1], Ln.fork2, 0));
        CK(launch_dec(d_jobs, (uint32_t)n, true, pool, Ln.aux[1]));
        CK(cudaEventRecord(Ln.join[njoin], Ln.aux[1]));
        njoin++;
        CK(launch_dec_staged_rest(d_jobs, (uint32_t)n, st));
        C.launches += 5;
    } else if (n_o1) { CK(launch_dec(d_jobs, (uint32_t)n, true, pool, st)); C.launches++; }
    for (int q = 0; q < njoin; q++) CK(cudaStreamWaitEvent(st, Ln.join[q], 0));
    if (C.prof) { CK(cudaEventRecord(C.pe[3], st)); C.pe_valid[1] = true; }
    CK(cudaEventRecord(S->ev, st)); S->busy = true;
    CK(launch_dec_results(d_jobs, (uint32_t)n, d_osz, d_status, st));
    C.launches++;
    return 0;
}

// copy a list of host ranges to/from device offsets, merging neighbours that are
// laid out identically on both sides into one cudaMemcpyAsync.  Host-to-device
// copies may also bridge gaps of up to 64 bytes (alignment padding inside one
// host arena; fewer than a page, so the bytes between two valid ranges are
// readable); device-to-host copies only merge exact neighbours.
struct Span { const uint8_t *h; size_t d; size_t len; };
int copy_spans(std::vector<Span> &sp, uint8_t *dbase, bool to_device, cudaStream_t st) {
    const size_t max_gap = to_device ? 64 : 0;
    size_t i = 0;
    while (i < sp.size()) {
        size_t j = i + 1, len = sp[i].len;
        while (j < sp.size() && sp[j].h >= sp[i].h + len && (size_t)(sp[j].h - (sp[i].h + len)) <= max_gap &&
               sp[j].d - sp[i].d == (size_t)(sp[j].h - sp[i].h)) {
            len = (size_t)(sp[j].h - sp[i].h) + sp[j].len;
            j++;
        }
        if (len) {
            if (to_device) CK(cudaMemcpyAsync(dbase + sp[i].d, sp[i].h, len, cudaMemcpyHostToDevice, st));
            else CK(cudaMemcpyAsync((void *)sp[i].h, dbase + sp[i].d, len, cudaMemcpyDeviceToHost, st));
        }
        i = j;
    }
    return 0;
}

// device placement of host buffers [k0,k1): mirror the host layout where buffers
// follow each other closely so that copies merge; otherwise start a new
// 256-byte aligned run with the same alignment mod 16 as on the host
void place_spans(int k0, int k1, const unsigned char *const *ptr, const uint32_t *len, uint64_t *off,
                 size_t *total) {
    size_t o = 0;
    for (int k = k0; k < k1; k++) {
        if (k > k0 && ptr[k] >= ptr[k - 1] + len[k - 1] && (size_t)(ptr[k] - (ptr[k - 1] + len[k - 1])) <= 64)
            o = off[k - 1] + (size_t)(ptr[k] - ptr[k - 1]);
        else o = al(o, 256) + ((uintptr_t)ptr[k] & 15);
        off[k] = o;
        o += len[k];
    }
    *total = o + 256;
}

// ------------------------------------------------------------ host-buffer encode
struct EncChunk {
    int k0 = 0, k1 = 0;
    Lane *L = nullptr;
    size_t o_in = 0, o_out = 0, o_off = 0, o_sz = 0, o_tot = 0, bound_total = 0;
    size_t o_cs = 0, o_ji = 0, o_best = 0, o_first = 0;   // method trial: candidate sizes, job index scratch, winners, group offsets
    size_t base = 0;            // where this chunk's streams start in the caller's arena
};

// With mfirst != null this is the method trial: input k is encoded under
// methods[mfirst[k] .. mfirst[k+1]), the first smallest stream is returned, best[k] is
// its index in that list (-1: all failed) and csize[c] the size of candidate c (0: that
// call failed).
int compress_batch_impl(int n, const unsigned char *const *in, const unsigned int *in_size, const int *order,
                        const uint32_t *caps, unsigned char *out, size_t out_cap, size_t *out_off,
                        unsigned int *out_size, const uint32_t *mfirst, const int *methods,
                        int *best, unsigned int *csize) {
    const bool M = mfirst != nullptr;
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    if (n <= 0) return 0;
    // ---- split into pipeline chunks of ~chunk_bytes() of input
    bool slow = false;
    {
        const int *o = M ? methods : order;
        const size_t no = M ? mfirst[n] : (size_t)n;
        for (size_t i = 0; i < no && !slow; i++) slow = (o[i] & (X_PACK | X_RLE)) != 0;
    }
    // large 4-lane streams: a 256 KiB stream is 65 536 dependent steps per lane (several ms) whatever the size of the
    // launch, so a chunk should hold as many of them as the GPU has warp slots
    bool lanes4 = false;
    for (int k = 0; k < n && !lanes4; k++) {
        if (in_size[k] < 32768) continue;
        const uint32_t c0 = M ? mfirst[k] : (uint32_t)k, c1 = M ? mfirst[k + 1] : (uint32_t)k + 1;
        const int *o = M ? methods : order;
        for (uint32_t c2 = c0; c2 < c1 && !lanes4; c2++)
            lanes4 = !(o[c2] & X_32) && !((o[c2] & ORDER_SIMD_AUTO) && in_size[k] >= 50000) && !(o[c2] & (X_STRIPE | X_CAT));
    }
    const int min_streams = lanes4 ? chunk_min_streams_4lane() : slow ? chunk_min_streams_slow(false) : chunk_min_streams();
    std::vector<EncChunk> ch;
    for (int k = 0; k < n;) {
        EncChunk c;
        c.k0 = k;
        size_t acc = 0;
        auto streams = [&](int k1) { return M ? (int)(mfirst[k1] - mfirst[c.k0]) : k1 - c.k0; };
        while (k < n && (k == c.k0 || ((acc + in_size[k] <= chunk_bytes() || streams(k) < min_streams) &&
                                       streams(k) < CHUNK_MAX_STREAMS)))
            acc += in_size[k++];
        c.k1 = k;
        ch.push_back(c);
    }
    std::vector<uint64_t> ioff(n);
    size_t out_base = 0;
    int rc = 0;

    auto submit = [&](EncChunk &c) -> int {
        Lane &Ln = *c.L;
        int m = c.k1 - c.k0;
        size_t in_total;
        place_spans(c.k0, c.k1, in, in_size, ioff.data(), &in_total);
        c.bound_total = 0;
        for (int k = c.k0; k < c.k1; k++) {
            size_t b = 0;
            if (M) for (uint32_t c2 = mfirst[k]; c2 < mfirst[k + 1]; c2++)
                b = std::max<size_t>(b, compress_bound(in_size[k], methods[c2]));
            else b = compress_bound(in_size[k], order[k]);
            c.bound_total += al(b, 16) + 16;
        }
        Layout L;
        c.o_in = L.take(in_total); c.o_out = L.take(c.bound_total);
        c.o_off = L.take((size_t)m * 8); c.o_sz = L.take((size_t)m * 4); c.o_tot = L.take(8);
        const size_t mm = M ? mfirst[c.k1] - mfirst[c.k0] : 0;      // candidate calls of this chunk
        c.o_cs = L.take(mm * 4); c.o_ji = L.take(mm * 4); c.o_best = L.take((size_t)m * 4);
        c.o_first = L.take((size_t)(m + 1) * 4);
        int r = Ln.io.ensure(L.off + 256);
        if (r) return r;
        uint8_t *D = Ln.io.p;
        std::vector<Span> sp(m);
        for (int k = c.k0; k < c.k1; k++) sp[k - c.k0] = Span{in[k], c.o_in + ioff[k], in_size[k]};
        if ((r = copy_spans(sp, D, true, Ln.st))) return r;
        if (M) {
            // the input is staged once; its candidate calls share it
            std::vector<uint64_t> xoff(mm);
            std::vector<uint32_t> xsz(mm);
            const uint32_t f0 = mfirst[c.k0];
            Stage *SF;
            if ((r = Ln.get_stage((size_t)(m + 1) * 4, &SF))) return r;
            uint32_t *hf = (uint32_t *)SF->h.p;
            for (int k = c.k0; k <= c.k1; k++) hf[k - c.k0] = mfirst[k] - f0;
            for (int k = c.k0; k < c.k1; k++)
                for (uint32_t c2 = mfirst[k]; c2 < mfirst[k + 1]; c2++) { xoff[c2 - f0] = ioff[k]; xsz[c2 - f0] = in_size[k]; }
            CK(cudaMemcpyAsync(D + c.o_first, hf, (size_t)(m + 1) * 4, cudaMemcpyHostToDevice, Ln.st));
            CK(cudaEventRecord(SF->ev, Ln.st)); SF->busy = true;
            Trial T{(uint32_t)m, (const uint32_t *)(D + c.o_first), (uint32_t *)(D + c.o_cs),
                    (uint32_t *)(D + c.o_ji), (int32_t *)(D + c.o_best)};
            r = enc_core(*C, Ln, Ln.st, (int)mm, D + c.o_in, xoff.data(), xsz.data(), methods + f0, nullptr,
                         D + c.o_out, c.bound_total, (uint64_t *)(D + c.o_off), (uint32_t *)(D + c.o_sz),
                         (uint64_t *)(D + c.o_tot), &T);
        } else
            r = enc_core(*C, Ln, Ln.st, m, D + c.o_in, ioff.data() + c.k0, in_size + c.k0, order + c.k0,
                         caps ? caps + c.k0 : nullptr, D + c.o_out, c.bound_total, (uint64_t *)(D + c.o_off),
                         (uint32_t *)(D + c.o_sz), (uint64_t *)(D + c.o_tot));
        if (r) return r;
        if ((r = Ln.hio.ensure((size_t)m * 16 + mm * 4 + 32))) return r;
        uint8_t *H = Ln.hio.p;
        CK(cudaMemcpyAsync(H, D + c.o_tot, 8, cudaMemcpyDeviceToHost, Ln.st));
        CK(cudaMemcpyAsync(H + 16, D + c.o_off, (size_t)m * 8, cudaMemcpyDeviceToHost, Ln.st));
        CK(cudaMemcpyAsync(H + 16 + (size_t)m * 8, D + c.o_sz, (size_t)m * 4, cudaMemcpyDeviceToHost, Ln.st));
        if (M) {
            CK(cudaMemcpyAsync(H + 16 + (size_t)m * 12, D + c.o_best, (size_t)m * 4, cudaMemcpyDeviceToHost, Ln.st));
            CK(cudaMemcpyAsync(H + 16 + (size_t)m * 16, D + c.o_cs, mm * 4, cudaMemcpyDeviceToHost, Ln.st));
        }
        return 0;
    };
    // sizes known: start the copy of exactly the bytes produced
    auto readback = [&](EncChunk &c) -> int {
        Lane &Ln = *c.L;
        int m = c.k1 - c.k0;
        CK(cudaStreamSynchronize(Ln.st));
        uint8_t *H = Ln.hio.p;
        uint64_t total = *(uint64_t *)H;
        const uint64_t *h_off = (const uint64_t *)(H + 16);
        const uint32_t *h_sz = (const uint32_t *)(H + 16 + (size_t)m * 8);
        c.base = out_base;
        if (out_base + total > out_cap) return B200RANS_ESPACE;
        if (total) CK(cudaMemcpyAsync(out + out_base, Ln.io.p + c.o_out, total, cudaMemcpyDeviceToHost, Ln.st));
        for (int k = c.k0; k < c.k1; k++) {
            out_off[k] = out_base + (size_t)h_off[k - c.k0];
            out_size[k] = h_sz[k - c.k0];
        }
        if (M) {
            const int32_t *h_best = (const int32_t *)(H + 16 + (size_t)m * 12);
            const uint32_t *h_cs = (const uint32_t *)(H + 16 + (size_t)m * 16);
            if (best) for (int k = c.k0; k < c.k1; k++) best[k] = h_best[k - c.k0];
            if (csize) memcpy(csize + mfirst[c.k0], h_cs, (size_t)(mfirst[c.k1] - mfirst[c.k0]) * 4);
        }
        out_base += total;
        return 0;
    };
    auto finish = [&](EncChunk &c) -> int { CK(cudaStreamSynchronize(c.L->st)); return 0; };

    int nc = (int)ch.size();
    const int depth = pipe_depth();
    for (int c = 0; c < nc && !rc; c++) {
        ch[c].L = &C->lane[c % NPIPE];
        if (c >= NPIPE) rc = finish(ch[c - NPIPE]);                  // the lane's previous chunk has left
        if (!rc) rc = submit(ch[c]);
        if (!rc && c >= depth) rc = readback(ch[c - depth]);
    }
    for (int c = nc > depth ? nc - depth : 0; c < nc && !rc; c++) rc = readback(ch[c]);
    for (auto &l : C->lane) cudaStreamSynchronize(l.st);
    return rc;
}

// host-side peek at a stream header: flag, stored length
bool peek_header(const unsigned char *in, unsigned int in_size, int *flag, uint32_t *ulen, int *hdr) {
    if (in_size == 0) return false;
    *flag = in[0];
    *hdr = 1;
    *ulen = 0;
    if (in[0] & X_NOSZ) return true;
    uint32_t v = 0;
    unsigned i = 1, cnt = 0;
    uint8_t c = 0x80;
    while ((c & 0x80) && i < in_size && cnt < 6) { c = in[i++]; v = (v << 7) | (c & 0x7f); cnt++; }
    *ulen = v;
    *hdr = (int)i;
    return true;
}

// ------------------------------------------------------------ host-buffer decode
struct DecChunk {
    int k0 = 0, k1 = 0;         // items
    int j0 = 0, j1 = 0;         // jobs
    Lane *L = nullptr;
    size_t o_in = 0, o_out = 0, o_scr = 0, o_osz = 0, o_st = 0;
};

int uncompress_batch_impl(int n, const unsigned char *const *in, const unsigned int *in_size,
                          unsigned char *const *out, unsigned int *out_size, int *status) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    if (n <= 0) return 0;
    // ---- plan: expand STRIPE streams into their sub-streams (host reads only headers)
    std::vector<DecItem> items(n);
    std::vector<const unsigned char *> jin;
    std::vector<uint32_t> jin_size, jcap;
    std::vector<uint8_t> jflag;
    std::vector<uint32_t> ocap(n, 0);
    std::vector<DecChunk> ch;
    {
        DecChunk c;
        size_t acc = 0;
        int dec_min_streams = chunk_min_streams();
        for (int k = 0; k < n; k++) {
            DecItem &it = items[k];
            int flag = 0, hdr = 0;
            uint32_t ulen = 0;
            it.first_job = (uint32_t)jin.size();
            if (!in[k] || !out[k] || !peek_header(in[k], in_size[k], &flag, &ulen, &hdr)) it.fail = true;
            else if (flag & X_STRIPE) {
                if (!stripe_plan_decode(it, in[k], in_size[k], out_size[k])) it.fail = true;
                else {
                    ocap[k] = it.ulen;
                    for (uint32_t s = 0; s < it.N; s++) {
                        jin.push_back(in[k] + it.sub_off[s]);
                        jin_size.push_back(it.sub_clen[s]);
                        jcap.push_back(it.sub_ulen[s]);
                        jflag.push_back(in[k][it.sub_off[s]]);
                    }
                }
            } else {
                if (flag & X_NOSZ) ulen = out_size[k];
                if (out_size[k] < ulen) it.fail = true;
                else {
                    ocap[k] = it.ulen = ulen;
                    jin.push_back(in[k]); jin_size.push_back(in_size[k]); jcap.push_back(ulen);
                    jflag.push_back((uint8_t)flag);
                }
            }
            it.njobs = (uint32_t)jin.size() - it.first_job;
            acc += ocap[k];
            if (it.njobs && (jflag.back() & (X_PACK | X_RLE))) dec_min_streams = std::max(dec_min_streams, chunk_min_streams_slow(true));
            if (it.njobs && !(jflag.back() & (X_32 | X_CAT)) && ocap[k] >= 32768)       // large 4-lane stream (see the encoder)
                dec_min_streams = std::max(dec_min_streams, chunk_min_streams_4lane());
            if ((acc >= chunk_bytes() && k - c.k0 + 1 >= dec_min_streams) || k - c.k0 + 1 >= CHUNK_MAX_STREAMS || k == n - 1) {
                c.k1 = k + 1; c.j1 = (int)jin.size();
                ch.push_back(c);
                c = DecChunk(); c.k0 = k + 1; c.j0 = (int)jin.size();
                acc = 0;
            }
        }
    }
    int nj = (int)jin.size();
    std::vector<uint64_t> cioff(nj ? nj : 1), jout(nj ? nj : 1), ooff(n);
    int rc = 0;

    auto submit = [&](DecChunk &c) -> int {
        Lane &Ln = *c.L;
        int mj = c.j1 - c.j0;
        // outputs: mirror host contiguity so that the copies back merge
        size_t o = 0, scratch = 0;
        for (int k = c.k0; k < c.k1; k++) {
            DecItem &it = items[k];
            if (it.fail) continue;
            int p = k - 1;
            while (p >= c.k0 && items[p].fail) p--;
            if (p >= c.k0 && out[k] == out[p] + ocap[p]) o = ooff[p] + ocap[p];
            else o = al(o, 256) + ((uintptr_t)out[k] & 15);
            ooff[k] = o;
            o += ocap[k];
            if (it.stripe) { it.o_tmp = scratch; scratch = al(scratch + it.ulen, 256); }
        }
        size_t in_total = 0;
        if (mj) place_spans(c.j0, c.j1, jin.data(), jin_size.data(), cioff.data(), &in_total);
        Layout L;
        c.o_in = L.take(in_total + 256); c.o_out = L.take(o + 256); c.o_scr = L.take(scratch + 256);
        c.o_osz = L.take((size_t)mj * 4 + 4); c.o_st = L.take((size_t)mj * 4 + 4);
        int r = Ln.io.ensure(L.off + 256);
        if (r) return r;
        uint8_t *D = Ln.io.p;
        for (int k = c.k0; k < c.k1; k++) {
            DecItem &it = items[k];
            if (it.fail) continue;
            if (it.stripe)      // sub-streams decode into scratch, joined afterwards
                for (uint32_t s = 0; s < it.N; s++) jout[it.first_job + s] = (c.o_scr + it.o_tmp + it.sub_idx[s]) - c.o_out;
            else jout[it.first_job] = ooff[k];
        }
        if (!mj) return 0;
        std::vector<Span> sp(mj);
        for (int j = c.j0; j < c.j1; j++) sp[j - c.j0] = Span{jin[j], c.o_in + cioff[j], jin_size[j]};
        if ((r = copy_spans(sp, D, true, Ln.st))) return r;
        r = dec_core(*C, Ln, Ln.st, mj, D + c.o_in, cioff.data() + c.j0, jin_size.data() + c.j0,
                     jflag.data() + c.j0, D + c.o_out, jout.data() + c.j0, jcap.data() + c.j0,
                     (uint32_t *)(D + c.o_osz), (int *)(D + c.o_st));
        if (r) return r;
        for (int k = c.k0; k < c.k1; k++) {
            DecItem &it = items[k];
            if (it.fail || !it.stripe) continue;
            CK(launch_stripe_join(D + c.o_scr + it.o_tmp, D + c.o_out + ooff[k], it.ulen, it.N, Ln.st));
            C->launches++;
        }
        if ((r = Ln.hio.ensure((size_t)mj * 8 + 32))) return r;
        CK(cudaMemcpyAsync(Ln.hio.p, D + c.o_osz, (size_t)mj * 4, cudaMemcpyDeviceToHost, Ln.st));
        CK(cudaMemcpyAsync(Ln.hio.p + (size_t)mj * 4 + 8, D + c.o_st, (size_t)mj * 4, cudaMemcpyDeviceToHost, Ln.st));
        return 0;
    };
    // verdict per item, then copy the good ones out
    auto readback = [&](DecChunk &c) -> int {
        Lane &Ln = *c.L;
        int mj = c.j1 - c.j0;
        CK(cudaStreamSynchronize(Ln.st));
        const uint32_t *h_osz = (const uint32_t *)Ln.hio.p;
        const int *h_st = (const int *)(Ln.hio.p + (size_t)mj * 4 + 8);
        std::vector<Span> sp;
        for (int k = c.k0; k < c.k1; k++) {
            DecItem &it = items[k];
            int s = it.fail ? ST_FAIL : ST_OK;
            uint32_t got = 0;
            if (!it.fail) {
                for (uint32_t j = 0; j < it.njobs; j++) {
                    uint32_t q = it.first_job + j - c.j0;
                    if (h_st[q] != ST_OK) s = h_st[q];
                    else if (it.stripe && h_osz[q] != it.sub_ulen[j]) s = ST_FAIL;
                    else if (!it.stripe) got = h_osz[q];
                }
                if (it.stripe) got = it.ulen;
            }
            if (status) status[k] = s;
            if (s == ST_OK) { out_size[k] = got; sp.push_back(Span{out[k], c.o_out + ooff[k], got}); }
            else out_size[k] = 0;
        }
        return copy_spans(sp, Ln.io.p, false, Ln.st);
    };
    auto finish = [&](DecChunk &c) -> int { CK(cudaStreamSynchronize(c.L->st)); return 0; };

    int nc = (int)ch.size();
    const int depth = pipe_depth();
    for (int c = 0; c < nc && !rc; c++) {
        ch[c].L = &C->lane[c % NPIPE];
        if (c >= NPIPE) rc = finish(ch[c - NPIPE]);
        if (!rc) rc = submit(ch[c]);
        if (!rc && c >= depth) rc = readback(ch[c - depth]);
    }
    for (int c = nc > depth ? nc - depth : 0; c < nc && !rc; c++) rc = readback(ch[c]);
    for (auto &l : C->lane) cudaStreamSynchronize(l.st);
    return rc;
}

}  // namespace b200rt

// ============================================================ C ABI: part 1
API unsigned int rans_compress_bound_4x16(unsigned int size, int order) {
    return compress_bound(size, order);
}

API void rans_set_cpu(int) {}

API unsigned char *rans_compress_to_4x16(unsigned char *in, unsigned int in_size, unsigned char *out,
                                         unsigned int *out_size, int order) {
    if (in_size > INT_MAX || (out && *out_size == 0)) { *out_size = 0; return NULL; }
    unsigned char *out_free = NULL;
    uint32_t cap;
    if (!out) {
        cap = compress_bound(in_size, order);
        if (!(out_free = out = (unsigned char *)malloc(cap))) { *out_size = 0; return NULL; }
    } else cap = *out_size;
    const unsigned char *ins[1] = {in ? in : (const unsigned char *)""};
    unsigned int isz[1] = {in_size}, osz[1] = {0};
    int ord[1] = {order};
    size_t ooff[1] = {0};
    // the stream leaves the device into a reusable pinned arena of bound size; exactly its bytes
    // are then handed to the caller (whose buffer may be smaller than the bound)
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) { free(out_free); *out_size = 0; return NULL; }
    const size_t acap = (size_t)std::max(cap, compress_bound(in_size, order)) + 64;
    if (C->single.ensure(acap)) { free(out_free); *out_size = 0; return NULL; }
    unsigned char *tmp = C->single.p;
    int r = compress_batch_impl(1, ins, isz, ord, &cap, tmp, acap, ooff, osz);
    if (r || osz[0] == 0 || osz[0] > cap) { free(out_free); *out_size = 0; return NULL; }
    memcpy(out, tmp + ooff[0], osz[0]);
    *out_size = osz[0];
    return out;
}

API unsigned char *rans_compress_4x16(unsigned char *in, unsigned int in_size, unsigned int *out_size,
                                      int order) {
    return rans_compress_to_4x16(in, in_size, NULL, out_size, order);
}

API unsigned char *rans_uncompress_to_4x16(unsigned char *in, unsigned int in_size, unsigned char *out,
                                           unsigned int *out_size) {
    int flag, hdr;
    uint32_t ulen;
    if (!in || !peek_header(in, in_size, &flag, &ulen, &hdr)) return NULL;
    unsigned char *out_free = NULL;
    if (flag & X_NOSZ) {
        if (!out) return NULL;
        ulen = *out_size;
    }
    if (!out) {
        if (ulen >= INT_MAX) return NULL;
        if (!(out_free = out = (unsigned char *)malloc(ulen ? ulen : 1))) return NULL;
        *out_size = ulen;
    } else if (*out_size < ulen) return NULL;
    const unsigned char *ins[1] = {in};
    unsigned char *outs[1] = {out};
    unsigned int isz[1] = {in_size}, osz[1] = {(flag & X_STRIPE) ? *out_size : ((flag & X_NOSZ) ? *out_size : ulen)};
    int st[1] = {0};
    int r = uncompress_batch_impl(1, ins, isz, outs, osz, st);
    if (r || st[0] != ST_OK) { free(out_free); return NULL; }
    *out_size = osz[0];
    return out;
}

API unsigned char *rans_uncompress_4x16(unsigned char *in, unsigned int in_size, unsigned int *out_size) {
    return rans_uncompress_to_4x16(in, in_size, NULL, out_size);
}

// ============================================================ C ABI: part 2
API int b200rans_set_device(int device) { tls_device = device; int e = 0; return get_ctx(&e) ? 0 : e; }

API int b200rans_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

API void *b200rans_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        fprintf(stderr, "libb200rans: cudaHostAlloc(%zu) failed\n", bytes);
        return nullptr;
    }
    return p;
}
API void b200rans_host_free(void *p) { if (p) cudaFreeHost(p); }

API int b200rans_compress_batch(int n, const unsigned char *const *in, const unsigned int *in_size,
                                const int *order, unsigned char *out, size_t out_cap, size_t *out_off,
                                unsigned int *out_size) {
    if (n < 0 || (n && (!in || !in_size || !order || !out || !out_off || !out_size))) return B200RANS_EINVAL;
    return compress_batch_impl(n, in, in_size, order, nullptr, out, out_cap, out_off, out_size);
}

API int b200rans_compress_trials(int n, const unsigned char *const *in, const unsigned int *in_size,
                                 const unsigned int *method_first, const int *methods, unsigned char *out,
                                 size_t out_cap, size_t *out_off, unsigned int *out_size, int *best,
                                 unsigned int *csize) {
    if (n < 0 || (n && (!in || !in_size || !method_first || !methods || !out || !out_off || !out_size)))
        return B200RANS_EINVAL;
    for (int k = 0; k < n; k++)
        if (method_first[k + 1] <= method_first[k] || method_first[k + 1] - method_first[k] > 64) return B200RANS_EINVAL;
    return compress_batch_impl(n, in, in_size, nullptr, nullptr, out, out_cap, out_off, out_size, method_first,
                               methods, best, csize);
}

API int b200rans_compress_methods_batch(int n, const unsigned char *const *in, const unsigned int *in_size,
                                        int n_methods, const int *methods, unsigned char *out, size_t out_cap,
                                        size_t *out_off, unsigned int *out_size, int *best, unsigned int *csize) {
    if (n < 0 || n_methods < 1 || n_methods > 64 || !methods) return B200RANS_EINVAL;
    std::vector<uint32_t> first((size_t)n + 1);
    std::vector<int> flat((size_t)n * n_methods);
    for (int k = 0; k <= n; k++) first[k] = (uint32_t)k * (uint32_t)n_methods;
    for (int k = 0; k < n; k++) memcpy(flat.data() + (size_t)k * n_methods, methods, sizeof(int) * n_methods);
    return b200rans_compress_trials(n, in, in_size, first.data(), flat.data(), out, out_cap, out_off, out_size,
                                    best, csize);
}

API unsigned char *b200rans_compress_methods(unsigned char *in, unsigned int in_size, int n_methods,
                                             const int *methods, unsigned int *out_size, int *best,
                                             unsigned int *csize) {
    if (out_size) *out_size = 0;
    if (!in || !out_size || n_methods < 1 || n_methods > 64 || !methods) return nullptr;
    size_t cap = 0;
    for (int j = 0; j < n_methods; j++) cap = std::max<size_t>(cap, compress_bound(in_size, methods[j]));
    cap += 64;
    unsigned char *out = (unsigned char *)malloc(cap);
    if (!out) return nullptr;
    const unsigned char *ins[1] = {in};
    size_t off = 0;
    unsigned int sz = 0;
    int b = -1;
    const uint32_t first[2] = {0, (uint32_t)n_methods};
    int r = compress_batch_impl(1, ins, &in_size, nullptr, nullptr, out, cap, &off, &sz, first, methods, &b,
                                csize);
    if (best) *best = b;
    if (r || !sz || b < 0) { free(out); return nullptr; }
    if (off) memmove(out, out + off, sz);
    *out_size = sz;
    return out;
}

API int b200rans_uncompress_batch(int n, const unsigned char *const *in, const unsigned int *in_size,
                                  unsigned char *const *out, unsigned int *out_size, int *status) {
    if (n < 0 || (n && (!in || !in_size || !out || !out_size))) return B200RANS_EINVAL;
    return uncompress_batch_impl(n, in, in_size, out, out_size, status);
}

API int64_t b200rans_uncompressed_size(const unsigned char *in, unsigned int in_size) {
    int flag, hdr;
    uint32_t ulen;
    if (!in || !peek_header(in, in_size, &flag, &ulen, &hdr) || (flag & X_NOSZ)) return -1;
    return ulen;
}

API size_t b200rans_compress_batch_dev_bound(int n, const unsigned int *in_size, const int *order) {
    size_t t = 0;
    for (int k = 0; k < n; k++) t += al(compress_bound(in_size[k], order[k]), 16) + 16;
    return t + 256;
}

API int b200rans_compress_batch_dev(void *stream, int n, const unsigned char *d_in, const uint64_t *in_off,
                                    const unsigned int *in_size, const int *order, unsigned char *d_out,
                                    size_t out_cap, uint64_t *d_out_off, unsigned int *d_out_size) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    if (n < 0 || (n && (!d_in || !in_off || !in_size || !order || !d_out))) return B200RANS_EINVAL;
    Lane &Ln = C->dlane;
    cudaStream_t st = stream ? (cudaStream_t)stream : Ln.st;
    return enc_core(*C, Ln, st, n, d_in, in_off, in_size, order, nullptr, d_out, out_cap, d_out_off, d_out_size,
                    nullptr);
}

API int b200rans_uncompress_batch_dev(void *stream, int n, const unsigned char *d_in, const uint64_t *in_off,
                                      const unsigned int *in_size, const unsigned char *flags,
                                      unsigned char *d_out,
                                      const uint64_t *out_off, const unsigned int *out_size,
                                      unsigned int *d_out_size, int *d_status) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    if (n < 0 || (n && (!d_in || !in_off || !in_size || !d_out || !out_off || !out_size || !d_out_size || !d_status)))
        return B200RANS_EINVAL;
    Lane &Ln = C->dlane;
    cudaStream_t st = stream ? (cudaStream_t)stream : Ln.st;
    return dec_core(*C, Ln, st, n, d_in, in_off, in_size, flags, d_out, out_off, out_size, d_out_size, d_status);
}

// ---------------------------------------------------------------- FASTQ split / join
static_assert(sizeof(b200fq_info) == sizeof(FqInfo), "b200fq_info mirrors FqInfo");

API size_t b200fq_split_scratch_bytes(uint32_t n, uint32_t max_records) {
    return fq_split_scratch_bytes(n, max_records);
}
API size_t b200fq_join_scratch_bytes(uint32_t name_len, uint32_t num_records) {
    return fq_join_scratch_bytes(name_len, num_records);
}

API int b200fq_split_dev(void *stream, const unsigned char *d_text, uint32_t n, unsigned char *d_name,
                         uint32_t name_cap, unsigned char *d_seq, unsigned char *d_qual, uint32_t seq_cap,
                         uint32_t *d_len, uint32_t *d_flag, uint32_t *d_name_off, uint32_t *d_seq_off,
                         uint32_t max_records, void *d_scratch, size_t scratch_bytes, b200fq_info *d_info) {
    return b200fq_split_dev_mode(stream, B200FQ_MODE_LOAD_SEQS, 0, d_text, n, d_name, name_cap, d_seq, d_qual, seq_cap,
                                 d_len, d_flag, d_name_off, d_seq_off, max_records, d_scratch, scratch_bytes, d_info);
}

API int b200fq_split_dev_mode(void *stream, int mode, uint32_t blk_size, const unsigned char *d_text, uint32_t n,
                              unsigned char *d_name, uint32_t name_cap, unsigned char *d_seq, unsigned char *d_qual,
                              uint32_t seq_cap, uint32_t *d_len, uint32_t *d_flag, uint32_t *d_name_off,
                              uint32_t *d_seq_off, uint32_t max_records, void *d_scratch, size_t scratch_bytes,
                              b200fq_info *d_info) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    if (mode != B200FQ_MODE_LOAD_SEQS && mode != B200FQ_MODE_KSEQ) return B200RANS_EINVAL;
    if (!d_text || !d_name || !d_seq || !d_qual || !d_len || !d_flag || !d_name_off || !d_seq_off || !d_scratch ||
        !d_info || n > 0x7fffffffu || ((uintptr_t)d_text & 15) || ((uintptr_t)d_scratch & 255) ||
        scratch_bytes < fq_split_scratch_bytes(n, max_records))
        return B200RANS_EINVAL;
    int l = 0;
    CK(fq_split_launch(d_text, n, d_name, d_seq, d_qual, name_cap, seq_cap, d_len, d_flag, d_name_off, d_seq_off,
                       max_records, (uint8_t *)d_scratch, (FqInfo *)d_info,
                       stream ? (cudaStream_t)stream : C->dlane.st, &l, mode == B200FQ_MODE_KSEQ, blk_size));
    C->launches += l;
    return 0;
}

API int b200fq_join_dev(void *stream, const unsigned char *d_name, uint32_t name_len, const unsigned char *d_seq,
                        const unsigned char *d_qual, const uint32_t *d_len, uint32_t num_records, int plus_name,
                        unsigned char *d_text, uint32_t text_cap, void *d_scratch, size_t scratch_bytes,
                        b200fq_info *d_info) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    if (!d_name || !d_seq || !d_qual || !d_len || !d_text || !d_scratch || !d_info || ((uintptr_t)d_name & 15) ||
        ((uintptr_t)d_scratch & 255) || scratch_bytes < fq_join_scratch_bytes(name_len, num_records))
        return B200RANS_EINVAL;
    int l = 0;
    CK(fq_join_launch(d_name, name_len, d_seq, d_qual, d_len, num_records, plus_name, d_text, text_cap,
                      (uint8_t *)d_scratch, (FqInfo *)d_info, stream ? (cudaStream_t)stream : C->dlane.st, &l));
    C->launches += l;
    return 0;
}

API int b200fq_split(const unsigned char *text, uint32_t n, unsigned char *name, uint32_t name_cap,
                     unsigned char *seq, unsigned char *qual, uint32_t seq_cap, uint32_t *len, uint32_t *flag,
                     uint32_t max_records, b200fq_info *info) {
    return b200fq_split_mode(B200FQ_MODE_LOAD_SEQS, 0, text, n, name, name_cap, seq, qual, seq_cap, len, flag,
                             max_records, info);
}

API int b200fq_split_mode(int mode, uint32_t blk_size, const unsigned char *text, uint32_t n, unsigned char *name,
                          uint32_t name_cap, unsigned char *seq, unsigned char *qual, uint32_t seq_cap, uint32_t *len,
                          uint32_t *flag, uint32_t max_records, b200fq_info *info) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    if (mode != B200FQ_MODE_LOAD_SEQS && mode != B200FQ_MODE_KSEQ) return B200RANS_EINVAL;
    if ((n && !text) || !name || !seq || !qual || !len || !flag || !info || n > 0x7fffffffu) return B200RANS_EINVAL;
    Lane &Ln = C->lane[0];
    cudaStream_t st = Ln.st;
    const size_t mr = max_records;
    Layout L;
    size_t o_text = L.take((size_t)n + 64), o_name = L.take((size_t)name_cap + 64);
    size_t o_seq = L.take((size_t)seq_cap + 64), o_qual = L.take((size_t)seq_cap + 64);
    size_t o_len = L.take(mr * 4 + 4), o_flag = L.take(mr * 4 + 4), o_no = L.take(mr * 4 + 4), o_so = L.take(mr * 4 + 4);
    size_t o_info = L.take(sizeof(FqInfo));
    size_t sb = fq_split_scratch_bytes(n, max_records);
    size_t o_scr = L.take(sb);
    int r = Ln.io.ensure(L.off + 256);
    if (r) return r;
    if ((r = Ln.hio.ensure(256))) return r;
    uint8_t *D = Ln.io.p;
    if (n) CK(cudaMemcpyAsync(D + o_text, text, n, cudaMemcpyHostToDevice, st));
    int l = 0;
    CK(fq_split_launch(D + o_text, n, D + o_name, D + o_seq, D + o_qual, name_cap, seq_cap, (uint32_t *)(D + o_len),
                       (uint32_t *)(D + o_flag), (uint32_t *)(D + o_no), (uint32_t *)(D + o_so), max_records,
                       D + o_scr, (FqInfo *)(D + o_info), st, &l, mode == B200FQ_MODE_KSEQ, blk_size));
    C->launches += l;
    CK(cudaMemcpyAsync(Ln.hio.p, D + o_info, sizeof(FqInfo), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(info, Ln.hio.p, sizeof(FqInfo));
    if (info->status) return 0;                 // the reference returns NULL; nothing to read back
    const size_t R = info->num_records;
    if (info->name_len) CK(cudaMemcpyAsync(name, D + o_name, info->name_len, cudaMemcpyDeviceToHost, st));
    if (info->seq_len) CK(cudaMemcpyAsync(seq, D + o_seq, info->seq_len, cudaMemcpyDeviceToHost, st));
    if (info->qual_len) CK(cudaMemcpyAsync(qual, D + o_qual, info->qual_len, cudaMemcpyDeviceToHost, st));
    if (R) {
        CK(cudaMemcpyAsync(len, D + o_len, R * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(flag, D + o_flag, R * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

API int b200fq_join(const unsigned char *name, uint32_t name_len, const unsigned char *seq,
                    const unsigned char *qual, uint32_t seq_len, const uint32_t *len, uint32_t num_records,
                    int plus_name, unsigned char *text, uint32_t text_cap, b200fq_info *info) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    if (!text || !info || (name_len && !name) || (seq_len && (!seq || !qual)) || (num_records && !len))
        return B200RANS_EINVAL;
    Lane &Ln = C->lane[0];
    cudaStream_t st = Ln.st;
    Layout L;
    size_t o_name = L.take((size_t)name_len + 64), o_seq = L.take((size_t)seq_len + 64);
    size_t o_qual = L.take((size_t)seq_len + 64), o_len = L.take((size_t)num_records * 4 + 4);
    size_t o_text = L.take((size_t)text_cap + 64), o_info = L.take(sizeof(FqInfo));
    size_t sb = fq_join_scratch_bytes(name_len, num_records);
    size_t o_scr = L.take(sb);
    int r = Ln.io.ensure(L.off + 256);
    if (r) return r;
    if ((r = Ln.hio.ensure(256))) return r;
    uint8_t *D = Ln.io.p;
    if (name_len) CK(cudaMemcpyAsync(D + o_name, name, name_len, cudaMemcpyHostToDevice, st));
    if (seq_len) {
        CK(cudaMemcpyAsync(D + o_seq, seq, seq_len, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(D + o_qual, qual, seq_len, cudaMemcpyHostToDevice, st));
    }
    if (num_records) CK(cudaMemcpyAsync(D + o_len, len, (size_t)num_records * 4, cudaMemcpyHostToDevice, st));
    int l = 0;
    CK(fq_join_launch(D + o_name, name_len, D + o_seq, D + o_qual, (const uint32_t *)(D + o_len), num_records,
                      plus_name, D + o_text, text_cap, D + o_scr, (FqInfo *)(D + o_info), st, &l));
    C->launches += l;
    CK(cudaMemcpyAsync(Ln.hio.p, D + o_info, sizeof(FqInfo), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(info, Ln.hio.p, sizeof(FqInfo));
    if (info->status) return 0;
    if (info->text_len) CK(cudaMemcpyAsync(text, D + o_text, info->text_len, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

// ---------------------------------------------------------------- CRC-32 and block framing
API int b200fqz_crc32_dev(void *stream, const unsigned char *d_buf, uint64_t n, uint32_t crc_in, uint32_t *d_crc) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    if ((n && !d_buf) || !d_crc) return B200RANS_EINVAL;
    Lane &Ln = C->dlane;
    cudaStream_t st = stream ? (cudaStream_t)stream : Ln.st;
    int r = Ln.crc.ensure(crc32_scratch_bytes(n) + 256);
    if (r) return r;
    int l = 0;
    CK(crc32_launch(d_buf, n, crc_in, d_crc, nullptr, 0, Ln.crc.p, st, &l));
    C->launches += l;
    return 0;
}

API int b200fqz_crc32(uint32_t crc_in, const unsigned char *buf, uint64_t n, uint32_t *crc_out) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    if ((n && !buf) || !crc_out) return B200RANS_EINVAL;
    Lane &Ln = C->lane[0];
    int r = Ln.io.ensure(n + 512);
    if (r) return r;
    if ((r = Ln.hio.ensure(256))) return r;
    if (n) CK(cudaMemcpyAsync(Ln.io.p + 256, buf, n, cudaMemcpyHostToDevice, Ln.st));
    if ((r = b200fqz_crc32_dev(Ln.st, Ln.io.p + 256, n, crc_in, (uint32_t *)Ln.io.p))) return r;
    CK(cudaMemcpyAsync(Ln.hio.p, Ln.io.p, 4, cudaMemcpyDeviceToHost, Ln.st));
    CK(cudaStreamSynchronize(Ln.st));
    *crc_out = *(uint32_t *)Ln.hio.p;
    return 0;
}

API int b200fqz_assemble_block_dev(void *stream, uint32_t num_records, int n_pieces, const b200fqz_piece *pieces,
                                   unsigned char *d_block, uint64_t block_cap, uint32_t *block_len) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    if (n_pieces < 0 || (n_pieces && !pieces) || !d_block || !block_len) return B200RANS_EINVAL;
    uint64_t total = 12;
    for (int i = 0; i < n_pieces; i++) {
        if (pieces[i].len && !pieces[i].ptr) return B200RANS_EINVAL;
        total += pieces[i].len;
    }
    if (total > block_cap || total > 0xffffffffull) return B200RANS_ESPACE;
    Lane &Ln = C->dlane;
    cudaStream_t st = stream ? (cudaStream_t)stream : Ln.st;
    int r = Ln.crc.ensure(crc32_scratch_bytes(total) + 256);
    if (r) return r;
    Stage *S;
    if ((r = Ln.get_stage(16, &S))) return r;
    // [block size][num_records][crc]: size and CRC are patched in by the finishing kernel
    uint32_t *h = (uint32_t *)S->h.p;
    h[0] = 0; h[1] = num_records; h[2] = 0;
    CK(cudaMemcpyAsync(d_block, h, 12, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(S->ev, st)); S->busy = true;
    uint64_t o = 12;
    for (int i = 0; i < n_pieces; i++) {
        if (pieces[i].len)
            CK(cudaMemcpyAsync(d_block + o, pieces[i].ptr, pieces[i].len,
                               pieces[i].on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        o += pieces[i].len;
    }
    int l = 0;
    CK(crc32_launch(d_block + 12, total - 12, 0, nullptr, d_block, (uint32_t)(total - 4), Ln.crc.p, st, &l));
    C->launches += l;
    *block_len = (uint32_t)total;
    return 0;
}

API uint64_t b200rans_launch_count(void) {
    return (tls_ctx ? tls_ctx->launches : 0) + g_worker_launches.load(std::memory_order_relaxed);
}

API int b200rans_set_profiling(int on) {
    int e = 0;
    Ctx *C = get_ctx(&e);
    if (!C) return e;
    C->prof = on != 0;
    return 0;
}
API float b200rans_last_kernel_ms(int which) {
    Ctx *C = tls_ctx;
    if (!C || which < 0 || which > 1 || !C->pe_valid[which]) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(C->pe[2 * which + 1]) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, C->pe[2 * which], C->pe[2 * which + 1]) != cudaSuccess) return -1.f;
    return ms;
}
API int b200rans_dec_staged_stats(unsigned long long out16[16], int reset) {
    int e = 0;
    Ctx *C = get_ctx(&e);
    if (!C) return e;
    if (!out16) return B200RANS_EINVAL;
    CK(cudaDeviceSynchronize());
    CK(dec_staged_stats(out16, reset != 0));
    return 0;
}
API const char *b200rans_version(void) { return "b200rans 0.1 (sm_100a)"; }
