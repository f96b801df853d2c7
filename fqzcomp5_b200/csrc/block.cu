// block.cu -- the callers either side of the codec, as the C ABI exposes them
// (include/b200rans.h parts 2 and 5):
//   * persistent per-device worker threads and the `_multi` entry points built on them
//     (the hts_tpool contract of thread_pool.c:113-164: independent jobs, results in
//     dispatch order; no collective);
//   * device-resident method trial and the in-slot encode (`_dev2`);
//   * one call per fqzcomp5 block: encode_block / decode_block (fqzcomp5.c:2147-2280,
//     :2290-2547) with the split, the method trial over section slices, the framing and
//     its CRC-32 all on the device;
//   * tok3's method tables for C callers (tokenise_name3.c:1283-1357).
// No codec arithmetic happens on the host.
#include <limits.h>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

#include "runtime.h"
#include "fastq.h"
#include "crc32.h"

using namespace b200;
using namespace b200rt;

#define API extern "C" __attribute__((visibility("default")))

namespace {

// ------------------------------------------------------------------ workers
// One thread per (device, slot), created on first use and kept: its thread-local context (streams,
// device and pinned arenas) therefore persists from call to call.  A call hands every worker one
// closure and waits for all of them.
struct Worker {
    int dev = 0, slot = 0;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false;
    int rc = 0;
    std::thread th;

    void loop() {
        set_thread_device(dev);
        for (;;) {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return has_job; });
            std::function<int()> f = std::move(job);
            lk.unlock();
            int e = 0;
            Ctx *C = get_ctx(&e);
            const uint64_t l0 = C ? C->launches : 0;
            int r = C ? f() : e;
            if (C) g_worker_launches.fetch_add(C->launches - l0, std::memory_order_relaxed);
            lk.lock();
            rc = r; has_job = false; done = true;
            lk.unlock();
            cv.notify_all();
        }
    }
    void submit(std::function<int()> f) {
        std::lock_guard<std::mutex> lk(m);
        job = std::move(f); has_job = true; done = false;
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return done; });
        done = false;
        return rc;
    }
};

struct WorkerPool {
    std::mutex call;                                     // one multi-GPU call at a time
    std::vector<std::unique_ptr<Worker>> w;              // [dev * WPD + slot]
    Worker *get(int dev, int slot) {
        const size_t i = (size_t)dev * B200RANS_WORKERS_PER_DEVICE + slot;
        if (w.size() <= i) w.resize(i + 1);
        if (!w[i]) {
            w[i].reset(new Worker());
            w[i]->dev = dev; w[i]->slot = slot;
            Worker *p = w[i].get();
            p->th = std::thread([p] { p->loop(); });
            p->th.detach();
        }
        return w[i].get();
    }
};
// never destroyed: the workers wait on it until the process ends
WorkerPool &pool() { static WorkerPool *p = new WorkerPool(); return *p; }

// fn(device, slot) on `slots` workers of each of the first ngpu devices; first failure wins
int run_on_workers(int ngpu, int slots, const std::function<int(int, int)> &fn) {
    int have = b200rans_device_count();
    if (have <= 0) { fprintf(stderr, "libb200rans: no CUDA device; there is no CPU path\n"); return B200RANS_ENODEV; }
    if (ngpu <= 0 || ngpu > have || slots < 1 || slots > B200RANS_WORKERS_PER_DEVICE) return B200RANS_EINVAL;
    WorkerPool &P = pool();
    std::lock_guard<std::mutex> lk(P.call);
    std::vector<Worker *> ws;
    for (int g = 0; g < ngpu; g++)
        for (int s = 0; s < slots; s++) {
            Worker *w = P.get(g, s);
            w->submit([&fn, g, s] { return fn(g, s); });
            ws.push_back(w);
        }
    int rc = 0;
    for (Worker *w : ws) { int r = w->wait(); if (r && !rc) rc = r; }
    return rc;
}

// workers per device the block calls use (each holds the arenas of a block in flight)
inline int block_workers() {
    static int v = env_int("B200RANS_BLOCK_WORKERS", 2, 1, B200RANS_WORKERS_PER_DEVICE);
    return v;
}

inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline void put32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }
inline uint32_t get32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int host_var_put(uint8_t *p, uint32_t v) {            // varint.h:205-237, most significant group first
    int n = 1;
    while (n < 5 && (v >> (7 * n))) n++;
    for (int k = n - 1; k >= 0; k--) *p++ = (uint8_t)(((v >> (7 * k)) & 0x7f) | (k ? 0x80 : 0));
    return n;
}
inline int host_var_get(const uint8_t *p, const uint8_t *end, uint32_t *v) {
    const uint8_t *s = p;
    uint32_t x = 0;
    int cnt = 0;
    uint8_t c = 0x80;
    while ((c & 0x80) && p < end && cnt < 6) { c = *p++; x = (x << 7) | (c & 0x7f); cnt++; }
    *v = x;
    return (int)(p - s);
}

__global__ void fill_u32_kernel(uint32_t *p, uint32_t n, uint32_t v) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------ encode one block
struct Sec { uint32_t ulen = 0, S = 0, nsl = 0, first = 0; std::vector<int> meth; size_t d_off = 0; };

int encode_block_impl(const unsigned char *text, uint32_t n, const b200fqz_block_opts *o, unsigned char *block,
                      size_t block_cap, b200fqz_block_report *rep) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    memset(rep, 0, sizeof(*rep));
    if ((n && !text) || !block || n > 0x7fffffffu) return B200RANS_EINVAL;
    if (o->n_name_methods < 0 || o->n_seq_methods < 1 || o->n_qual_methods < 1 ||
        o->n_name_methods > B200FQZ_MAX_METHODS || o->n_seq_methods > B200FQZ_MAX_METHODS ||
        o->n_qual_methods > B200FQZ_MAX_METHODS || (o->n_name_methods == 0 && !o->name_coder))
        return B200RANS_EINVAL;
    Lane &Ln = C->dlane;
    cudaStream_t st = Ln.st;
    const double t0 = now_ms();

    // ---- 1. one layout for everything the block needs, then text to the device and the split
    // (load_seqs, fqzcomp5.c:279-410).  Sections total at most n bytes; with a fixed read length a slice
    // shrinks to whole reads (never below half of slice_bytes), so there are at most 2n/S + 6 slices.
    uint32_t mr = n / 16 + 64;
    FqInfo info;
    const uint32_t name_cap = n + 64, seq_cap = n / 2 + 64;
    const size_t max_inputs = (o->slice_bytes ? 2 * ((size_t)n / o->slice_bytes) : 0) + 6;
    const size_t max_mm = max_inputs * B200FQZ_MAX_METHODS;
    const size_t out_cap = (size_t)n + n / 16 + max_inputs * 2048 + 4096;
    const size_t blk_cap = out_cap + 4096 + 12 * max_inputs + (size_t)n / 4 + 64 +
                           (o->n_name_methods ? 0 : 2 * (size_t)n + 4096);
    size_t o_text = 0, o_name = 0, o_seq = 0, o_qual = 0, o_len = 0, o_flag = 0, o_info = 0;
    size_t o_out = 0, o_off = 0, o_sz = 0, o_tot = 0, o_cs = 0, o_ji = 0, o_best = 0, o_first = 0, o_blk = 0;
    int r;
    if ((r = C->hblk.ensure(4096))) return r;
    for (int attempt = 0;; attempt++) {
        Layout L;
        o_text = L.take((size_t)n + 64); o_name = L.take((size_t)name_cap + 64);
        o_seq = L.take((size_t)seq_cap + 64); o_qual = L.take((size_t)seq_cap + 64);
        o_len = L.take((size_t)mr * 4 + 4); o_flag = L.take((size_t)mr * 4 + 4);
        const size_t o_no = L.take((size_t)mr * 4 + 4), o_so = L.take((size_t)mr * 4 + 4);
        o_info = L.take(sizeof(FqInfo));
        const size_t sb = fq_split_scratch_bytes(n, mr);
        const size_t o_scr = L.take(sb);
        o_out = L.take(out_cap); o_off = L.take(max_inputs * 8); o_sz = L.take(max_inputs * 4); o_tot = L.take(16);
        o_cs = L.take(max_mm * 4); o_ji = L.take(max_mm * 4); o_best = L.take(max_inputs * 4);
        o_first = L.take((max_inputs + 1) * 4);
        o_blk = L.take(blk_cap);
        if ((r = C->blk.ensure(L.off + 256))) return r;
        uint8_t *D0 = C->blk.p;
        if (n) CK(cudaMemcpyAsync(D0 + o_text, text, n, cudaMemcpyHostToDevice, st));
        int l = 0;
        CK(fq_split_launch(D0 + o_text, n, D0 + o_name, D0 + o_seq, D0 + o_qual, name_cap, seq_cap,
                           (uint32_t *)(D0 + o_len), (uint32_t *)(D0 + o_flag), (uint32_t *)(D0 + o_no),
                           (uint32_t *)(D0 + o_so), mr, D0 + o_scr, (FqInfo *)(D0 + o_info), st, &l,
                           o->kseq_blk_size != 0, o->kseq_blk_size));
        C->launches += l;
        CK(cudaMemcpyAsync(C->hblk.p, D0 + o_info, sizeof(FqInfo), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        memcpy(&info, C->hblk.p, sizeof(info));
        if (info.status == 2 && attempt == 0) { mr = n / 6 + 64; continue; }    // more records than guessed
        break;
    }
    rep->status = info.status;
    if (info.status) return 0;                       // the reference returns NULL for this block
    rep->num_records = info.num_records; rep->consumed = info.consumed; rep->fixed_len = info.fixed_len;
    const uint32_t R = info.num_records;
    uint8_t *D = C->blk.p;
    const double t1 = now_ms();

    // ---- 2. plan the slices and their method lists
    Sec sec[3];
    sec[0].ulen = info.name_len; sec[1].ulen = info.seq_len; sec[2].ulen = info.qual_len;
    sec[0].d_off = o_name; sec[1].d_off = o_seq; sec[2].d_off = o_qual;
    const int *lists[3] = {o->name_methods, o->seq_methods, o->qual_methods};
    const int nlist[3] = {o->n_name_methods, o->n_seq_methods, o->n_qual_methods};
    uint32_t ninputs = 0;
    for (int s = 0; s < 3; s++) {
        Sec &X = sec[s];
        rep->ulen[s] = X.ulen;
        for (int j = 0; j < nlist[s]; j++) {
            int m = lists[s][j];
            if (m == B200FQZ_RANSXN1) {              // fqzcomp5.c:2013-2022
                if (info.fixed_len <= 0) continue;
                m = (info.fixed_len << 8) + 9;
            }
            X.meth.push_back(m);
        }
        if (s == 0 && nlist[0] == 0) continue;       // names go through the caller's coder
        if (X.meth.empty()) return B200RANS_EINVAL;
        uint32_t S = o->slice_bytes ? o->slice_bytes : X.ulen;
        if (s && info.fixed_len > 0 && S > (uint32_t)info.fixed_len) S -= S % (uint32_t)info.fixed_len;
        if (!S) S = 1;
        X.S = S;
        X.nsl = X.ulen ? (uint32_t)(((uint64_t)X.ulen + S - 1) / S) : 1;
        X.first = ninputs;
        ninputs += X.nsl;
        rep->nslices[s] = X.nsl;
    }
    std::vector<uint32_t> mfirst(ninputs + 1);
    std::vector<int> methods;
    std::vector<uint64_t> xoff;
    std::vector<uint32_t> xsz;
    if (ninputs > max_inputs) return B200RANS_ESPACE;
    {
        uint32_t k = 0;
        for (int s = 0; s < 3; s++) {
            Sec &X = sec[s];
            for (uint32_t i = 0; i < X.nsl; i++, k++) {
                const uint32_t off = i * X.S, len = std::min(X.S, X.ulen - off);
                mfirst[k] = (uint32_t)methods.size();
                for (int m : X.meth) { methods.push_back(m); xoff.push_back(X.d_off + off); xsz.push_back(len); }
            }
        }
        mfirst[ninputs] = (uint32_t)methods.size();
    }
    const size_t mm = methods.size();
    // ---- 3. names (and flags) start their way to the host for the caller's coder
    const size_t h_need = 4096 + (size_t)ninputs * 8 + mm * 4 + 64 +
                          (nlist[0] ? 0 : (size_t)info.name_len + (size_t)R * 4 + 64) +
                          (info.fixed_len > 0 ? 0 : (size_t)R * 4 + 64);
    if ((r = C->hblk.ensure(h_need))) return r;
    uint8_t *H = C->hblk.p;
    size_t ho = 4096;
    const size_t h_sz = ho; ho += (size_t)ninputs * 4;
    const size_t h_best = ho; ho += (size_t)ninputs * 4;
    const size_t h_cs = ho; ho += mm * 4 + 64;
    size_t h_names = 0, h_flags = 0, h_len = 0;
    if (!nlist[0]) {
        h_names = ho; ho += ((size_t)info.name_len + 63) & ~(size_t)63;
        h_flags = ho; ho += (size_t)R * 4 + 64;
        if (info.name_len) CK(cudaMemcpyAsync(H + h_names, D + o_name, info.name_len, cudaMemcpyDeviceToHost, st));
        if (R) CK(cudaMemcpyAsync(H + h_flags, D + o_flag, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
    }
    if (info.fixed_len <= 0) {
        h_len = ho; ho += (size_t)R * 4 + 64;
        if (R) CK(cudaMemcpyAsync(H + h_len, D + o_len, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaEventRecord(C->ev_blk, st));              // names, flags and lengths are on the host once this has passed
    const double t2 = now_ms();

    // ---- 4. the method trial over every slice of every section: one batch
    if (ninputs) {
        Stage *SF;
        if ((r = Ln.get_stage((size_t)(ninputs + 1) * 4, &SF))) return r;
        memcpy(SF->h.p, mfirst.data(), (size_t)(ninputs + 1) * 4);
        CK(cudaMemcpyAsync(D + o_first, SF->h.p, (size_t)(ninputs + 1) * 4, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(SF->ev, st)); SF->busy = true;
        Trial T{ninputs, (const uint32_t *)(D + o_first), (uint32_t *)(D + o_cs), (uint32_t *)(D + o_ji),
                (int32_t *)(D + o_best)};
        r = enc_core(*C, Ln, st, (int)mm, D, xoff.data(), xsz.data(), methods.data(), nullptr, D + o_out, out_cap,
                     (uint64_t *)(D + o_off), (uint32_t *)(D + o_sz), (uint64_t *)(D + o_tot), &T, 1, false);
        if (r) return r;
        CK(cudaMemcpyAsync(H + h_sz, D + o_sz, (size_t)ninputs * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(H + h_best, D + o_best, (size_t)ninputs * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(H + h_cs, D + o_cs, mm * 4, cudaMemcpyDeviceToHost, st));
    }
    const double t2b = now_ms();                     // everything is queued; from here on the host only waits
    // ---- 5. the caller's name coder runs on this thread while the device works
    unsigned char *name_sec = nullptr;
    uint32_t name_sec_len = 0;
    if (!nlist[0]) {
        CK(cudaEventSynchronize(C->ev_blk));
        int e = o->name_coder(o->name_user, H + h_names, info.name_len, (const uint32_t *)(H + h_flags), R, &name_sec,
                              &name_sec_len);
        if (e || !name_sec) { free(name_sec); cudaStreamSynchronize(st); return B200RANS_EINVAL; }
    }
    CK(cudaStreamSynchronize(st));
    const double t3 = now_ms();

    // ---- 6. framing: [u32 size][u32 num_records][u32 crc] + sections (fqzcomp5.c:2147-2280)
    const uint32_t *hs = (const uint32_t *)(H + h_sz);
    const int32_t *hb = (const int32_t *)(H + h_best);
    const uint32_t *hc = (const uint32_t *)(H + h_cs);
    std::vector<uint8_t> hp;                         // host-built bytes of the block, in order
    struct Piece { bool dev; size_t off; size_t len; };
    std::vector<Piece> pieces;
    auto host_piece = [&](size_t from) { if (hp.size() > from) pieces.push_back(Piece{false, from, hp.size() - from}); };
    hp.resize(12);
    put32(hp.data() + 4, R);
    size_t mark = 0;
    uint64_t dev_cursor = 0;                         // winners are packed without gaps, in input order
    for (int s = 0; s < 3; s++) {
        Sec &X = sec[s];
        if (s == 0 && !nlist[0]) {
            host_piece(mark); mark = hp.size();
            pieces.push_back(Piece{false, (size_t)-1, name_sec_len});      // the caller's bytes
            rep->clen[0] = name_sec_len;
        } else {
            uint64_t clen = 0;
            for (uint32_t i = 0; i < X.nsl; i++) {
                const uint32_t k = X.first + i;
                if (!hs[k] || hb[k] < 0) { free(name_sec); rep->status = 1; return 0; }
                clen += hs[k];
                rep->wins[s][hb[k]]++;
                for (size_t j = 0; j < X.meth.size(); j++) rep->csize[s][j] += hc[mfirst[k] + j];
            }
            const bool sliced = o->slice_bytes != 0;
            const uint64_t pay = clen + (sliced ? 8 + 4ull * X.nsl : 0);
            if (pay > 0xffffffffull) { free(name_sec); return B200RANS_ESPACE; }
            uint8_t meta[9];
            const uint8_t strat = sliced ? B200FQZ_STRAT_SLICED : (s == 0 ? 0xB0 : 0);
            if (s == 0) { put32(meta, X.ulen); meta[4] = strat; put32(meta + 5, (uint32_t)pay); }    // :1415-1416
            else { meta[0] = strat; put32(meta + 1, X.ulen); put32(meta + 5, (uint32_t)pay); }       // :2222-2225
            hp.insert(hp.end(), meta, meta + 9);
            if (sliced) {
                uint8_t w[4];
                put32(w, X.nsl); hp.insert(hp.end(), w, w + 4);
                put32(w, X.S); hp.insert(hp.end(), w, w + 4);
                for (uint32_t i = 0; i < X.nsl; i++) { put32(w, hs[X.first + i]); hp.insert(hp.end(), w, w + 4); }
            }
            host_piece(mark); mark = hp.size();
            pieces.push_back(Piece{true, (size_t)dev_cursor, (size_t)clen});
            dev_cursor += clen;
            rep->clen[s] = (uint32_t)pay;
        }
        if (s == 0) {                                // read lengths follow the names (:2190-2214)
            if (info.fixed_len > 0) {
                uint8_t buf[6];
                int nb = 1 + host_var_put(buf + 1, (uint32_t)info.fixed_len);
                buf[0] = (uint8_t)(nb - 1);
                hp.insert(hp.end(), buf, buf + nb);
            } else {
                const uint32_t *len = (const uint32_t *)(H + h_len);
                const size_t at = hp.size();
                hp.resize(at + 5 + (size_t)R * 5);
                size_t nb = 5;
                hp[at] = 0;
                for (uint32_t i = 0; i < R; i++) nb += host_var_put(hp.data() + at + nb, len[i]);
                put32(hp.data() + at + 1, (uint32_t)(nb - 5));
                hp.resize(at + nb);
            }
        }
    }
    host_piece(mark);
    uint64_t total = 0;
    for (auto &p : pieces) total += p.len;
    if (total > block_cap || total > blk_cap || total > 0xffffffffull) { free(name_sec); return B200RANS_ESPACE; }
    if (dev_cursor > out_cap) { free(name_sec); return B200RANS_ESPACE; }
    {
        Stage *SH;
        if ((r = Ln.get_stage(hp.size() + 64, &SH))) { free(name_sec); return r; }
        memcpy(SH->h.p, hp.data(), hp.size());
        uint8_t *B = D + o_blk;
        uint64_t at = 0;
        for (auto &p : pieces) {
            if (p.len) {
                if (p.dev) CK(cudaMemcpyAsync(B + at, D + o_out + p.off, p.len, cudaMemcpyDeviceToDevice, st));
                else if (p.off == (size_t)-1) CK(cudaMemcpyAsync(B + at, name_sec, p.len, cudaMemcpyHostToDevice, st));
                else CK(cudaMemcpyAsync(B + at, SH->h.p + p.off, p.len, cudaMemcpyHostToDevice, st));
            }
            at += p.len;
        }
        CK(cudaEventRecord(SH->ev, st)); SH->busy = true;
        if ((r = Ln.crc.ensure(crc32_scratch_bytes(total) + 256))) { free(name_sec); return r; }
        int l = 0;
        // CRC over everything after the CRC field; size and CRC are patched in by the finishing kernel
        CK(crc32_launch(B + 12, total - 12, 0, (uint32_t *)(D + o_tot), B, (uint32_t)(total - 4), Ln.crc.p, st, &l));
        C->launches += l;
        CK(cudaMemcpyAsync(block, B, total, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(H, D + o_tot, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    free(name_sec);
    rep->block_len = (uint32_t)total;
    rep->crc = get32(H);
    const double t4 = now_ms();
    (void)t2;
    rep->ms[0] = (float)(t1 - t0); rep->ms[1] = (float)(t2b - t1); rep->ms[2] = (float)(t3 - t2b); rep->ms[3] = (float)(t4 - t3);
    return 0;
}

// ------------------------------------------------------------------ decode one block
int decode_block_impl(const unsigned char *block, uint32_t block_len, int plus_name, unsigned char *text,
                      size_t text_cap, b200fqz_block_report *rep) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    memset(rep, 0, sizeof(*rep));
    if (!block || !text) return B200RANS_EINVAL;
    rep->status = 1;
    if (block_len < 12 + 9 + 1 + 9 + 9) return 0;
    if (get32(block) != block_len - 4) return 0;                   // :2302
    const uint32_t R = get32(block + 4);
    rep->num_records = R;
    rep->crc = get32(block + 8);
    Lane &Ln = C->dlane;
    cudaStream_t st = Ln.st;

    // ---- parse the framing (host reads only the headers)
    struct DSec { uint32_t ulen = 0, S = 0, nsl = 0; std::vector<uint32_t> csz; size_t pay = 0; };
    DSec sec[3];
    const uint8_t *p = block + 12, *end = block + block_len;
    std::vector<uint32_t> lens;
    int32_t fixed_len = 0;
    for (int s = 0; s < 3; s++) {
        DSec &X = sec[s];
        if (end - p < 9) return 0;
        uint8_t strat;
        uint32_t clen;
        if (s == 0) { X.ulen = get32(p); strat = p[4]; clen = get32(p + 5); }
        else { strat = p[0]; X.ulen = get32(p + 1); clen = get32(p + 5); }
        p += 9;
        if ((size_t)(end - p) < clen) return 0;
        const uint8_t *q = p, *qend = p + clen;
        if (strat == B200FQZ_STRAT_SLICED) {
            if (clen < 8) return 0;
            X.nsl = get32(q); X.S = get32(q + 4); q += 8;
            if ((uint64_t)X.nsl * 4 > (uint64_t)(qend - q) || !X.S) return 0;
            if ((uint64_t)X.nsl * X.S < X.ulen || (X.nsl > 1 && (uint64_t)(X.nsl - 1) * X.S >= X.ulen)) return 0;
            uint64_t tot = 0;
            for (uint32_t i = 0; i < X.nsl; i++, q += 4) { X.csz.push_back(get32(q)); tot += X.csz.back(); }
            if (tot != (uint64_t)(qend - q)) return 0;
        } else if (strat == (s == 0 ? 0xB0 : 0)) {
            X.nsl = 1; X.S = X.ulen ? X.ulen : 1; X.csz.push_back(clen);
        } else return 0;                             // a section this path did not write (tok3, LZP, fqz models)
        X.pay = (size_t)(q - block);
        rep->ulen[s] = X.ulen; rep->clen[s] = clen; rep->nslices[s] = X.nsl;
        p = qend;
        if (s == 0) {                                // read lengths (:2381-2404)
            if (p >= end) return 0;
            const uint8_t c = *p++;
            if (c > 0) {
                uint32_t len;
                int vl = host_var_get(p, end, &len);
                if (!vl) return 0;
                p += vl;
                fixed_len = (int32_t)len;
            } else {
                if (end - p < 4) return 0;
                p += 4;
                lens.resize(R);
                for (uint32_t i = 0; i < R; i++) {
                    int vl = host_var_get(p, end, &lens[i]);
                    if (!vl) return 0;
                    p += vl;
                }
            }
        }
    }
    if (p != end || sec[1].ulen != sec[2].ulen) return 0;
    rep->fixed_len = fixed_len;
    const uint64_t text_bound = (uint64_t)sec[0].ulen * (plus_name ? 2 : 1) + 2ull * sec[1].ulen + 6ull * R + 64;
    if (text_bound > 0xffffffffull) return B200RANS_ESPACE;

    // ---- device layout: block, fields, text
    Layout L;
    const size_t o_blk = L.take((size_t)block_len + 64);
    size_t o_f[3];
    for (int s = 0; s < 3; s++) o_f[s] = L.take((size_t)sec[s].ulen + 64);
    const size_t o_len = L.take((size_t)R * 4 + 4);
    uint32_t nj = sec[0].nsl + sec[1].nsl + sec[2].nsl;
    const size_t o_osz = L.take((size_t)nj * 4 + 4), o_st = L.take((size_t)nj * 4 + 4), o_crc = L.take(16);
    const size_t o_info = L.take(sizeof(FqInfo));
    const size_t jsb = fq_join_scratch_bytes(sec[0].ulen, R);
    const size_t o_scr = L.take(jsb);
    const size_t o_text = L.take((size_t)text_bound + 64);
    int r;
    if ((r = C->blk.ensure(L.off + 256))) return r;
    if ((r = C->hblk.ensure(4096 + (size_t)nj * 8 + 64))) return r;
    uint8_t *D = C->blk.p, *H = C->hblk.p;
    CK(cudaMemcpyAsync(D + o_blk, block, block_len, cudaMemcpyHostToDevice, st));
    // ---- CRC of everything after the CRC field (:2306-2318)
    if ((r = Ln.crc.ensure(crc32_scratch_bytes(block_len) + 256))) return r;
    int l = 0;
    CK(crc32_launch(D + o_blk + 12, block_len - 12, 0, (uint32_t *)(D + o_crc), nullptr, 0, Ln.crc.p, st, &l));
    C->launches += l;
    CK(cudaMemcpyAsync(H, D + o_crc, 4, cudaMemcpyDeviceToHost, st));
    // ---- every slice of every section in one decode batch
    std::vector<uint64_t> ioff(nj), ooff(nj);
    std::vector<uint32_t> isz(nj), ocap(nj);
    std::vector<uint8_t> flags(nj);
    {
        uint32_t k = 0;
        for (int s = 0; s < 3; s++) {
            DSec &X = sec[s];
            size_t at = X.pay;
            for (uint32_t i = 0; i < X.nsl; i++, k++) {
                ioff[k] = o_blk + at; isz[k] = X.csz[i];
                const uint32_t off = i * X.S;
                ooff[k] = o_f[s] + off; ocap[k] = std::min(X.S, X.ulen - std::min(off, X.ulen));
                flags[k] = X.csz[i] ? block[at] : 0;
                if (flags[k] & X_STRIPE) flags[k] = 0xff;            // handled below
                at += X.csz[i];
            }
        }
    }
    // STRIPE streams carry a sub-stream table that the host-buffer decoder plans from the header:
    // RANSXN1 winners go through it, everything else through the device-resident batch.
    std::vector<uint32_t> plain, striped;
    for (uint32_t k = 0; k < nj; k++) (flags[k] == 0xff ? striped : plain).push_back(k);
    if (!plain.empty()) {
        std::vector<uint64_t> io2, oo2; std::vector<uint32_t> is2, oc2; std::vector<uint8_t> f2;
        for (uint32_t k : plain) { io2.push_back(ioff[k]); oo2.push_back(ooff[k]); is2.push_back(isz[k]); oc2.push_back(ocap[k]); f2.push_back(flags[k]); }
        r = dec_core(*C, Ln, st, (int)plain.size(), D, io2.data(), is2.data(), f2.data(), D, oo2.data(), oc2.data(),
                     (uint32_t *)(D + o_osz), (int *)(D + o_st));
        if (r) return r;
        CK(cudaMemcpyAsync(H + 64, D + o_osz, plain.size() * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(H + 64 + (size_t)nj * 4, D + o_st, plain.size() * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    if (get32(H) != rep->crc) { rep->status = 3; return 0; }       // "Block CRC mismatch" (:2312-2317)
    {
        const uint32_t *hz = (const uint32_t *)(H + 64);
        const int *hst = (const int *)(H + 64 + (size_t)nj * 4);
        for (size_t i = 0; i < plain.size(); i++)
            if (hst[i] != ST_OK || hz[i] != ocap[plain[i]]) return 0;
    }
    if (!striped.empty()) {
        // decoded on the host-buffer path straight from the caller's block into a pinned bounce buffer,
        // then placed in the field buffer
        size_t tot = 0;
        for (uint32_t k : striped) tot += ocap[k];
        std::vector<const unsigned char *> ins;
        std::vector<unsigned char *> outs;
        std::vector<uint32_t> is2, os2;
        std::vector<int> st2(striped.size());
        Arena &bounce = C->single;                   // pinned, kept from call to call
        if ((r = bounce.ensure(tot + 64))) return r;
        size_t at = 0;
        for (uint32_t k : striped) {
            ins.push_back(block + (ioff[k] - o_blk)); is2.push_back(isz[k]);
            outs.push_back(bounce.p + at); os2.push_back(ocap[k]); at += ocap[k];
        }
        r = uncompress_batch_impl((int)striped.size(), ins.data(), is2.data(), outs.data(), os2.data(), st2.data());
        if (r) return r;
        at = 0;
        bool ok = true;
        for (size_t i = 0; i < striped.size(); i++) {
            const uint32_t k = striped[i];
            if (st2[i] != ST_OK || os2[i] != ocap[k]) { ok = false; break; }
            CK(cudaMemcpyAsync(D + ooff[k], bounce.p + at, ocap[k], cudaMemcpyHostToDevice, st));
            at += ocap[k];
        }
        CK(cudaStreamSynchronize(st));
        if (!ok) return 0;
    }
    // ---- read lengths, then output_fastq (:3440-3480) with the +33 of :2532-2533
    if (fixed_len > 0) {
        if (R) { fill_u32_kernel<<<(R + 255) / 256, 256, 0, st>>>((uint32_t *)(D + o_len), R, (uint32_t)fixed_len); C->launches++; }
    } else if (R) {
        Stage *SL;
        if ((r = Ln.get_stage((size_t)R * 4, &SL))) return r;
        memcpy(SL->h.p, lens.data(), (size_t)R * 4);
        CK(cudaMemcpyAsync(D + o_len, SL->h.p, (size_t)R * 4, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(SL->ev, st)); SL->busy = true;
    }
    l = 0;
    CK(fq_join_launch(D + o_f[0], sec[0].ulen, D + o_f[1], D + o_f[2], (const uint32_t *)(D + o_len), R, plus_name,
                      D + o_text, (uint32_t)text_bound, D + o_scr, (FqInfo *)(D + o_info), st, &l));
    C->launches += l;
    CK(cudaMemcpyAsync(H + 32, D + o_info, sizeof(FqInfo), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    FqInfo info;
    memcpy(&info, H + 32, sizeof(info));
    if (info.status) return 0;
    if (info.text_len > text_cap) return B200RANS_ESPACE;
    if (info.text_len) CK(cudaMemcpyAsync(text, D + o_text, info.text_len, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    rep->block_len = info.text_len;
    rep->status = 0;
    return 0;
}

}  // namespace

// ============================================================ C ABI: part 2 (device-resident, second form)
API size_t b200rans_compress_slots_bound(int n, const unsigned int *in_size, const int *order) {
    return (n > 0 && in_size && order) ? enc_slots_bound(n, in_size, order) : 256;
}

API int b200rans_compress_batch_dev2(void *stream, int n, const unsigned char *d_in, const uint64_t *in_off,
                                     const unsigned int *in_size, const int *order, unsigned char *d_out,
                                     size_t out_cap, uint64_t *d_out_off, unsigned int *d_out_size,
                                     unsigned int flags) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    if (n < 0 || (n && (!d_in || !in_off || !in_size || !order || !d_out)) || (flags & ~B200RANS_OUT_IN_SLOT))
        return B200RANS_EINVAL;
    const bool inslot = (flags & B200RANS_OUT_IN_SLOT) != 0;
    if (inslot && n && (!d_out_off || !d_out_size)) return B200RANS_EINVAL;
    Lane &Ln = C->dlane;
    cudaStream_t st = stream ? (cudaStream_t)stream : Ln.st;
    return enc_core(*C, Ln, st, n, d_in, in_off, in_size, order, nullptr, d_out, out_cap, d_out_off, d_out_size,
                    nullptr, nullptr, 16, inslot);
}

API int b200rans_compress_trials_dev(void *stream, int n, const unsigned char *d_in, const uint64_t *in_off,
                                     const unsigned int *in_size, const unsigned int *method_first,
                                     const int *methods, unsigned char *d_out, size_t out_cap,
                                     unsigned int pack_align, uint64_t *d_out_off, unsigned int *d_out_size,
                                     int *d_best, unsigned int *d_csize) {
    int err = 0;
    Ctx *C = get_ctx(&err);
    if (!C) return err;
    if (n < 0 || (n && (!d_in || !in_off || !in_size || !method_first || !methods || !d_out || !d_out_off ||
                        !d_out_size)))
        return B200RANS_EINVAL;
    if (n == 0) return 0;
    for (int k = 0; k < n; k++)
        if (method_first[k + 1] <= method_first[k] || method_first[k + 1] - method_first[k] > 64) return B200RANS_EINVAL;
    Lane &Ln = C->dlane;
    cudaStream_t st = stream ? (cudaStream_t)stream : Ln.st;
    const size_t mm = method_first[n];
    std::vector<uint64_t> xoff(mm);
    std::vector<uint32_t> xsz(mm);
    for (int k = 0; k < n; k++)
        for (uint32_t c = method_first[k]; c < method_first[k + 1]; c++) { xoff[c] = in_off[k]; xsz[c] = in_size[k]; }
    // candidate sizes, job index scratch, winners and group offsets live in the lane's io arena
    Layout L;
    const size_t o_cs = L.take(mm * 4), o_ji = L.take(mm * 4), o_best = L.take((size_t)n * 4);
    const size_t o_first = L.take((size_t)(n + 1) * 4);
    int r = Ln.io.ensure(L.off + 256);
    if (r) return r;
    uint8_t *D = Ln.io.p;
    Stage *SF;
    if ((r = Ln.get_stage((size_t)(n + 1) * 4, &SF))) return r;
    memcpy(SF->h.p, method_first, (size_t)(n + 1) * 4);
    CK(cudaMemcpyAsync(D + o_first, SF->h.p, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(SF->ev, st)); SF->busy = true;
    Trial T{(uint32_t)n, (const uint32_t *)(D + o_first), d_csize ? d_csize : (uint32_t *)(D + o_cs),
            (uint32_t *)(D + o_ji), d_best ? d_best : (int32_t *)(D + o_best)};
    return enc_core(*C, Ln, st, (int)mm, d_in, xoff.data(), xsz.data(), methods, nullptr, d_out, out_cap, d_out_off,
                    d_out_size, nullptr, &T, pack_align ? pack_align : 16, false);
}

// tok3's per-token-type method tables, R[level][type] of tokenise_name3.c:1283-1357 (rANS build),
// one row per level group (-1, -3, -5, -7, -9); token types in enum order (tokenise_name3.c:96-112):
// TYPE ALPHA CHAR DIGITS0 DZLEN DUP DIFF DIGITS DDELTA DDELTA0 MATCH NOP END
API int b200rans_tok3_methods(int level, int token_type, unsigned int in_len, int *out) {
    static const int R[5][13][7] = {
        {{128, -1}, {129, -1}, {0, -1}, {8, -1}, {0, -1}, {8, -1}, {8, -1}, {8, -1}, {0, -1}, {128, -1}, {0, -1}, {0, -1}, {0, -1}},
        {{192, 0, -1}, {129, 1, -1}, {0, -1}, {136, 0, -1}, {0, -1}, {200, -1}, {136, -1}, {200, -1}, {0, -1}, {128, -1}, {0, -1}, {0, -1}, {0, -1}},
        {{192, 0, -1}, {1, 128, 0, 129, -1}, {0, -1}, {200, 0, -1}, {0, -1}, {200, -1}, {192, 200, -1}, {132, 201, -1}, {0, -1}, {128, -1}, {0, -1}, {0, -1}, {0, -1}},
        {{193, 0, 1, -1}, {128, 1, 128, 0, 129, -1}, {1, 0, -1}, {200, 0, -1}, {0, -1}, {201, -1}, {192, 200, -1}, {132, 201, -1}, {0, -1}, {128, -1}, {0, -1}, {0, -1}, {0, -1}},
        {{192, 0, 1, 65, 193, 132, -1}, {132, 1, 0, 129, -1}, {1, 0, 192, -1}, {201, 0, 192, 64, -1}, {0, 128, 1, -1}, {201, -1}, {192, 201, 65, -1}, {132, 201, 1, 192, 129, 193, -1}, {1, 0, 192, -1}, {192, 1, 0, -1}, {0, -1}, {0, -1}, {0, -1}},
    };
    if (!out || token_type < 0 || token_type > 12) return B200RANS_EINVAL;
    int row = (level - 1) / 2;                       // tokenise_name3.c:1275-1278
    if (row < 0) row = 0;
    if (row > 4) row = 4;
    int n = 0;
    for (int j = 0; j < 7 && R[row][token_type][j] >= 0; j++) {
        int m = R[row][token_type][j] & ~X_32;       // :1374-1375
        if ((in_len % 4) != 0 && (m & X_STRIPE)) continue;          // :1377-1378
        out[n++] = m;
    }
    return n;
}

// ============================================================ multi-GPU
API int b200rans_compress_batch_multi(int ngpu, int n, const unsigned char *const *in,
                                      const unsigned int *in_size, const int *order, const int *block_of,
                                      unsigned char *out, size_t out_cap, size_t *out_off,
                                      unsigned int *out_size) {
    if (n < 0 || ngpu <= 0 || (n && (!in || !in_size || !order || !out || !out_off || !out_size))) return B200RANS_EINVAL;
    if (block_of) for (int k = 0; k < n; k++) if (block_of[k] < 0) return B200RANS_EINVAL;
    // every device gets a private slice of the arena sized by its streams' bounds
    std::vector<std::vector<int>> idx(ngpu);
    for (int k = 0; k < n; k++) idx[(block_of ? block_of[k] : k) % ngpu].push_back(k);
    std::vector<size_t> base(ngpu + 1, 0);
    for (int g = 0; g < ngpu; g++) {
        size_t t = 0;
        for (int k : idx[g]) t += al(compress_bound(in_size[k], order[k]), 16) + 16;
        base[g + 1] = base[g] + al(t + 256, 256);
    }
    if (base[ngpu] > out_cap) return B200RANS_ESPACE;
    return run_on_workers(ngpu, 1, [&](int g, int) {
        int m = (int)idx[g].size();
        if (!m) return 0;
        std::vector<const unsigned char *> i2(m);
        std::vector<unsigned int> s2(m), z2(m);
        std::vector<int> o2(m);
        std::vector<size_t> f2(m);
        for (int j = 0; j < m; j++) { int k = idx[g][j]; i2[j] = in[k]; s2[j] = in_size[k]; o2[j] = order[k]; }
        int r = compress_batch_impl(m, i2.data(), s2.data(), o2.data(), nullptr, out + base[g],
                                    base[g + 1] - base[g], f2.data(), z2.data());
        if (r) return r;
        for (int j = 0; j < m; j++) { int k = idx[g][j]; out_off[k] = base[g] + f2[j]; out_size[k] = z2[j]; }
        return 0;
    });
}

API int b200rans_uncompress_batch_multi(int ngpu, int n, const unsigned char *const *in,
                                        const unsigned int *in_size, const int *block_of,
                                        unsigned char *const *out, unsigned int *out_size, int *status) {
    if (n < 0 || ngpu <= 0 || (n && (!in || !in_size || !out || !out_size))) return B200RANS_EINVAL;
    if (block_of) for (int k = 0; k < n; k++) if (block_of[k] < 0) return B200RANS_EINVAL;
    std::vector<std::vector<int>> idx(ngpu);
    for (int k = 0; k < n; k++) idx[(block_of ? block_of[k] : k) % ngpu].push_back(k);
    return run_on_workers(ngpu, 1, [&](int g, int) {
        int m = (int)idx[g].size();
        if (!m) return 0;
        std::vector<const unsigned char *> i2(m);
        std::vector<unsigned char *> o2(m);
        std::vector<unsigned int> s2(m), z2(m);
        std::vector<int> st2(m);
        for (int j = 0; j < m; j++) { int k = idx[g][j]; i2[j] = in[k]; s2[j] = in_size[k]; o2[j] = out[k]; z2[j] = out_size[k]; }
        int r = uncompress_batch_impl(m, i2.data(), s2.data(), o2.data(), z2.data(), st2.data());
        if (r) return r;
        for (int j = 0; j < m; j++) { int k = idx[g][j]; out_size[k] = z2[j]; if (status) status[k] = st2[j]; }
        return 0;
    });
}

// ============================================================ the method learner (host logic only)
// metrics_method / metrics_update of fqzcomp5.c:1899-1958, per section
API void b200fqz_learner_init(b200fqz_learner *L) {
    if (!L) return;
    memset(L, 0, sizeof(*L));                        // review <= 0: the first call starts a trial
    for (int s = 0; s < 3; s++) L->trial[s] = -99999;
}

static const int *sec_list(const b200fqz_block_opts *o, int s, int *n) {
    *n = s == 0 ? o->n_name_methods : s == 1 ? o->n_seq_methods : o->n_qual_methods;
    return s == 0 ? o->name_methods : s == 1 ? o->seq_methods : o->qual_methods;
}

API void b200fqz_learner_methods(b200fqz_learner *L, const b200fqz_block_opts *all, b200fqz_block_opts *out) {
    if (!L || !all || !out) return;
    *out = *all;
    for (int s = 0; s < 3; s++) {
        int n;
        const int *lst = sec_list(all, s, &n);
        if (n <= 1) { L->on_trial[s] = 0; continue; }
        if (L->review[s] <= 0) {                                     // :1902-1908
            L->review[s] = 100;                                      // METRICS_REVIEW
            L->trial[s] = 3;                                         // METRICS_TRIAL
            memset(L->usize[s], 0, sizeof(L->usize[s]));
            memset(L->csize[s], 0, sizeof(L->csize[s]));
        }
        L->on_trial[s] = 0;
        if (L->trial[s] > 0) {                                       // under evaluation: all methods
            L->on_trial[s] = 1;
            continue;
        }
        if (L->trial[s] > -99999) {                                  // trial finished: pick the best (:1916-1931)
            int best = 0;
            double best_sz = 1e30;
            for (int m = 0; m < n; m++)
                if (L->usize[s][m] && best_sz > (L->csize[s][m] + 1.0) / L->usize[s][m]) {
                    best_sz = (L->csize[s][m] + 1.0) / L->usize[s][m];
                    best = m;
                }
            L->used[s] = best;
            L->trial[s] = -99999;
        } else L->review[s]--;                                       // :1938-1940
        int *dst = s == 0 ? out->name_methods : s == 1 ? out->seq_methods : out->qual_methods;
        dst[0] = lst[L->used[s]];
        if (s == 0) out->n_name_methods = 1; else if (s == 1) out->n_seq_methods = 1; else out->n_qual_methods = 1;
    }
}

API void b200fqz_learner_update(b200fqz_learner *L, const b200fqz_block_opts *out, const b200fqz_block_report *rep) {
    if (!L || !out || !rep || rep->status) return;
    for (int s = 0; s < 3; s++) {
        if (!L->on_trial[s] || L->trial[s] <= 0) continue;           // :1951-1952
        int n;
        sec_list(out, s, &n);
        // a RANSXN1 placeholder that was skipped (reads of different lengths) leaves a hole: the report's sizes
        // are indexed by the RESOLVED list, the learner's by the caller's
        const int *lst = sec_list(out, s, &n);
        int k = 0;
        for (int m = 0; m < n; m++) {
            if (lst[m] == B200FQZ_RANSXN1 && rep->fixed_len <= 0) continue;
            L->usize[s][m] += rep->ulen[s];
            L->csize[s][m] += rep->csize[s][k++];
        }
        L->trial[s]--;                                               // compress_with_methods, :2131
    }
}

// ============================================================ C ABI: part 5 (blocks)
API size_t b200fqz_block_bound(uint32_t n) {
    // sections cannot grow beyond their bound; names may be coded by the caller (2x + 1000, fqzcomp5.c:1412)
    return (size_t)n + (size_t)n / 8 + 3 * (257 * 257 * 3 + 4096) + 65536 + (size_t)(n / 1024) * 8;
}

API int b200fqz_encode_block(const unsigned char *text, uint32_t n, const b200fqz_block_opts *opts,
                             unsigned char *block, size_t block_cap, b200fqz_block_report *rep) {
    if (!opts || !rep) return B200RANS_EINVAL;
    return encode_block_impl(text, n, opts, block, block_cap, rep);
}

API int b200fqz_decode_block(const unsigned char *block, uint32_t block_len, int plus_name, unsigned char *text,
                             size_t text_cap, b200fqz_block_report *rep) {
    if (!rep) return B200RANS_EINVAL;
    return decode_block_impl(block, block_len, plus_name, text, text_cap, rep);
}

API int b200fqz_encode_blocks_multi(int ngpu, int nblocks, const unsigned char *const *text, const uint32_t *n,
                                    const b200fqz_block_opts *opts, unsigned char *const *block,
                                    const size_t *block_cap, b200fqz_block_report *rep) {
    if (nblocks < 0 || ngpu <= 0 || !opts || (nblocks && (!text || !n || !block || !block_cap || !rep))) return B200RANS_EINVAL;
    const int W = block_workers();
    return run_on_workers(ngpu, W, [&](int g, int s) {
        for (int b = g + s * ngpu; b < nblocks; b += ngpu * W) {   // block b on device b % ngpu
            int r = encode_block_impl(text[b], n[b], opts, block[b], block_cap[b], &rep[b]);
            if (r) return r;
        }
        return 0;
    });
}

API int b200fqz_decode_blocks_multi(int ngpu, int nblocks, const unsigned char *const *block,
                                    const uint32_t *block_len, int plus_name, unsigned char *const *text,
                                    const size_t *text_cap, b200fqz_block_report *rep) {
    if (nblocks < 0 || ngpu <= 0 || (nblocks && (!block || !block_len || !text || !text_cap || !rep))) return B200RANS_EINVAL;
    const int W = block_workers();
    return run_on_workers(ngpu, W, [&](int g, int s) {
        for (int b = g + s * ngpu; b < nblocks; b += ngpu * W) {
            int r = decode_block_impl(block[b], block_len[b], plus_name, text[b], text_cap[b], &rep[b]);
            if (r) return r;
        }
        return 0;
    });
}
