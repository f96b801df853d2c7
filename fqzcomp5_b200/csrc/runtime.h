// runtime.h -- per-thread CUDA contexts, arenas and pipeline lanes of libb200rans.so, shared by
// the C-ABI files (api.cu: codec entry points; block.cu: block pipeline and multi-GPU workers).
// Mirrors the role of the per-thread scratch of htscodecs/utils.c:119-208: every host thread that
// calls the library owns a private context, so hts_tpool workers may call concurrently.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <vector>

#include "../../include/b200rans.h"
#include "kernels.h"
#include "stripe.h"

namespace b200rt {
using namespace b200;

inline bool cuda_ok(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    fprintf(stderr, "libb200rans: %s failed: %s\n", what, cudaGetErrorString(e));
    return false;
}
#define CK(call) do { if (!cuda_ok((call), #call)) return B200RANS_ECUDA; } while (0)

inline size_t al(size_t v, size_t a = 256) { return (v + a - 1) & ~(a - 1); }

struct Arena {
    uint8_t *p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    int ensure(size_t need) {
        if (need <= cap) return 0;
        size_t want = need + (need < ((size_t)2 << 30) ? need / 4 : need / 32) + (1 << 20);   // big arenas: little slack
        if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); p = nullptr; cap = 0; }
        cudaError_t e = pinned ? cudaHostAlloc((void **)&p, want, cudaHostAllocDefault)
                               : cudaMalloc((void **)&p, want);
        if (e != cudaSuccess) {
            fprintf(stderr, "libb200rans: %s of %zu bytes failed: %s\n",
                    pinned ? "cudaHostAlloc" : "cudaMalloc", want, cudaGetErrorString(e));
            p = nullptr;
            return B200RANS_ENOMEM;
        }
        cap = want;
        return 0;
    }
    void release() { if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); } p = nullptr; cap = 0; }
};

// bump sub-allocator over an arena laid out before allocation (two passes: size, place)
struct Layout {
    size_t off = 0;
    size_t take(size_t bytes, size_t a = 256) { off = al(off, a); size_t o = off; off += bytes; return o; }
};

constexpr int NSTAGE = 4;
constexpr int NPIPE = 6;                  // lanes: chunks in flight in the host-buffer API
// a pipeline chunk is ~chunk_bytes() of uncompressed data but at least chunk_min_streams() streams
// (a chunk of few streams is latency bound); see the knobs below
constexpr int CHUNK_MIN_STREAMS_SLOW = 1536;  // PACK / RLE streams take several ms each whatever their number: a
                                              // chunk should fill most of the GPU's stream slots
constexpr int CHUNK_MAX_STREAMS = 16384;
struct Stage { Arena h; cudaEvent_t ev = nullptr; bool busy = false; };

// Tuning knobs of the host-buffer pipeline (environment overrides are for measurement only).
inline int env_int(const char *name, int dflt, int lo, int hi) {
    const char *e = getenv(name);
    if (!e) return dflt;
    int v = atoi(e);
    return v < lo ? lo : v > hi ? hi : v;
}
// chunks submitted (copy in + kernels queued) ahead of the chunk whose results are being read back:
// the host blocks on that chunk's kernels, and without work queued behind it the copy engines idle
inline int pipe_depth() { static int v = env_int("B200RANS_PIPE_DEPTH", 2, 1, NPIPE - 1); return v; }
// plain streams of at least this many bytes get their counts from hist_kernel (one CTA per stream)
// instead of counting inside the coder warp
inline uint32_t hist_min_bytes(bool o1) {
    static uint32_t v0 = (uint32_t)env_int("B200RANS_HIST_MIN_O0", 4096, 0, 0x7fffffff);
    static uint32_t v1 = (uint32_t)env_int("B200RANS_HIST_MIN_O1", 4096, 0, 0x7fffffff);
    return o1 ? v1 : v0;
}
// PACK / RLE streams: transforms, counts and order-1 model by one CTA per stream (prep_kernel) in front of the
// coder warp; 0 = the coder warp does everything itself (kept for measurement)
inline bool use_prep() { static int v = env_int("B200RANS_PREP", 1, 0, 1); return v != 0; }
// order-1 streams behind PACK / RLE: staged decode (dec_staged.cu); 0 = the general kernel alone (kept for measurement)
inline bool use_dec_staged() { static int v = env_int("B200RANS_DEC_STAGED", 1, 0, 1); return v != 0; }
inline size_t chunk_bytes() { static size_t v = (size_t)env_int("B200RANS_CHUNK_MB", 48, 1, 1024) << 20; return v; }
inline int chunk_min_streams_slow(bool dec) {
    static int ve = env_int("B200RANS_CHUNK_STREAMS_SLOW", CHUNK_MIN_STREAMS_SLOW, 1, 16384);
    static int vd = env_int("B200RANS_CHUNK_STREAMS_SLOW_DEC", CHUNK_MIN_STREAMS_SLOW, 1, 16384);
    return dec ? vd : ve;
}
inline int chunk_min_streams_4lane() { static int v = env_int("B200RANS_CHUNK_STREAMS_4LANE", 4096, 1, 16384); return v; }
inline int chunk_min_streams() { static int v = env_int("B200RANS_CHUNK_STREAMS", 256, 1, 16384); return v; }

// One pipeline lane: a stream plus the arenas a chunk of work needs.  Chunks of a
// large host-buffer batch rotate over NPIPE lanes so that the H2D copy of one
// chunk, the kernels of the previous and the D2H copy of the one before overlap.
struct Lane {
    cudaStream_t st = nullptr;
    cudaStream_t aux[3] = {nullptr, nullptr, nullptr};   // side streams: coder launches of different routes run side by side
    cudaEvent_t fork = nullptr, fork2 = nullptr, join[3] = {nullptr, nullptr, nullptr};
    Arena work;                 // device: jobs, slots, scratch, pool
    Arena io;                   // device: staged inputs / outputs of the host-buffer API
    Arena crc;                  // device: CRC-32 tables and tile values (kept apart from `work`, which an
                                // encode still in flight on another stream may be using)
    Arena hio;                  // pinned: results read back
    Stage stage[NSTAGE];        // pinned: job descriptors in flight
    int next_stage = 0;

    int init() {
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        for (auto &a : aux) CK(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&fork2, cudaEventDisableTiming));
        for (auto &j : join) CK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming));
        hio.pinned = true;
        for (auto &s : stage) { s.h.pinned = true; CK(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming)); }
        return 0;
    }
    // pinned staging block for descriptors; waits until its previous use has been consumed
    int get_stage(size_t bytes, Stage **out) {
        Stage &s = stage[next_stage];
        next_stage = (next_stage + 1) % NSTAGE;
        if (s.busy) { CK(cudaEventSynchronize(s.ev)); s.busy = false; }
        int r = s.h.ensure(bytes);
        if (r) return r;
        *out = &s;
        return 0;
    }
    void destroy() {
        if (st) cudaStreamSynchronize(st);
        work.release(); io.release(); crc.release(); hio.release();
        for (auto &s : stage) { s.h.release(); if (s.ev) cudaEventDestroy(s.ev); }
        for (auto &a : aux) if (a) { cudaStreamSynchronize(a); cudaStreamDestroy(a); a = nullptr; }
        if (fork) cudaEventDestroy(fork);
        if (fork2) cudaEventDestroy(fork2);
        for (auto &j : join) if (j) cudaEventDestroy(j);
        if (st) cudaStreamDestroy(st);
        st = nullptr;
    }
};

struct Ctx {
    int dev = 0;
    bool ok = false;
    Lane lane[NPIPE];           // host-buffer API: pipeline chunks rotate over these
    Lane dlane;                 // device-resident (`_dev`) API: its own stream, scratch and staging, so that
                                // work queued on a caller's stream never shares scratch with a host-buffer call
    Arena single;               // pinned: the stream of a single rans_compress_to_4x16 call on its way out
    Arena blk, hblk;            // block pipeline (block.cu): device buffers of the block in flight / pinned results
    uint64_t launches = 0;
    bool prof = false;          // bracket the coder kernels with timing events (bench.py roofline)
    cudaEvent_t pe[4] = {nullptr, nullptr, nullptr, nullptr};   // enc start/stop, dec start/stop
    cudaEvent_t ev_blk = nullptr;   // block pipeline: "the split's small results are on the host"
    bool pe_valid[2] = {false, false};

    int init(int device) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            fprintf(stderr, "libb200rans: no usable CUDA device (%s); this library has no CPU path\n",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
            return B200RANS_ENODEV;
        }
        if (device < 0 || device >= n) return B200RANS_EINVAL;
        dev = device;
        CK(cudaSetDevice(dev));
        for (auto &l : lane) { int r = l.init(); if (r) return r; }
        { int r = dlane.init(); if (r) return r; }
        single.pinned = true;
        hblk.pinned = true;
        for (auto &e2 : pe) CK(cudaEventCreate(&e2));
        CK(cudaEventCreateWithFlags(&ev_blk, cudaEventDisableTiming));
        ok = true;
        return 0;
    }
    ~Ctx() {
        if (!ok) return;
        cudaSetDevice(dev);
        for (auto &l : lane) l.destroy();
        dlane.destroy();
        single.release(); blk.release(); hblk.release();
        for (auto &e : pe) if (e) cudaEventDestroy(e);
        if (ev_blk) cudaEventDestroy(ev_blk);
    }
};

// Method trial (compress_with_methods, fqzcomp5.c:1979-2119; tok3's compress(),
// tokenise_name3.c:1268-1417): the n calls of an enc_core batch are groups of candidate
// encodings of the same input, calls d_first[k] .. d_first[k+1]-1 belonging to input k.
// Every candidate's size goes to d_csize, the first smallest of each group is kept and
// d_out_off / d_out_size / d_best are per input.
struct Trial {
    uint32_t inputs;
    const uint32_t *d_first;    // [inputs + 1]
    uint32_t *d_csize;          // [n]
    uint32_t *d_jobidx;         // [n] scratch
    int32_t *d_best;            // [inputs]
};

// kernel launches issued by the multi-GPU worker threads (their contexts are not the caller's)
extern std::atomic<uint64_t> g_worker_launches;

// the calling thread's context (created on first use); *err receives a b200rans_status on failure
Ctx *get_ctx(int *err);
void set_thread_device(int device);      // device used by contexts this thread creates from now on

// Encode / decode of a batch whose inputs are on the device; asynchronous on st.  pack_align: the
// packed streams start on multiples of it (16 for the codec API, 1 for block sections); inslot: no
// packing at all, every stream stays in its own bound-sized slot inside d_out (see api.cu).
int enc_core(Ctx &C, Lane &Ln, cudaStream_t st, int n, const uint8_t *d_in, const uint64_t *in_off,
             const uint32_t *in_size, const int *order, const uint32_t *caps,
             uint8_t *d_out, size_t out_cap, uint64_t *d_out_off, uint32_t *d_out_size,
             uint64_t *d_total, const Trial *trial = nullptr, uint32_t pack_align = 16, bool inslot = false);
size_t enc_slots_bound(int n, const uint32_t *in_size, const int *order);   // out_cap needed by in-slot mode
int dec_core(Ctx &C, Lane &Ln, cudaStream_t st, int n, const uint8_t *d_in, const uint64_t *in_off,
             const uint32_t *in_size, const uint8_t *flags, uint8_t *d_out, const uint64_t *out_off,
             const uint32_t *out_cap, uint32_t *d_osz, int *d_status);

// host-side peek at a stream header: flag byte, stored length, header bytes
bool peek_header(const unsigned char *in, unsigned int in_size, int *flag, uint32_t *ulen, int *hdr);

// host-buffer batches (pipelined over the context's lanes)
int compress_batch_impl(int n, const unsigned char *const *in, const unsigned int *in_size, const int *order,
                        const uint32_t *caps, unsigned char *out, size_t out_cap, size_t *out_off,
                        unsigned int *out_size, const uint32_t *mfirst = nullptr, const int *methods = nullptr,
                        int *best = nullptr, unsigned int *csize = nullptr);
int uncompress_batch_impl(int n, const unsigned char *const *in, const unsigned int *in_size,
                          unsigned char *const *out, unsigned int *out_size, int *status);

}  // namespace b200rt
