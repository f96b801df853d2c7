// common.cuh -- shared definitions for the B200 rANS Nx16 kernels.
//
// Execution model: one warp owns one rANS stream.  The N (4 or 32)
// interleaved rANS states of the reference (rANS_static32x16pr.c:130-135) map to
// lanes 0..N-1 of that warp; everything around the serial state chain (table
// parsing, histograms, normalisation, copies) is done by all 32 lanes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t RANS_L = 1u << 15;            // rANS_word.h:64
constexpr int X_PACK = 0x80, X_RLE = 0x40, X_CAT = 0x20, X_NOSZ = 0x10,
              X_STRIPE = 0x08, X_32 = 0x04;      // rANS_static16_int.h:48-53
constexpr int ORDER_STRIPE_NO0 = 1 << 16, ORDER_SIMD_AUTO = 1 << 17;

// Per-stream job records.  Host fills the inputs; kernels fill the results.
struct EncJob {
    const uint8_t *in;      // uncompressed input (device)
    uint8_t *slot;          // private scratch slot, 16-byte aligned
    uint32_t in_size;
    int32_t  order;         // the caller's order argument, unmodified
    uint32_t slot_cap;      // bytes in slot (>= rans_compress_bound_4x16, multiple of 16)
    uint32_t cap;           // *out_size on entry to the emulated call
    // results: the stream is slot[0,head_len) followed by tail[0,tail_len)
    const uint8_t *tail;
    uint32_t head_len;
    uint32_t tail_len;
    uint32_t status;        // 0 ok; !=0: the reference call would have returned NULL
    uint32_t need_cap;      // smallest *out_size for which the call succeeds (STRIPE selection)
    // optional pre-transformed inputs (PACK / RLE pre-pass), else null
    uint8_t *work;          // scratch for transforms: 4.25 * in_size + 8192 bytes
    uint32_t item;          // index into the caller's arrays; 0xffffffff for STRIPE sub-streams
    uint32_t stripe_n;      // >0: STRIPE parent record; its tail is the chosen sub-streams,
                            //     whose job indices sit at (uint32_t*)(slot + slot_cap) (the host
                            //     allocates STRIPE_LIST_BYTES behind the slot of a parent)
    uint32_t route;         // ROUTE_*: which launch codes this stream
    uint32_t stripe_nmeth;  // STRIPE parent: candidate methods per stripe; sub-stream (i, j) is the
                            //     job at parent + 1 + i * stripe_nmeth + j
    uint32_t *model;        // counts precomputed by hist_kernel, or null (the coder counts itself):
                            //   [256] order-0 counts, [MODEL_HDR_WORDS..] order-1 pair counts in rank space
    uint8_t *prep;          // PACK / RLE streams: the area prep_kernel fills (prep.cuh), or null (the coder warp
                            //   does the transforms and the model itself)
};
enum : uint32_t {
    ROUTE_O0 = 0,           // order-0 kernel
    ROUTE_O1 = 1,           // order-1 kernel
    ROUTE_NONE = 2,         // not coded (STRIPE parent, assembled by stripe_select)
    ROUTE_O1_WIDE = 3,      // order-1 kernel with more shared memory per stream (PACK / RLE in front)
    ROUTE_O1_PREP = 4,      // order-1 stream prepared by prep_kernel: the lean chains-only kernel
};
constexpr uint32_t MODEL_HDR_WORDS = 260;   // 256 counts, nsym, 3 pad
constexpr uint32_t STRIPE_LIST_BYTES = 1024;  // 255 job indices behind a STRIPE parent's slot

struct DecJob {
    const uint8_t *in;      // compressed stream (device)
    uint8_t *out;           // destination (device)
    uint8_t *tmp;           // scratch: 2*out_cap + 1024 bytes when PACK/RLE may be present, else null
    uint32_t in_size;
    uint32_t out_cap;       // capacity; exact length for NOSZ streams
    uint32_t out_size;      // result
    int32_t  status;        // result: 0 ok
    uint32_t route;         // 0: order-0 kernel, 1: order-1 (general) kernel, 2: staged decode (dec_staged.cuh)
    uint32_t pad_;
    uint8_t *prep;          // route 2: the stream's DecPrep record
};

enum : int32_t {
    ST_OK = 0,
    ST_FAIL = 1,            // the reference would return NULL
    ST_UNSUPPORTED = 2,
};

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// inclusive scan across the warp
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// ---- varints: 7 bits per byte, most significant first (varint.h:205-299) ----
__device__ __forceinline__ int var_put_u32(uint8_t *p, uint32_t v) {
    int n = 1;
    while (n < 5 && (v >> (7 * n))) n++;
    for (int k = n - 1; k >= 0; k--) *p++ = (uint8_t)(((v >> (7 * k)) & 0x7f) | (k ? 0x80 : 0));
    return n;
}
__device__ __forceinline__ int var_size_u32(uint32_t v) {
    int n = 1;
    while (n < 5 && (v >> (7 * n))) n++;
    return n;
}
// Bounded read; at the end of the buffer yields 0 and consumes nothing.
__device__ __forceinline__ int var_get_u32(const uint8_t *p, const uint8_t *end, uint32_t *v) {
    const uint8_t *s = p;
    uint32_t x = 0;
    int cnt = 0;
    uint8_t c = 0x80;
    while ((c & 0x80) && p < end && cnt < 6) {
        c = *p++;
        x = (x << 7) | (c & 0x7f);
        cnt++;
    }
    *v = x;
    return (int)(p - s);
}

// rANS_static4x16pr.c:93-106.  The reference evaluates this in double, left to
// right, with separately rounded multiply and adds; the library is compiled with
// -fmad=false so the device does the same.
__host__ __device__ inline uint32_t compress_bound(uint32_t size, int order) {
    int N = (order >> 8) & 0xff;
    if (!N) N = 4;
    order &= 0xff;
    double t = 1.05 * (double)size;
    if (order == 0) { t = t + (257 * 3); t = t + 4; }
    else { t = t + (257 * 257 * 3); t = t + 4; t = t + (257 * 3); t = t + 4; }
    t = t + ((order & X_PACK) ? 1 : 0);
    t = t + ((order & X_RLE) ? 1 + 257 * 3 + 4 : 0);
    t = t + 20;
    t = t + ((order & X_32) ? (32 - 4) * 4 : 0);
    t = t + ((order & X_STRIPE) ? 7 + 5 * N : 0);
    uint32_t sz = (uint32_t)t;
    return sz + (sz & 1) + 2;
}

// ---- warp-cooperative byte copy, any alignment, regions may not overlap ----
__device__ inline void warp_copy(uint8_t *dst, const uint8_t *src, uint32_t n, int lane) {
    if (n == 0) return;
    // head: bring dst to 16-byte alignment
    uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (head > n) head = n;
    if ((uint32_t)lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    uint32_t nv = n >> 4;
    if ((((uintptr_t)src) & 15) == 0) {
        const uint4 *s4 = (const uint4 *)src;
        uint4 *d4 = (uint4 *)dst;
        for (uint32_t i = lane; i < nv; i += 32) d4[i] = s4[i];
    } else {
        // assemble aligned 16-byte stores from aligned 4-byte loads with a byte funnel shift
        const uint32_t sh = ((uintptr_t)src & 3) * 8;
        const uint32_t *sw = (const uint32_t *)((uintptr_t)src & ~(uintptr_t)3);
        uint4 *d4 = (uint4 *)dst;
        if (sh == 0) {
            for (uint32_t i = lane; i < nv; i += 32) {
                const uint32_t *q = sw + 4 * i;
                d4[i] = make_uint4(q[0], q[1], q[2], q[3]);
            }
        } else {
            for (uint32_t i = lane; i < nv; i += 32) {
                const uint32_t *q = sw + 4 * i;
                uint32_t a = q[0], b = q[1], c = q[2], d = q[3], e = q[4];
                d4[i] = make_uint4(__funnelshift_r(a, b, sh), __funnelshift_r(b, c, sh),
                                   __funnelshift_r(c, d, sh), __funnelshift_r(d, e, sh));
            }
        }
    }
    uint32_t done = nv << 4;
    for (uint32_t i = done + lane; i < n; i += 32) dst[i] = src[i];
}

// Move n bytes UP to dst > src where the two ranges may overlap (memmove semantics): blocks of
// 128 bytes from the highest down; a block is read completely before it is written, and writing
// it can only touch source bytes of blocks already moved.
__device__ inline void warp_move_up(uint8_t *dst, const uint8_t *src, uint32_t n, int lane) {
    if (n == 0 || dst == src) return;
    if (dst + 0 >= src + n) { warp_copy(dst, src, n, lane); return; }      // disjoint
    uint32_t done = 0;
    while (done < n) {
        const uint32_t blk = n - done < 128 ? n - done : 128;
        const uint32_t base = n - done - blk;
        uint32_t v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const uint32_t i = lane + 32 * q; v[q] = i < blk ? src[base + i] : 0; }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; q++) { const uint32_t i = lane + 32 * q; if (i < blk) dst[base + i] = (uint8_t)v[q]; }
        __syncwarp();
        done += blk;
    }
}

// cp.async 16 bytes, bytes beyond src_bytes are zero-filled and not read
__device__ __forceinline__ void cp_async16_zfill(void *smem, const void *gmem, uint32_t src_bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

}  // namespace b200
