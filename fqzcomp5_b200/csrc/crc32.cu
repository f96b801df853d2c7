// crc32.cu -- zlib-compatible CRC-32 of a buffer in HBM, and fqzcomp5's block framing
// (encode_block, fqzcomp5.c:2147-2280: [block size][records][crc32][sections...], the CRC
// taken over everything after the CRC field, :2266-2274) assembled on the device so that a
// finished block leaves the GPU in one copy.
//
// CRC-32 is linear over GF(2).  With raw(M) = the register after M starting from 0 (no
// conditioning) and v * x^k the register shifted through k zero bits,
//     raw(A || B) = raw(A) * x^(8|B|)  ^  raw(B),        raw(0..0 || M) = raw(M),
//     crc32(c, M) = raw(M) ^ (~c) * x^(8|M|) ^ ~0          (zlib's crc32(): c = 0 to start).
// So the buffer is cut into 512-byte pieces, one per thread, whose raw values are combined
// pairwise up a tree, each step multiplying the left value by a constant power of x:
//   * the grid is laid over the 16-byte aligned words that hold the buffer; bytes outside it
//     are masked to zero.  Zeros in front change nothing; the k zeros behind are undone at the
//     end by multiplying with x^(-8k) (x is invertible modulo the CRC polynomial, its order
//     divides 2^32 - 1);
//   * a word is absorbed nibble by nibble from 8 x 16-entry tables replicated once per
//     shared-memory bank (16 KiB), so the look-ups of a warp never conflict.
// Algorithmic bytes: n read.  One pass over the data plus a one-CTA finishing kernel.
#include "crc32.h"
#include "common.cuh"

namespace b200 {
namespace {

constexpr uint32_t POLY = 0xEDB88320u;       // reflected CRC-32 (ISO-HDLC) polynomial, as zlib's
constexpr uint32_t PIECE = 512;              // bytes per thread
constexpr uint32_t CTPB = 256;               // threads per CTA
constexpr uint32_t CTILE = PIECE * CTPB;     // 128 KiB per CTA

// a * b modulo the polynomial, reflected bit order (x^0 is bit 31)
__host__ __device__ inline uint32_t mulmod(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (int i = 0; i < 32; i++) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = (b >> 1) ^ ((b & 1) ? POLY : 0u);
    }
    return p;
}
// x^(e) modulo the polynomial, e in bits, by square and multiply
__host__ __device__ inline uint32_t xpow(uint64_t e) {
    uint32_t r = 0x80000000u, s = 0x40000000u;          // 1 and x
    while (e) {
        if (e & 1) r = mulmod(r, s);
        s = mulmod(s, s);
        e >>= 1;
    }
    return r;
}

// register after absorbing the 32-bit value v from state 0: v * x^32
__host__ __device__ inline uint32_t step32(uint32_t v) {
    for (int i = 0; i < 32; i++) v = (v >> 1) ^ ((v & 1) ? POLY : 0u);
    return v;
}

struct CrcWork { uint32_t nib[128]; };                  // nib[k*16+v] = step32(v << 4k)
struct CrcConst {                                       // powers of x computed on the host per call
    uint32_t kp[6];                                     // x^(8 * PIECE * 2^j), j = 0..5
    uint32_t kt, kper;                                  // x^(8 * CTILE), x^(8 * CTILE * per)
    uint32_t xinv_pad, xn;                              // x^(-8 * pad), x^(8 * n)
    uint32_t per;
};

__global__ void crc_tables_kernel(CrcWork *W) {
    uint32_t i = threadIdx.x;
    if (i < 128) W->nib[i] = step32((i & 15u) << (4 * (i >> 4)));
}

__global__ void __launch_bounds__(CTPB)
crc_tiles_kernel(const uint8_t *__restrict__ base, uint32_t lead, uint64_t end, const CrcWork *__restrict__ W,
                 uint32_t *__restrict__ tile_crc, CrcConst C) {
    // base is 16-byte aligned; the buffer is base[lead .. end)
    __shared__ uint32_t T[128 * 32];                    // T[(k*16+v)*32 + bank]
    __shared__ uint32_t wc[CTPB / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 128 * 32; i += CTPB) T[i] = W->nib[i >> 5];
    __syncthreads();
    const uint32_t ts = (uint32_t)__cvta_generic_to_shared(T) + 4 * lane;
    const uint64_t p0 = (uint64_t)blockIdx.x * CTILE + (uint64_t)threadIdx.x * PIECE;
    uint32_t c = 0;
    if (p0 < end && p0 + PIECE > lead) {
        const uint4 *src = (const uint4 *)(base + p0);
#pragma unroll 2
        for (uint32_t q = 0; q < PIECE / 16; q++) {
            const uint64_t p = p0 + 16 * q;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (p < end && p + 16 > lead) {
                v = __ldg(src + q);
                if (p < lead || p + 16 > end) {         // mask the bytes outside the buffer
                    uint32_t w[4] = {v.x, v.y, v.z, v.w};
                    for (int j = 0; j < 16; j++) {
                        uint64_t a = p + j;
                        if (a < lead || a >= end) w[j >> 2] &= ~(0xffu << (8 * (j & 3)));
                    }
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t x = c ^ w4[j], r = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    uint32_t t;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(ts + ((((x >> (4 * k)) & 15u) + 16u * k) << 7)));
                    r ^= t;
                }
                c = r;
            }
        }
    }
    // combine up the tree: left * x^(8 * bytes of the right part) ^ right
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const int s = 1 << j;
        uint32_t right = __shfl_down_sync(FULL, c, s);
        if ((lane & (2 * s - 1)) == 0) c = mulmod(c, C.kp[j]) ^ right;
    }
    if (lane == 0) wc[wid] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t r = wc[0];
        for (uint32_t w = 1; w < CTPB / 32; w++) r = mulmod(r, C.kp[5]) ^ wc[w];
        tile_crc[blockIdx.x] = r;
    }
}

// One CTA: fold the tile values, undo the padding, apply zlib's conditioning.
__global__ void __launch_bounds__(1024)
crc_finish_kernel(const uint32_t *__restrict__ tile_crc, uint32_t ntiles, uint32_t crc_in, uint32_t *__restrict__ out,
                  uint8_t *patch, uint32_t patch_size_val, CrcConst C) {
    __shared__ uint32_t sc[1024];
    // thread t folds tiles [t*per, (t+1)*per) left to right; tiles past ntiles are virtual zeros
    // (part of the padding undone below)
    uint32_t c = 0;
    for (uint32_t i = 0; i < C.per; i++) {
        uint32_t t = threadIdx.x * C.per + i;
        c = mulmod(c, C.kt) ^ (t < ntiles ? tile_crc[t] : 0u);
    }
    sc[threadIdx.x] = c;
    __syncthreads();
    uint32_t K = C.kper;
    for (uint32_t s = 1; s < 1024; s <<= 1) {
        if ((threadIdx.x & (2 * s - 1)) == 0) sc[threadIdx.x] = mulmod(sc[threadIdx.x], K) ^ sc[threadIdx.x + s];
        K = mulmod(K, K);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint32_t raw = mulmod(sc[0], C.xinv_pad);
        const uint32_t crc = raw ^ mulmod(~crc_in, C.xn) ^ 0xffffffffu;
        if (out) *out = crc;
        if (patch) {                                               // block framing: size and CRC fields
            patch[0] = (uint8_t)patch_size_val; patch[1] = (uint8_t)(patch_size_val >> 8);
            patch[2] = (uint8_t)(patch_size_val >> 16); patch[3] = (uint8_t)(patch_size_val >> 24);
            patch[8] = (uint8_t)crc; patch[9] = (uint8_t)(crc >> 8);
            patch[10] = (uint8_t)(crc >> 16); patch[11] = (uint8_t)(crc >> 24);
        }
    }
}

}  // namespace

size_t crc32_scratch_bytes(uint64_t n) {
    return 1024 + ((n + 32) / CTILE + 2) * 4;
}

cudaError_t crc32_launch(const uint8_t *d_buf, uint64_t n, uint32_t crc_in, uint32_t *d_out, uint8_t *d_patch,
                         uint32_t patch_size_val, uint8_t *d_scratch, cudaStream_t st, int *launches) {
    CrcWork *W = (CrcWork *)d_scratch;
    uint32_t *tile_crc = (uint32_t *)(d_scratch + 1024);
    const uint32_t lead = (uint32_t)((uintptr_t)d_buf & 15);
    const uint8_t *base = d_buf - lead;
    const uint64_t end = (uint64_t)lead + n;
    const uint32_t ntiles = (uint32_t)((end + CTILE - 1) / CTILE);
    CrcConst C;
    C.kp[0] = xpow(8ull * PIECE);
    for (int j = 1; j < 6; j++) C.kp[j] = mulmod(C.kp[j - 1], C.kp[j - 1]);
    C.per = (ntiles + 1023) / 1024;
    C.kt = xpow(8ull * CTILE);
    C.kper = xpow(8ull * CTILE * C.per);
    const uint64_t V = (uint64_t)CTILE * C.per * 1024;             // virtual bytes covered by the tree
    const uint64_t pad = V - end;                                  // zero bytes behind the buffer
    const uint64_t ord = 0xffffffffull;                            // the order of x divides 2^32 - 1
    C.xinv_pad = xpow((ord - (8 * pad) % ord) % ord);
    C.xn = xpow(8 * n);
    crc_tables_kernel<<<1, 128, 0, st>>>(W);
    if (ntiles) crc_tiles_kernel<<<ntiles, CTPB, 0, st>>>(base, lead, end, W, tile_crc, C);
    crc_finish_kernel<<<1, 1024, 0, st>>>(tile_crc, ntiles, crc_in, d_out, d_patch, patch_size_val, C);
    if (launches) *launches += ntiles ? 3 : 2;
    return cudaGetLastError();
}

}  // namespace b200
