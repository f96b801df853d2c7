// rans_encode.cuh -- device side of rans_compress_to_4x16: histograms, the
// reference's exact frequency normalisation and table serialisation, and the
// backward N-lane rANS encode with warp-ballot compaction of the 16-bit words.
// One warp per stream.
//
// Reference behaviour restated here (never its code):
//   o0: rANS_static4x16pr.c:112-232, rANS_static32x16pr.c:67-254
//   o1: rANS_static4x16pr.c:422-518, rANS_static32x16pr.c:414-525
//   model: rANS_static16_int.h:97-146 (normalise_freq), :165-189, :240-252,
//          :278-306, :312-421 (encode_freq1); rANS_static4x16pr.c:357-420 (shift)
//   symbol: rANS_word.h:201-272 (RansEncSymbolInit), :287-336 (RansEncPutSymbol)
#pragma once
#include "common.cuh"
#include "rans_decode.cuh"   // Pool

namespace b200 {

// ------------------------------------------------------------------------
// Order-0 histogram of a byte range by one warp into shared memory.
// 16-byte loads; equal neighbours inside a lane's 16 bytes are merged into one
// shared-memory atomic (quality strings are sticky).
// ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ldg_u8(const uint8_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(__cvta_generic_to_global(p)));
    return v;
}
__device__ __forceinline__ uint4 ldg_u128(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(__cvta_generic_to_global(p)));
    return v;
}
// 16 bytes from any address: aligned 32-bit words funnel-shifted into place.  Reads only
// words that hold at least one of the 16 bytes.  Plain (coherent) loads: the bytes may have
// been produced by this kernel (PACK / RLE output).
__device__ __forceinline__ uint4 ld16_any(const uint8_t *p) {
    const uint32_t a = (uint32_t)((uintptr_t)p & 3);
    const uint32_t *w = (const uint32_t *)(p - a);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = a ? w[4] : 0u;
    const uint32_t sh = a * 8;
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                      __funnelshift_r(w3, w4, sh));
}
// 16 bytes into shared-memory bins; equal neighbours are merged before an atomic is
// issued (quality strings are sticky).  (A word-level "four equal bytes" shortcut was
// measured slower: the two paths diverge within the warp.)
__device__ __forceinline__ void hist16(uint4 q, uint32_t *F) {
    const uint32_t F_s = (uint32_t)__cvta_generic_to_shared(F);
    uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t prev = w[0] & 0xff, cnt = 0;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            uint32_t c = (w[a] >> (8 * b)) & 0xff;
            // branch-free: a predicated shared-memory reduction closes the run when the byte changes
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, %1;\n\t@p red.shared.add.u32 [%2], %3;\n\t}"
                         ::"r"(c), "r"(prev), "r"(F_s + prev * 4), "r"(cnt) : "memory");
            cnt = (c == prev) ? cnt + 1 : 1;
            prev = c;
        }
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(F_s + prev * 4), "r"(cnt) : "memory");
}

__device__ inline void warp_hist8(const uint8_t *in, uint32_t n, uint32_t *F, int lane) {
    for (int j = lane; j < 256; j += 32) F[j] = 0;
    __syncwarp();
    uint32_t head = (uint32_t)((16 - ((uintptr_t)in & 15)) & 15);
    if (head > n) head = n;
    if ((uint32_t)lane < head) atomicAdd(&F[in[lane]], 1u);
    const uint8_t *p = in + head;
    uint32_t rest = n - head, nv = rest >> 4;
    const uint4 *v = (const uint4 *)p;
    uint32_t i = lane;
    for (; i + 96 < nv; i += 128) {                  // four 16-byte loads in flight per lane
        uint4 q0 = ldg_u128(v + i), q1 = ldg_u128(v + i + 32), q2 = ldg_u128(v + i + 64), q3 = ldg_u128(v + i + 96);
        hist16(q0, F); hist16(q1, F); hist16(q2, F); hist16(q3, F);
    }
    for (; i < nv; i += 32) hist16(ldg_u128(v + i), F);
    for (uint32_t i = (nv << 4) + lane; i < rest; i += 32) atomicAdd(&F[p[i]], 1u);
    __syncwarp();
}

__device__ __forceinline__ uint32_t round2(uint32_t v) {      // rANS_static16_int.h:86-95
    v--;
    v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    return v + 1;
}

// ------------------------------------------------------------------------
// normalise_freq (rANS_static16_int.h:97-146), warp version over 256 entries
// in shared memory (lane owns entries 8*lane..8*lane+7).
// 31-bit fixed-point scale; zero stays zero, non-zero stays >= 1; the FIRST most
// frequent symbol absorbs the error; one rescale from the scaled counts if that
// would more than halve it; greedy shave as the last resort.
// ------------------------------------------------------------------------
__device__ inline int normalise_freq_warp(uint32_t *F, uint32_t size_in, uint32_t tot, int lane) {
    if (!size_in) return 0;
    int size = (int)size_in;
    int big = 0;
    for (int pass = 0; pass < 2; pass++) {
        uint64_t tr = ((uint64_t)tot << 31) / (uint32_t)size + (uint32_t)((1 << 30) / size);
        uint32_t top = 0, sum = 0;
        int arg = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            int j = lane * 8 + t;
            uint32_t f = F[j];
            if (!f) continue;
            if (top < f) { top = f; arg = j; }
            f = (uint32_t)((f * tr) >> 31);
            if (!f) f = 1;
            F[j] = f;
            sum += f;
        }
        sum = warp_sum(sum);
#pragma unroll
        for (int o = 16; o; o >>= 1) {          // max, ties -> lowest index
            uint32_t t2 = __shfl_xor_sync(FULL, top, o);
            int a2 = __shfl_xor_sync(FULL, arg, o);
            if (t2 > top || (t2 == top && a2 < arg)) { top = t2; arg = a2; }
        }
        big = top ? arg : 0;
        __syncwarp();
        int adjust = (int)tot - (int)sum;
        uint32_t fb = F[big];
        if (adjust >= 0) { if (lane == 0) F[big] = fb + adjust; break; }
        if (fb > (uint32_t)-adjust && (pass == 1 || fb / 2 >= (uint32_t)-adjust)) {
            if (lane == 0) F[big] = fb + adjust;
            break;
        }
        if (pass == 0) { size = (int)sum; continue; }
        if (lane == 0) {
            adjust += fb - 1;
            F[big] = 1;
            for (int j = 0; adjust && j < 256; j++) {
                if (F[j] < 2) continue;
                int d = (F[j] > (uint32_t)-adjust) ? adjust : 1 - (int)F[j];
                F[j] += d;
                adjust -= d;
            }
        }
    }
    __syncwarp();
    return F[big] > 0 ? 0 : -1;
}

// Same algorithm, one thread, over the n entries of one order-1 row.
__device__ inline int normalise_freq_row(uint32_t *F, uint32_t n, uint32_t size_in, uint32_t tot) {
    if (!size_in) return 0;
    int size = (int)size_in;
    uint32_t big = 0;
    for (int pass = 0; pass < 2; pass++) {
        uint64_t tr = ((uint64_t)tot << 31) / (uint32_t)size + (uint32_t)((1 << 30) / size);
        uint32_t top = 0, sum = 0;
        big = 0;
        for (uint32_t j = 0; j < n; j++) {
            uint32_t f = F[j];
            if (!f) continue;
            if (top < f) { top = f; big = j; }
            f = (uint32_t)((f * tr) >> 31);
            if (!f) f = 1;
            F[j] = f;
            sum += f;
        }
        int adjust = (int)tot - (int)sum;
        if (adjust >= 0) { F[big] += adjust; break; }
        if (F[big] > (uint32_t)-adjust && (pass == 1 || F[big] / 2 >= (uint32_t)-adjust)) {
            F[big] += adjust;
            break;
        }
        if (pass == 0) { size = (int)sum; continue; }
        adjust += F[big] - 1;
        F[big] = 1;
        for (uint32_t j = 0; adjust && j < n; j++) {
            if (F[j] < 2) continue;
            int d = (F[j] > (uint32_t)-adjust) ? adjust : 1 - (int)F[j];
            F[j] += d;
            adjust -= d;
        }
    }
    return F[big] > 0 ? 0 : -1;
}

// Alphabet list (rANS_static16_int.h:165-189), single thread.
__device__ inline int put_alphabet(uint8_t *cp, const uint32_t *F) {
    uint8_t *op = cp;
    int j = 0;
    while (j < 256) {
        if (!F[j]) { j++; continue; }
        *cp++ = (uint8_t)j;
        if (j && F[j - 1]) {
            int k = j + 1;
            while (k < 256 && F[k]) k++;
            *cp++ = (uint8_t)(k - (j + 1));
            j = k;
        } else j++;
    }
    *cp++ = 0;
    return (int)(cp - op);
}

// Encoder symbol (rANS_word.h:171-179,201-272) packed into 16 bytes:
//   x = x_max + 1, y = rcp_freq, z = bias, w = cmpl_freq << 16 | (rcp_shift - 32)
// (the shift in the low bits: a wrapping funnel shift takes its count from the low five bits of a
// register, so a step needs no instruction to extract it)
__device__ __forceinline__ uint4 enc_sym_init(uint32_t start, uint32_t freq, uint32_t bits) {
    uint4 s;
    s.x = ((RANS_L >> bits) << 16) * freq;           // a state renormalises when it is >= this (at most 2^31)
    uint32_t cmpl = ((1u << bits) - freq) & 0xffff;
    if (freq < 2) {
        s.y = ~0u;
        s.z = start + (1u << bits) - 1;
        s.w = cmpl << 16;
    } else {
        uint32_t sh = 32 - __clz(freq - 1);                 // smallest sh with freq <= 1<<sh
        s.y = (uint32_t)(((1ull << (sh + 31)) + freq - 1) / freq);
        s.z = start;
        s.w = (cmpl << 16) | (sh - 1);
    }
    return s;
}

// Compact 4-byte form for the order-1 tables (nsym^2 entries per stream):
//   bias | freq << 13 | (rcp_shift - 32) << 26.
// x_max and cmpl_freq follow from freq and the stream's precision; the 32-bit reciprocal
// depends on freq alone and comes from a 4097-entry table in global memory that every
// stream shares (L1-resident), so shared memory holds four times as many pairs as with full
// 16-byte symbols and the order-1 kernels keep twice the warps resident.
__device__ uint32_t g_rcp_freq[4097];
__device__ __forceinline__ uint32_t enc_sym_pack(uint4 s, uint32_t freq) {
    return s.z | (freq << 13) | ((s.w & 31) << 26);
}
__device__ __forceinline__ uint32_t rcp_of_freq(uint32_t f) {
    uint32_t v;
    asm("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(__cvta_generic_to_global(g_rcp_freq + f)));
    return v;
}
// What a step needs, as separate values: a symbol fetched from the 16-byte table pays one shift for cmpl, one
// unpacked from the 4-byte form is used as it is unpacked (packing it into the uint4 layout and taking it apart
// again cost three instructions per order-1 step).  shw: the shift is its low five bits.
struct EncSym {
    uint32_t xlim, rcp, bias, cmpl, shw;
    __device__ __forceinline__ EncSym() {}
    __device__ __forceinline__ EncSym(uint4 s) : xlim(s.x), rcp(s.y), bias(s.z), cmpl(s.w >> 16), shw(s.w) {}
};
__device__ __forceinline__ EncSym enc_sym_unpack(uint32_t c, uint32_t bits) {
    const uint32_t f = (c >> 13) & 0x1fff;
    EncSym s;
    s.xlim = f << (31 - bits);
    s.rcp = rcp_of_freq(f);
    s.bias = c & 0x1fff;
    s.cmpl = (1u << bits) - f;
    s.shw = c >> 26;
    return s;
}

// ------------------------------------------------------------------------
// Output staging.  The encoder's 16-bit words (and the final states) are written
// DOWNWARD; one step emits up to 64 bytes in lane order.  They are collected in a
// 1 KiB shared-memory ring indexed by the low bits of the offset within the
// (256-byte aligned) slot and leave for global memory as aligned 16-byte stores,
// 512 bytes at a time, instead of millions of scattered 2-byte stores.
// ------------------------------------------------------------------------
constexpr uint32_t ORING = 1024;
// AL: the ring starts at a multiple of its size in the shared window, so that "base + (offset mod size)" is one
// LOP3 (and, or) instead of a mask and an add.
template <bool AL>
struct OutRingT {
    uint8_t *slot;       // slot base (global, 256-byte aligned)
    uint32_t ring_s;     // shared-space address of the ring (16-byte aligned)
    uint32_t off;        // next byte to write is off-1 (downward), offset from slot
    uint32_t hi;         // bytes [off, hi) are still in the ring; hi is a multiple of 16

    __device__ __forceinline__ uint32_t at(uint32_t o) const {
        return AL ? (ring_s | (o & (ORING - 1))) : (ring_s + (o & (ORING - 1)));
    }
    // lo: any address at or below everything that will be written; out_end: even address
    // where writing starts (downward).  Up to 15 bytes above out_end may be overwritten
    // when out_end is not 16-byte aligned (callers leave that slack).
    __device__ __forceinline__ void init(uint8_t *lo, uint8_t *out_end, uint8_t *ring) {
        slot = (uint8_t *)((uintptr_t)lo & ~(uintptr_t)255);
        ring_s = (uint32_t)__cvta_generic_to_shared(ring);
        off = (uint32_t)(out_end - slot);
        hi = (off + 15) & ~15u;
    }
    // flush the highest 512-byte block(s) while more than 512 bytes are pending
    __device__ __forceinline__ void maybe_flush(int lane) {
        while (hi - off > 512) {
            uint32_t c = (hi - 1) & ~511u;
            __syncwarp();
            uint32_t o = c + 16 * lane;
            if (o < hi) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(at(o)));
                asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(__cvta_generic_to_global(slot + o)),
                             "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
            __syncwarp();
            hi = c;
        }
    }
    // everything that is left: [off, hi)
    __device__ __forceinline__ void final_flush(int lane) {
        __syncwarp();
        uint32_t a = (off + 15) & ~15u;     // first 16-byte aligned offset
        if (a > hi) a = hi;
        for (uint32_t o = off + 2 * lane; o < a; o += 64) {          // leading 2-byte pieces (off is even)
            uint32_t v;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(at(o)));
            *(uint16_t *)(slot + o) = (uint16_t)v;
        }
        for (uint32_t o = a + 16 * lane; o < hi; o += 512) {
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(at(o)));
            asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(__cvta_generic_to_global(slot + o)),
                         "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
        hi = off;
        __syncwarp();
    }
};
using OutRing = OutRingT<false>;

// One encode step for the warp (rANS_word.h:287-336 + the lane order of
// rANS_static32x16pr.c:187-231): lanes whose state exceeds x_max emit their low
// 16 bits; lane 31's word lands at the highest address.  At most 4 steps may
// pass between two maybe_flush() calls.
// The renormalisation is written out in PTX so that one predicate serves the ballot, the store and the
// shift (compiled from C++ the comparison was made twice and the shifted state went through a second
// register), and the reciprocal's shift is taken straight from the symbol's last word.
// ALL: every lane is on (the main loops of the 32-lane coders).
#define B200_ENC_RENORM(PRED, ADDR)                                                              \
    "{\n\t.reg .pred p, q;\n\t.reg .b32 m, t, k, a;\n\t" PRED                                   \
    "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"                                                \
    "shr.b32 t, m, %3;\n\t"                                                                     \
    "popc.b32 k, t;\n\t"                                                                        \
    "popc.b32 t, m;\n\t"                                                                        \
    "sub.u32 a, %1, k;\n\t"                                                                     \
    "sub.u32 a, a, k;\n\t"                                                                      \
    "sub.u32 %1, %1, t;\n\t"                                                                    \
    "sub.u32 %1, %1, t;\n\t"                                                                    \
    "and.b32 a, a, 1023;\n\t" ADDR                                                              \
    "@p st.shared.u16 [a], %0;\n\t"                                                             \
    "@p shr.u32 %0, %0, 16;\n\t}"
template <bool ALL = false, bool AL>
__device__ __forceinline__ uint32_t enc_step(uint32_t R, bool on, EncSym e, OutRingT<AL> &w, int lane) {
    static_assert(ORING == 1024, "the mask in B200_ENC_RENORM");
    uint32_t off = w.off;
    if (ALL) {
        if (AL) asm volatile(B200_ENC_RENORM("setp.ge.u32 p, %0, %2;\n\t", "or.b32 a, a, %4;\n\t")
                             : "+r"(R), "+r"(off) : "r"(e.xlim), "r"(lane), "r"(w.ring_s) : "memory");
        else asm volatile(B200_ENC_RENORM("setp.ge.u32 p, %0, %2;\n\t", "add.u32 a, a, %4;\n\t")
                          : "+r"(R), "+r"(off) : "r"(e.xlim), "r"(lane), "r"(w.ring_s) : "memory");
    } else {
        const uint32_t on32 = on;
        if (AL) asm volatile(B200_ENC_RENORM("setp.ne.u32 q, %5, 0;\n\tsetp.ge.u32.and p, %0, %2, q;\n\t", "or.b32 a, a, %4;\n\t")
                             : "+r"(R), "+r"(off) : "r"(e.xlim), "r"(lane), "r"(w.ring_s), "r"(on32) : "memory");
        else asm volatile(B200_ENC_RENORM("setp.ne.u32 q, %5, 0;\n\tsetp.ge.u32.and p, %0, %2, q;\n\t", "add.u32 a, a, %4;\n\t")
                          : "+r"(R), "+r"(off) : "r"(e.xlim), "r"(lane), "r"(w.ring_s), "r"(on32) : "memory");
    }
    w.off = off;
    if (ALL || on) {
        uint32_t q;
        asm("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(q) : "r"(__umulhi(R, e.rcp)), "r"(0), "r"(e.shw));
        R = R + e.bias + q * e.cmpl;
    }
    return R;
}

// final states, lane 0 lowest (rANS_word.h:105-117); then everything leaves the ring
template <bool AL>
__device__ __forceinline__ void enc_flush(uint32_t R, bool act, int N, OutRingT<AL> &w, int lane) {
    w.maybe_flush(lane);
    w.off -= 4 * N;
    if (act) {
        uint32_t a = w.off + 4 * lane;       // off is 2-byte aligned only
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(w.at(a)), "r"(R) : "memory");
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(w.at(a + 2)), "r"(R >> 16) : "memory");
    }
    w.final_flush(lane);
}

// ======================================================================== o0
struct __align__(16) EncO0Smem {
    uint8_t  ring[ORING];   // output staging (OutRing)
    uint4    sym[256];      // encoder symbols (enc_sym_init)
    uint32_t F[256];
};

// Writes the frequency table at `out` (forwards) and the payload below
// `out_end` (backwards).  Returns 0 ok; *tab_len, *ptr_out give the two pieces.
// Not inlined: it is called from three places (payload, RLE meta-data, self-compressed order-1
// tables), and as a function of its own its hot loop gets a register allocation that does not
// depend on what surrounds the call.
// AL: S.ring sits at a multiple of ORING in the shared window (the order-0 kernel's streams)
template <int N, bool AL = false>
__device__ __noinline__ int enc_o0(const uint8_t *in, uint32_t n, uint8_t *out, uint8_t *out_end,
                      uint32_t *tab_len, uint8_t **ptr_out, EncO0Smem &S, int lane,
                      const uint32_t *model = nullptr) {
    *tab_len = 0;
    *ptr_out = out_end;
    if (n == 0) return 0;
    if (model) {                          // counts from hist_kernel (kernels.cu)
        for (int j = lane; j < 256; j += 32) S.F[j] = model[j];
        __syncwarp();
    } else warp_hist8(in, n, S.F, lane);

    uint32_t fsum = round2(n);
    if (fsum > 4096) fsum = 4096;
    if (normalise_freq_warp(S.F, n, fsum, lane) < 0) return 1;
    uint32_t tl = 0;
    if (lane == 0) {                                        // rANS_static16_int.h:240-252
        uint8_t *cp = out;
        cp += put_alphabet(cp, S.F);
        for (int j = 0; j < 256; j++)
            if (S.F[j]) cp += var_put_u32(cp, S.F[j]);
        tl = (uint32_t)(cp - out);
    }
    tl = __shfl_sync(FULL, tl, 0);
    *tab_len = tl;
    if (normalise_freq_warp(S.F, fsum, 4096, lane) < 0) return 1;

    {   // cumulative starts and encoder symbols
        uint32_t f[8], loc = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) { f[t] = S.F[lane * 8 + t]; loc += f[t]; }
        uint32_t x = warp_incl_scan(loc, lane) - loc;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            if (f[t]) S.sym[lane * 8 + t] = enc_sym_init(x, f[t], 12);
            x += f[t];
        }
    }
    __syncwarp();

    // NB every lane runs the same ballots: lanes >= N (N == 4) are predicated off.
    const bool act = (N == 32) ? true : lane < N;
    OutRingT<AL> w;
    w.init(out, out_end, S.ring);
    uint32_t R = RANS_L;
    const uint32_t rem = n % N;
    uint32_t i = n - rem;
    if (rem) {                                               // symbols i..n-1 on lanes 0..rem-1
        bool on = (uint32_t)lane < rem;
        uint4 e = S.sym[on ? in[i + lane] : 0];
        R = enc_step(R, on, e, w, lane);
    }
    const uint8_t *q = in + (act ? lane : 0);
    // Symbols are fetched three groups (12 steps) ahead of their use, into three register
    // sets that are refilled right after they are consumed (no register rotation, so no
    // instruction waits on a load younger than a full trip of this loop): the DRAM latency of
    // a new 128-byte line stays off the state chain.
    uint32_t sA[4], sB[4], sC[4];
    auto fetch = [&](uint32_t (&s4)[4], uint32_t at) {       // symbols of the group ending at `at`
#pragma unroll
        for (int u = 0; u < 4; u++) s4[u] = ldg_u8(q + at - (u + 1) * N);
    };
    auto group = [&](const uint32_t (&s4)[4]) {
        w.maybe_flush(lane);
        uint4 e0 = S.sym[s4[0]], e1 = S.sym[s4[1]], e2 = S.sym[s4[2]], e3 = S.sym[s4[3]];
        R = enc_step<N == 32>(R, act, e0, w, lane);
        R = enc_step<N == 32>(R, act, e1, w, lane);
        R = enc_step<N == 32>(R, act, e2, w, lane);
        R = enc_step<N == 32>(R, act, e3, w, lane);
    };
    if (i >= 4 * 512) {
        // The symbols of 512 / N steps are 512 consecutive bytes.  The warp brings them in with one coalesced
        // 16-byte load per lane, a whole chunk ahead of their use, and parks them in shared memory (the histogram
        // is dead by now: S.F holds two chunks); a step then takes its byte from there.  One global load per
        // 512 / N steps instead of one per step, and no state-chain instruction ever waits on global memory
        // (with 4 lanes a chunk is 128 steps: the table and token streams the 4-lane coder gets are read from
        // DRAM-cold buffers, and a byte load per step stalled every step on it).
        const uint32_t st_s = (uint32_t)__cvta_generic_to_shared(S.F);
        while (i & 511) {                                    // steps above the highest chunk boundary
            w.maybe_flush(lane);
            R = enc_step<N == 32>(R, act, S.sym[q[i - N]], w, lane);
            i -= N;
        }
        const bool al16 = (((uintptr_t)in) & 15) == 0;
        auto load16 = [&](uint32_t at) -> uint4 {            // plain loads: `in` may be PACK / RLE output
            const uint8_t *p = in + at + 16 * lane;
            return al16 ? *(const uint4 *)p : ld16_any(p);
        };
        uint4 nxt = load16(i - 512);
        uint32_t buf = 0;
        while (i >= 512) {
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(st_s + buf * 512 + 16 * lane), "r"(nxt.x),
                         "r"(nxt.y), "r"(nxt.z), "r"(nxt.w) : "memory");
            __syncwarp();
            if (i >= 1024) nxt = load16(i - 1024);
            const uint32_t b = st_s + buf * 512 + (act ? lane : 0);
#pragma unroll 4
            for (int gi = 512 / (4 * N) - 1; gi >= 0; gi--) {        // groups of four steps, from the chunk's end
                uint32_t s4[4];
#pragma unroll
                for (int u = 0; u < 4; u++)
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(s4[u]) : "r"(b + N * (4 * gi + 3 - u)));
                group(s4);
            }
            i -= 512;
            buf ^= 1;
        }
    }
    if (i >= 24 * N) {
        fetch(sA, i); fetch(sB, i - 4 * N); fetch(sC, i - 8 * N);
        for (; i >= 24 * N; i -= 12 * N) {
            group(sA); fetch(sA, i - 12 * N);
            group(sB); fetch(sB, i - 16 * N);
            group(sC); fetch(sC, i - 20 * N);
        }
        group(sA); group(sB); group(sC);                     // the three groups already in registers
        i -= 12 * N;
    }
    for (; i >= 4 * N; i -= 4 * N) { fetch(sA, i); group(sA); }
    for (; i > 0; i -= N) {
        w.maybe_flush(lane);
        R = enc_step<N == 32>(R, act, S.sym[q[i - N]], w, lane);
    }
    enc_flush(R, act, N, w, lane);
    *ptr_out = w.slot + w.off;
    __syncwarp();
    return 0;
}

// ======================================================================== o1
// rans_compute_shift (rANS_static4x16pr.c:357-420) helpers
__device__ __forceinline__ double fast_log(double a) {                    // utils.h:69-72
    return (double)(__double_as_longlong(a) - 4606921278410026770LL) * 1.539095918623324e-16;
}

struct __align__(16) EncO1Smem {
    uint32_t T[256];        // o0 counts, then order-1 row totals (symbol space)
    uint8_t  rank[256];
    uint8_t  sym[256];      // rank -> symbol
    uint16_t S[256];        // per-row stored total (rank space)
    uint32_t pres[8];       // alphabet membership bitmap (symbol space)
    uint32_t pad_[4];
    union {                 // the model-building scratch is dead when the output ring starts
        uint32_t rowlen[256];   // serialised row lengths / offsets (rank space), partition cursors
        uint8_t  ring[ORING];   // output staging (OutRing)
    };
};                          // followed by dynamic storage: nsym*nsym pair counts when they fit

// serialise one row against the alphabet (rANS_static16_int.h:278-306): every
// listed symbol gets a varint, a run of z zeros becomes 0,(z-1).  With cp==null
// only the length is computed.
__device__ inline uint32_t put_freq_row(uint8_t *cp, const uint32_t *F, uint32_t n) {
    uint32_t len = 0, j = 0;
    while (j < n) {
        if (F[j]) {
            if (cp) len += var_put_u32(cp + len, F[j]); else len += var_size_u32(F[j]);
            j++;
            continue;
        }
        uint32_t z = 0;
        while (j < n && !F[j]) { z++; j++; }
        if (cp) { cp[len] = 0; cp[len + 1] = (uint8_t)(z - 1); }
        len += 2;
    }
    return len;
}

// Alverson reciprocal of enc_sym_init without the 64-bit division: ceil(2^(sh+31) / freq) for
// 2 <= freq <= 4096, 2^(sh-1) < freq <= 2^sh, by two 32-bit long-division steps.
__device__ __forceinline__ uint32_t rcp_freq_small(uint32_t freq, uint32_t sh) {
    uint32_t a = 1u << (sh + 19);                 // <= 2^31
    uint32_t q1 = a / freq, r1 = a - q1 * freq;   // r1 < 4096
    uint32_t b = r1 << 12;
    uint32_t q2 = b / freq, r2 = b - q2 * freq;
    return (q1 << 12) + q2 + (r2 ? 1u : 0u);
}
__device__ __forceinline__ uint32_t enc_sym_make4(uint32_t start, uint32_t freq, uint32_t bits) {
    if (freq < 2) return (start + (1u << bits) - 1) | (freq << 13);
    uint32_t sh = 32 - __clz(freq - 1);
    return start | (freq << 13) | ((sh - 1) << 26);
}
// fills g_rcp_freq (once per device): rcp of enc_sym_init for every possible frequency
__global__ void rcp_table_kernel() {
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f > 4096) return;
    g_rcp_freq[f] = f < 2 ? ~0u : rcp_freq_small(f, 32 - __clz(f - 1));
}

// ------------------------------------------------------------------------
// Order-1 model for large alphabets (nsym > 64): the warp walks the context rows one at a
// time, lane l holding columns 8l..8l+7 in registers, so every access to the pair counts
// and to the encoder symbols is a contiguous 32/64-byte piece per lane and nothing is
// serialised over the row length.  Same arithmetic as the per-row code in enc_o1:
//   sweep 1  row totals and the statistics of rans_compute_shift (rANS_static4x16pr.c:357-420)
//   sweep 2  normalise_freq (rANS_static16_int.h:97-146), encode_freq_d (:278-306) at a
//            running offset, scaled starts and encoder symbols (rANS_word.h:201-272)
// H rows are left normalised (as the per-row code leaves them).  Returns 0 ok, 1 fail.
// ------------------------------------------------------------------------
__device__ inline int enc_o1_rows_wide(uint32_t *H, uint32_t nsym, EncO1Smem &S, const uint8_t *in, uint32_t n,
                                       uint8_t *out, uint32_t hdr, uint32_t *symtab, int lane,
                                       uint32_t *shift_out, uint32_t *tl_out) {
    const uint32_t j0 = (uint32_t)lane * 8;
    const uint32_t last_rank = S.rank[in[n - 1]];
    const bool vec = (nsym & 3) == 0;
    auto load_row = [&](const uint32_t *row, uint32_t (&f)[8]) {
        if (vec && j0 + 8 <= nsym) {
            uint4 a = *(const uint4 *)(row + j0), b = *(const uint4 *)(row + j0 + 4);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
#pragma unroll
            for (int t = 0; t < 8; t++) f[t] = j0 + t < nsym ? row[j0 + t] : 0;
        }
    };
    // ---- sweep 1
    double e10 = 0, e12 = 0;
    uint32_t max_tot = 0;
    for (uint32_t i = 0; i < nsym; i++) {
        uint32_t f[8], loc = 0;
        load_row(H + (size_t)i * nsym, f);
#pragma unroll
        for (int t = 0; t < 8; t++) loc += f[t];
        const uint32_t Ti = warp_sum(loc) + (i == last_rank ? 1u : 0u);
        if (lane == 0) S.T[i] = Ti;
        if (!Ti) { if (lane == 0) S.S[i] = 0; continue; }
        uint32_t max_val = round2(Ti);
        uint32_t cnt = 0;                           // ns | sm10 << 10 | sm12 << 20
#pragma unroll
        for (int t = 0; t < 8; t++) {
            if (!f[t]) continue;
            cnt += 1;
            if ((uint64_t)f[t] * 1025 <= max_val) cnt += 1u << 10;     // max_val / f > 1024
            if ((uint64_t)f[t] * 4097 <= max_val) cnt += 1u << 20;     // max_val / f > 4096
        }
        cnt = warp_sum(cnt);
        const uint32_t ns = cnt & 1023, sm10 = (cnt >> 10) & 1023, sm12 = cnt >> 20;
        const double l10 = log((double)(1024 + sm10)), l12 = log((double)(4096 + sm12));
        const double T_slow = (double)4096 / Ti, T_fast = (double)1024 / Ti;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            if (!f[t]) continue;
            double a = f[t] * T_fast, b = f[t] * T_slow;
            e10 -= f[t] * (fast_log(a > 1 ? a : 1) - l10);
            e12 -= f[t] * (fast_log(b > 1 ? b : 1) - l12);
            e10 += 1.3;
            e12 += 4.7;
        }
        if (ns < 64 && max_val > 128) max_val /= 2;
        if (max_val > 1024) max_val /= 2;
        if (max_val > 4096) max_val = 4096;
        if (lane == 0) S.S[i] = (uint16_t)max_val;
        if (max_tot < max_val) max_tot = max_val;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        e10 += __shfl_xor_sync(FULL, e10, o);
        e12 += __shfl_xor_sync(FULL, e12, o);
    }
    const uint32_t shift = (e10 / e12 < 1.01 || max_tot <= 1024) ? 10 : 12;
    __syncwarp();

    // ---- sweep 2
    uint32_t off = hdr;
    int err = 0;
    for (uint32_t i = 0; i < nsym; i++) {
        const uint32_t Ti = S.T[i];
        if (!Ti) continue;
        uint32_t *row = H + (size_t)i * nsym;
        uint32_t f[8];
        load_row(row, f);
        uint32_t mv = S.S[i];
        if (shift == 10 && mv > 1024) mv = 1024;
        // normalise_freq(row, Ti, mv)
        {
            uint32_t size = Ti;
            for (int pass = 0; pass < 2; pass++) {
                const uint64_t tr = ((uint64_t)mv << 31) / size + (uint32_t)((1 << 30) / (int)size);
                uint32_t top = 0, arg = 0, sum = 0;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    uint32_t v = f[t];
                    if (!v) continue;
                    if (top < v) { top = v; arg = j0 + t; }
                    v = (uint32_t)((v * tr) >> 31);
                    if (!v) v = 1;
                    f[t] = v;
                    sum += v;
                }
                sum = warp_sum(sum);
#pragma unroll
                for (int o = 16; o; o >>= 1) {      // max, ties -> lowest index
                    uint32_t t2 = __shfl_xor_sync(FULL, top, o), a2 = __shfl_xor_sync(FULL, arg, o);
                    if (t2 > top || (t2 == top && a2 < arg)) { top = t2; arg = a2; }
                }
                const uint32_t big = top ? arg : 0;
                uint32_t mine = 0;
#pragma unroll
                for (int t = 0; t < 8; t++) if ((big & 7) == (uint32_t)t) mine = f[t];
                const uint32_t fb = __shfl_sync(FULL, mine, big >> 3);
                int adjust = (int)mv - (int)sum;
                const bool own = (big >> 3) == (uint32_t)lane;
                if (adjust >= 0 || (fb > (uint32_t)-adjust && (pass == 1 || fb / 2 >= (uint32_t)-adjust))) {
                    if (own) {
#pragma unroll
                        for (int t = 0; t < 8; t++) if ((big & 7) == (uint32_t)t) f[t] += adjust;
                    }
                    break;
                }
                if (pass == 0) { size = sum; continue; }
                // greedy shave (rare): serial over the row through shared memory
                uint32_t *tmp = S.rowlen;
#pragma unroll
                for (int t = 0; t < 8; t++) tmp[j0 + t] = f[t];
                __syncwarp();
                if (lane == 0) {
                    adjust += (int)fb - 1;
                    tmp[big] = 1;
                    for (uint32_t j = 0; adjust && j < nsym; j++) {
                        if (tmp[j] < 2) continue;
                        int d = (tmp[j] > (uint32_t)-adjust) ? adjust : 1 - (int)tmp[j];
                        tmp[j] += d;
                        adjust -= d;
                    }
                    if (!tmp[big]) err = 1;
                }
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 8; t++) f[t] = tmp[j0 + t];
                __syncwarp();
            }
        }
        if (lane == 0) S.S[i] = (uint16_t)mv;
        // the row stays normalised in H (the table coder and later readers expect it)
        if (vec && j0 + 8 <= nsym) {
            *(uint4 *)(row + j0) = make_uint4(f[0], f[1], f[2], f[3]);
            *(uint4 *)(row + j0 + 4) = make_uint4(f[4], f[5], f[6], f[7]);
        } else {
#pragma unroll
            for (int t = 0; t < 8; t++) if (j0 + t < nsym) row[j0 + t] = f[t];
        }
        // zero bitmap of the row: zb[u] covers columns 32u..32u+31 (columns >= nsym read as non-zero)
        uint32_t m8 = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) if (j0 + t < nsym && !f[t]) m8 |= 1u << t;
        uint32_t wz = m8 << (8 * (lane & 3));
        wz |= __shfl_xor_sync(FULL, wz, 1);
        wz |= __shfl_xor_sync(FULL, wz, 2);
        uint32_t zb[8];
#pragma unroll
        for (int u = 0; u < 8; u++) zb[u] = __shfl_sync(FULL, wz, 4 * u);
        // first non-zero column at or after the start of word u (256 if none)
        uint32_t nz[9];
        nz[8] = 256;
#pragma unroll
        for (int u = 7; u >= 0; u--) nz[u] = ~zb[u] ? 32 * u + __ffs(~zb[u]) - 1 : nz[u + 1];
        // bytes per entry and scaled frequency, prefix over the row in column order
        int sh = 0;
        while ((mv << sh) < (1u << shift)) sh++;
        uint32_t len[8], run[8], tot8 = 0;
        const uint32_t myw = lane >> 2;                       // word holding this lane's 8 columns
        uint32_t zw = 0, nzn = 256;
#pragma unroll
        for (int u = 0; u < 8; u++) if (myw == (uint32_t)u) { zw = zb[u]; nzn = nz[u + 1]; }
        const uint32_t prevw_top = myw ? 0u : 0u;
        (void)prevw_top;
        uint32_t zprev = 0;                                   // bit 31 of the previous word
#pragma unroll
        for (int u = 1; u < 8; u++) if (myw == (uint32_t)u) zprev = zb[u - 1] >> 31;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const uint32_t j = j0 + t, bit = j & 31;
            uint32_t l = 0, r = 0;
            if (j < nsym) {
                if (f[t]) l = f[t] >= 128 ? 2 : 1;
                else {
                    const uint32_t pz = bit ? (zw >> (bit - 1)) & 1 : zprev;
                    if (!pz) {
                        const uint32_t w = (~zw) >> bit;
                        const uint32_t end = w ? j + __ffs(w) - 1 : nzn;
                        l = 2;
                        r = min(end, nsym) - j;
                    }
                }
            }
            len[t] = l; run[t] = r;
            tot8 += l | ((f[t] << sh) << 16);
        }
        uint32_t ex = warp_incl_scan(tot8, lane);
        const uint32_t rowbytes = __shfl_sync(FULL, ex, 31) & 0xffff;
        ex -= tot8;
        uint32_t o = off + (ex & 0xffff), x = ex >> 16;
        uint32_t e8[8];
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const uint32_t fs = f[t] << sh;
            if (len[t]) {
                if (f[t]) {
                    if (len[t] == 2) { out[o] = (uint8_t)(0x80 | (f[t] >> 7)); out[o + 1] = (uint8_t)(f[t] & 0x7f); }
                    else out[o] = (uint8_t)f[t];
                } else { out[o] = 0; out[o + 1] = (uint8_t)(run[t] - 1); }
                o += len[t];
            }
            e8[t] = enc_sym_make4(x, fs, shift);
            x += fs;
        }
        uint32_t *srow = symtab + (size_t)i * nsym;
        if (vec && j0 + 8 <= nsym) {
            *(uint4 *)(srow + j0) = make_uint4(e8[0], e8[1], e8[2], e8[3]);
            *(uint4 *)(srow + j0 + 4) = make_uint4(e8[4], e8[5], e8[6], e8[7]);
        } else {
#pragma unroll
            for (int t = 0; t < 8; t++) if (j0 + t < nsym) srow[j0 + t] = e8[t];
        }
        off += rowbytes;
    }
    if (__any_sync(FULL, err)) return 1;
    *shift_out = shift;
    *tl_out = off;
    return 0;
}

// ------------------------------------------------------------------------
// Pair counts for large alphabets without random read-modify-write on a 256 KiB table per
// stream (thousands of streams thrash L2 and every increment becomes a DRAM round trip):
//   1. the number of pairs per context follows from the order-0 counts, so the pairs can be
//      dealt straight into per-bucket regions (bucket = 8 consecutive context ranks) as
//      16-bit keys (rank(prev) & 7) * nsym + rank(cur): sequential writes to <= 32 cursors;
//   2. each bucket is counted in shared memory (8 rows x nsym counters) and its rows are
//      written to H once, contiguously.  H needs no zeroing.
// T = order-0 counts in symbol space; cur = 32 cursors in shared memory; cnt = 8*nsym words.
// ------------------------------------------------------------------------
__device__ inline void pair_counts_partitioned(const uint8_t *in, uint32_t n, uint32_t nsym, const uint32_t *T,
                                               const uint8_t *rank, const uint8_t *symof, uint32_t *H,
                                               uint16_t *keys, uint32_t *cur, uint32_t *cnt, int lane) {
    const uint32_t nb = (nsym + 7) >> 3;
    // pairs whose context is symbol s: occurrences of s except as the last byte, plus the
    // virtual 0 in front of the first byte
    uint32_t bsize = 0;
    if ((uint32_t)lane < nb) {
        for (uint32_t r = lane * 8; r < min(nsym, (uint32_t)lane * 8 + 8); r++) {
            const uint32_t s = symof[r];
            bsize += T[s] - (s == in[n - 1] ? 1u : 0u) + (s == 0 ? 1u : 0u);
        }
    }
    const uint32_t padded = (bsize + 7) & ~7u;                 // regions start on 16-byte boundaries
    const uint32_t boff = warp_incl_scan(padded, lane) - padded;
    cur[lane] = boff;
    __syncwarp();
    auto deal = [&](uint32_t rp, uint32_t rc) {
        const uint32_t pos = atomicAdd(&cur[rp >> 3], 1u);
        keys[pos] = (uint16_t)((rp & 7) * nsym + rc);
    };
    {
        uint32_t head = (uint32_t)((16 - ((uintptr_t)in & 15)) & 15);
        if (head > n) head = n;
        if ((uint32_t)lane < head) deal(rank[lane ? in[lane - 1] : 0], rank[in[lane]]);
        const uint8_t *p = in + head;
        const uint32_t rest = n - head, nv = rest >> 4;
        const uint4 *v = (const uint4 *)p;
        uint32_t carry_last = head ? in[head - 1] : 0;
        uint4 q = (uint32_t)lane < nv ? v[lane] : make_uint4(0, 0, 0, 0);
        for (uint32_t base = 0; base < nv; base += 32) {
            const uint32_t i = base + lane;
            const bool on = i < nv;
            const uint4 qn = i + 32 < nv ? v[i + 32] : make_uint4(0, 0, 0, 0);   // next round, in flight
            const uint32_t lastb = q.w >> 24;
            uint32_t pb = __shfl_up_sync(FULL, lastb, 1);
            if (lane == 0) pb = carry_last;
            const uint32_t nact = min(32u, nv - base);
            carry_last = __shfl_sync(FULL, lastb, nact - 1);
            if (on) {
                const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
                uint32_t rp = rank[pb];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const uint32_t rc = rank[(w4[a] >> (8 * b)) & 0xff];
                        deal(rp, rc);
                        rp = rc;
                    }
            }
            q = qn;
        }
        for (uint32_t i = (nv << 4) + lane; i < rest; i += 32) {
            const uint32_t pos = head + i;
            deal(rank[pos ? in[pos - 1] : 0], rank[in[pos]]);
        }
    }
    __threadfence_block();
    __syncwarp();
    const uint32_t words = 8 * nsym;
    for (uint32_t b = 0; b < nb; b++) {
        for (uint32_t j = lane * 4; j < words; j += 128) *(uint4 *)(cnt + j) = make_uint4(0, 0, 0, 0);
        __syncwarp();
        const uint32_t o = __shfl_sync(FULL, boff, b), sz = __shfl_sync(FULL, bsize, b);
        const uint4 *kv = (const uint4 *)(keys + o);             // 8 keys per load
        for (uint32_t i = lane; i * 8 < sz; i += 32) {
            const uint4 k8 = kv[i];
            const uint32_t w4[4] = {k8.x, k8.y, k8.z, k8.w};
            const uint32_t left = sz - i * 8;
#pragma unroll
            for (int a = 0; a < 8; a++)
                if ((uint32_t)a < left) atomicAdd(&cnt[(w4[a >> 1] >> (16 * (a & 1))) & 0xffff], 1u);
        }
        __syncwarp();
        const uint32_t rows = min(8u, nsym - 8 * b);
        uint32_t *dst = H + (size_t)8 * b * nsym;
        for (uint32_t j = lane; j < rows * nsym; j += 32) dst[j] = cnt[j];
        __syncwarp();
    }
    __threadfence_block();
    __syncwarp();
}

// ------------------------------------------------------------------------
// The order-1 state chains (rANS_static32x16pr.c:457-525, rANS_static4x16pr.c:460-518): lane z owns
// [z*seg,(z+1)*seg); lane N-1 also the tail; every symbol is coded in the context of its predecessor, a
// lane's first symbol in context 0.  symtab = 4-byte encoder symbols indexed (rank(ctx), rank(sym)), in
// shared memory when sym_smem; rank (shared memory) maps symbols to ranks; ring = ORING bytes of output staging.
// ------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void enc_o1_payload(const uint8_t *in, uint32_t n, uint8_t *out, uint8_t *out_end,
                                               uint8_t **ptr_out, uint8_t *ring, const uint8_t *rank,
                                               const uint32_t *symtab, uint32_t nsym, uint32_t shift, bool sym_smem,
                                               int lane) {
    const uint32_t seg = n / N;
    const bool act = lane < N;
    OutRing w;
    w.init(out, out_end, ring);
    uint32_t R = RANS_L;
    {   // tail on lane N-1, from the end down to N*seg
        const bool lastl = lane == N - 1;
        for (uint32_t p = n - 1; p >= N * seg && p > 0; p--) {
            EncSym e = make_uint4(0, 0, 0, 0);
            if (lastl) e = enc_sym_unpack(symtab[rank[in[p - 1]] * nsym + rank[in[p]]], shift);
            w.maybe_flush(lane);
            R = enc_step(R, lastl, e, w, lane);
        }
    }
    const uint8_t *q = in + (size_t)(act ? lane : 0) * seg;
    uint32_t rs = act && seg ? rank[q[seg - 1]] : 0;         // rank of the symbol being coded
    uint32_t kstart = seg;
    __syncwarp();                                            // enter the hot loop converged
    if (N == 32 && sym_smem && seg >= 32 && ((((uintptr_t)in) | seg) & 15) == 0) {
        // Hot loop: lane segments are 16-byte aligned, so every lane reads its symbols with
        // 128-bit loads (one group of 16 ahead) and the encoder symbols come from shared memory.
        const uint32_t rank_s = (uint32_t)__cvta_generic_to_shared(rank);
        const uint32_t sym_s = (uint32_t)__cvta_generic_to_shared(symtab);
        const uint4 *v = (const uint4 *)q;
        const uint32_t J = seg >> 4;
        uint4 cur = ldg_u128(v + J - 1), nxt = ldg_u128(v + J - 2);
        auto byte_of = [](const uint4 &x, int b) {
            uint32_t w = b < 4 ? x.x : b < 8 ? x.y : b < 12 ? x.z : x.w;
            return (w >> (8 * (b & 3))) & 0xff;
        };
        auto lds_sym = [&](uint32_t a) {
            uint32_t r;
            asm("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
            return enc_sym_unpack(r, shift);
        };
        auto rank_of = [&](uint32_t b) {
            uint32_t r;
            asm("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(rank_s + b));
            return r;
        };
        // the encoder symbol of the NEXT step is fetched (rank look-up, table look-up, unpack)
        // while the current step runs: none of it depends on the state
        uint32_t rc = rank_of(byte_of(cur, 14));
        EncSym e = lds_sym(sym_s + (rc * nsym + rs) * 4);
        for (uint32_t j = J - 1; j >= 1; j--) {
            uint4 nn = j >= 2 ? ldg_u128(v + j - 2) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int b = 15; b >= 0; b--) {
                // next step codes byte b-1 in the context of byte b-2 (crossing into nxt at the low end)
                uint32_t nb = b >= 2 ? byte_of(cur, b - 2) : byte_of(nxt, 14 + b);
                uint32_t rn = rank_of(nb);
                EncSym en = lds_sym(sym_s + (rn * nsym + rc) * 4);
                if ((b & 3) == 3) w.maybe_flush(lane);
                R = enc_step<true>(R, true, e, w, lane);
                e = en;
                rs = rc;
                rc = rn;
            }
            cur = nxt;
            nxt = nn;
        }
#pragma unroll
        for (int b = 15; b >= 1; b--) {                      // group 0: its byte 0 is the lane's first symbol
            uint32_t rn = b >= 2 ? rank_of(byte_of(cur, b - 2)) : 0;
            EncSym en = b >= 2 ? lds_sym(sym_s + (rn * nsym + rc) * 4) : e;
            if ((b & 3) == 3) w.maybe_flush(lane);
            R = enc_step<true>(R, true, e, w, lane);
            e = en;
            rs = rc;
            rc = rn;
        }
        kstart = 1;
    } else if (seg >= 32) {
        // Any lane count, any alignment, encoder symbols in shared or global memory: the 16 input bytes
        // and the 16 symbols of a group are requested together before the group's 16 steps run, so one
        // round trip to memory is paid per group instead of per step (a 4-lane stream of 256 KiB is
        // 65 536 steps per lane: with one exposed global load each it took ~35 ms).  Groups are counted
        // from the END of the lane's segment; the seg % 16 bytes at its start go through the loop below.
        const uint32_t rank_s = (uint32_t)__cvta_generic_to_shared(rank);
        const uint32_t J = seg >> 4, lead = seg & 15;
        const uint8_t *g0 = q + lead;                          // group j = bytes g0[16j .. 16j+15]
        auto rank_of = [&](uint32_t b) {
            uint32_t r;
            asm("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(rank_s + b));
            return r;
        };
        const uint32_t rank0 = rank_of(0);
        uint4 cur = ld16_any(g0 + 16 * (J - 1));
        for (uint32_t j = J; j-- > 0;) {
            uint4 prv = j ? ld16_any(g0 + 16 * (j - 1)) : make_uint4(0, 0, 0, 0);
            const uint32_t w4[4] = {cur.x, cur.y, cur.z, cur.w};
            uint32_t rk[17];
            rk[0] = j ? rank_of(prv.w >> 24) : (lead ? rank_of(g0[-1]) : rank0);
#pragma unroll
            for (int b = 0; b < 16; b++) rk[b + 1] = rank_of((w4[b >> 2] >> (8 * (b & 3))) & 0xff);
            uint32_t ev[16];
#pragma unroll
            for (int b = 0; b < 16; b++) ev[b] = symtab[rk[b] * nsym + rk[b + 1]];      // (written by this kernel: no ld.nc)
#pragma unroll
            for (int b = 15; b >= 0; b--) {
                if (b == 0 && j == 0 && !lead) break;        // the lane's first symbol is coded below
                if ((b & 3) == 3) w.maybe_flush(lane);
                R = enc_step<N == 32>(R, act, enc_sym_unpack(ev[b], shift), w, lane);
            }
            rs = lead ? rk[0] : rk[1];                       // rank of the next symbol to code
            cur = prv;
        }
        kstart = lead ? lead : 1;
    }
    if (kstart > 1 && kstart <= 16 && seg >= 16) {
        // the few bytes left at the start of the segment: fetched with one load, not one dependent global
        // byte load per step (short streams -- STRIPE sub-streams -- spend most of their steps here)
        const uint4 h16 = ld16_any(q);
        const uint32_t w4[4] = {h16.x, h16.y, h16.z, h16.w};
        for (uint32_t k = kstart; k-- > 1;) {
            const uint32_t rc = rank[(w4[(k - 1) >> 2] >> (8 * ((k - 1) & 3))) & 0xff];
            EncSym e = enc_sym_unpack(symtab[rc * nsym + rs], shift);
            w.maybe_flush(lane);
            R = enc_step<N == 32>(R, act, e, w, lane);
            rs = rc;
        }
        kstart = 1;
    }
    for (uint32_t k = kstart; k-- > 1;) {
        uint32_t rc = rank[q[k - 1]];
        EncSym e = enc_sym_unpack(symtab[rc * nsym + rs], shift);
        w.maybe_flush(lane);
        R = enc_step<N == 32>(R, act, e, w, lane);
        rs = rc;
    }
    if (seg) {
        EncSym e = enc_sym_unpack(symtab[rank[0] * nsym + rs], shift);
        w.maybe_flush(lane);
        R = enc_step<N == 32>(R, act, e, w, lane);
    }
    enc_flush(R, act, N, w, lane);
    *ptr_out = w.slot + w.off;
    __syncwarp();
}


template <int N>
__device__ int enc_o1(const uint8_t *in, uint32_t n, uint8_t *out, uint8_t *out_end,
                      uint32_t *tab_len, uint8_t **ptr_out, EncO1Smem &S, uint8_t *dyn, uint32_t dyn_bytes,
                      const Pool &pool, int lane, const uint32_t *model = nullptr) {
    *tab_len = 0;
    *ptr_out = out_end;
    if (N == 32 && n < 32) return 1;
    const uint32_t seg = n / N;

    // ---- alphabet = symbols present, plus 0 (rANS_static16_int.h:357-361)
    if (model) {                          // counts and pair counts from hist_kernel (kernels.cu)
        for (int j = lane; j < 256; j += 32) S.T[j] = model[j];
        __syncwarp();
    } else warp_hist8(in, n, S.T, lane);
    uint32_t nsym;
    {
        uint32_t loc = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) { int j = lane * 8 + t; loc += (S.T[j] || j == 0) ? 1 : 0; }
        uint32_t incl = warp_incl_scan(loc, lane);
        nsym = __shfl_sync(FULL, incl, 31);
        uint32_t r = incl - loc;
        uint32_t bits = 0;
        for (int t = 0; t < 8; t++) {
            int j = lane * 8 + t;
            bool p = S.T[j] || j == 0;
            S.rank[j] = p ? (uint8_t)r : 0xff;
            if (p) { S.sym[r++] = (uint8_t)j; bits |= 1u << t; }
        }
        // lane l holds symbols 8l..8l+7: four lanes make one 32-bit word
        uint32_t w = bits << (8 * (lane & 3));
        w |= __shfl_xor_sync(FULL, w, 1);
        w |= __shfl_xor_sync(FULL, w, 2);
        if ((lane & 3) == 0) S.pres[lane >> 2] = w;
    }
    __syncwarp();

    // ---- pair counts H[rank(prev)][rank(cur)], first symbol follows 0 (utils.h:279-357)
    // Pair counts H: with counts from hist_kernel they are used in place in global memory (L2)
    // and the dynamic shared memory holds only the encoder symbols (the table coder's scratch
    // aliases it, it runs before they are built).  Otherwise H sits in shared memory in front
    // of the symbols when it fits, else in the scratch pool.
    uint32_t *H;
    const uint32_t hw = nsym * nsym;
    const bool h_global = model != nullptr;
    const uint32_t h_bytes = h_global ? 0 : max((hw * 4 + 15) & ~15u, (uint32_t)sizeof(EncO0Smem));
    const bool h_smem = !h_global && h_bytes <= dyn_bytes;
    const bool sym_smem = (h_global || h_smem) && h_bytes + hw * 4 <= dyn_bytes;
    EncO0Smem *o0s = (EncO0Smem *)dyn;           // dead H, or not-yet-built symbols
    if (h_global) H = const_cast<uint32_t *>(model) + MODEL_HDR_WORDS;
    else if (h_smem) H = (uint32_t *)dyn;
    else { H = (uint32_t *)pool_alloc(pool, hw * 4, lane); if (!H) return 2; }
    if (!model && nsym > 64 && 8 * nsym * 4 <= dyn_bytes && n >= 2) {
        uint16_t *keys = (uint16_t *)pool_alloc(pool, 2 * n + 2048, lane);
        if (!keys) return 2;
        pair_counts_partitioned(in, n, nsym, S.T, S.rank, S.sym, H, keys, S.rowlen, (uint32_t *)dyn, lane);
    } else if (!model) {
        for (uint32_t j = lane; j < hw; j += 32) H[j] = 0;
        __syncwarp();
        const uint8_t *rank = S.rank;
        uint32_t head = (uint32_t)((16 - ((uintptr_t)in & 15)) & 15);
        if (head > n) head = n;
        if ((uint32_t)lane < head) {
            uint32_t prev = lane ? in[lane - 1] : 0;
            atomicAdd(&H[rank[prev] * nsym + rank[in[lane]]], 1u);
        }
        const uint8_t *p = in + head;
        uint32_t rest = n - head, nv = rest >> 4;
        const uint4 *v = (const uint4 *)p;
        // carry = rank of the byte just before this lane's 16 bytes
        uint32_t carry_last = head ? in[head - 1] : 0;      // byte before the vector body
        for (uint32_t base = 0; base < nv; base += 32) {
            uint32_t i = base + lane;
            bool on = i < nv;
            uint4 q = on ? ldg_u128(v + i) : make_uint4(0, 0, 0, 0);
            uint32_t lastb = q.w >> 24;
            uint32_t pb = __shfl_up_sync(FULL, lastb, 1);
            if (lane == 0) pb = carry_last;
            // last byte of the last active lane feeds lane 0 of the next round
            uint32_t nact = min(32u, nv - base);
            carry_last = __shfl_sync(FULL, lastb, nact - 1);
            if (on) {
                uint32_t w4[4] = {q.x, q.y, q.z, q.w};
                uint32_t rp = rank[pb], last = 0xffffffffu, cnt = 0;
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        uint32_t rc = rank[(w4[a] >> (8 * b)) & 0xff];
                        uint32_t idx = rp * nsym + rc;
                        if (idx == last) cnt++;
                        else { if (cnt) atomicAdd(&H[last], cnt); last = idx; cnt = 1; }
                        rp = rc;
                    }
                atomicAdd(&H[last], cnt);
            }
        }
        for (uint32_t i = (nv << 4) + lane; i < rest; i += 32) {
            uint32_t pos = head + i;
            uint32_t prev = pos ? in[pos - 1] : 0;
            atomicAdd(&H[rank[prev] * nsym + rank[in[pos]]], 1u);
        }
    }
    // lanes 1..N-1 start in context 0 (rANS_static16_int.h:325-327)
    if (lane >= 1 && lane < N) atomicAdd(&H[S.rank[0] * nsym + S.rank[in[lane * seg]]], 1u);
    __syncwarp();
    uint32_t *symtab;
    if (sym_smem) symtab = (uint32_t *)(dyn + h_bytes);
    else { symtab = (uint32_t *)pool_alloc(pool, hw * 4, lane); if (!symtab) return 2; }
    uint32_t shift = 12, tl = 0;
    const bool wide = nsym > 64;              // large alphabets: row-at-a-time, lanes across columns
    // Pair counts from hist_kernel, small alphabet: the passes below walk a row per lane, element by element, and on
    // the matrix in global memory every element was an L2 round trip of its own (8 % of the kernel's samples sat on
    // those loads).  The matrix is brought into the shared-memory area of the encoder symbols, worked on there, and
    // turned into the symbols in place; the table coder's scratch aliases that area, so around it the normalised
    // rows go back to their global home for a moment.
    uint32_t *const Hg = H;
    const bool staged = h_global && sym_smem && !wide;
    if (staged) {
        for (uint32_t j = lane; j < hw; j += 32) symtab[j] = Hg[j];
        __syncwarp();
        H = symtab;
    }
    if (wide) {
        uint32_t hdr = 0;
        if (lane == 0) {                      // alphabet of the contexts, 0 forced in (:357-361)
            uint8_t *cp = out;
            *cp++ = 0;
            auto present = [&](int s) { return (S.pres[s >> 5] >> (s & 31)) & 1; };
            int j = 0;
            while (j < 256) {
                if (!present(j)) { j++; continue; }
                *cp++ = (uint8_t)j;
                if (j && present(j - 1)) {
                    int k = j + 1;
                    while (k < 256 && present(k)) k++;
                    *cp++ = (uint8_t)(k - (j + 1));
                    j = k;
                } else j++;
            }
            *cp++ = 0;
            hdr = (uint32_t)(cp - out);
        }
        hdr = __shfl_sync(FULL, hdr, 0);
        __syncwarp();
        if (enc_o1_rows_wide(H, nsym, S, in, n, out, hdr, symtab, lane, &shift, &tl)) return 1;
        __threadfence_block();
        __syncwarp();
        out[0] = (uint8_t)(shift << 4);
    } else {
        // row totals; the last symbol's total gets one extra (utils.h:311,345)
        for (uint32_t i0 = 0; i0 < nsym; i0 += 32) { const uint32_t i = i0 + lane; if (i >= nsym) continue;
            uint32_t t = 0;
            for (uint32_t j = 0; j < nsym; j++) t += H[i * nsym + j];
            if (S.sym[i] == in[n - 1]) t++;
            S.T[i] = t;                     // now rank space
        }
        __syncwarp();

        // ---- precision: 10 or 12 bits (rANS_static4x16pr.c:357-420), lane per row
        double e10 = 0, e12 = 0;
        uint32_t max_tot = 0;
        for (uint32_t i0 = 0; i0 < nsym; i0 += 32) { const uint32_t i = i0 + lane; if (i >= nsym) continue;
            const uint32_t *row = H + i * nsym;
            uint32_t Ti = S.T[i];
            if (!Ti) { S.S[i] = 0; continue; }
            uint32_t max_val = round2(Ti);
            int ns = 0, sm10 = 0, sm12 = 0;
            for (uint32_t j = 0; j < nsym; j++) {
                uint32_t f = row[j];
                if (f && max_val / f > 1024) sm10++;
                if (f && max_val / f > 4096) sm12++;
            }
            double l10 = log((double)(1024 + sm10)), l12 = log((double)(4096 + sm12));
            double T_slow = (double)4096 / Ti, T_fast = (double)1024 / Ti;
            for (uint32_t j = 0; j < nsym; j++) {
                uint32_t f = row[j];
                if (!f) continue;
                ns++;
                double a = f * T_fast, b = f * T_slow;
                e10 -= f * (fast_log(a > 1 ? a : 1) - l10);
                e12 -= f * (fast_log(b > 1 ? b : 1) - l12);
                e10 += 1.3;
                e12 += 4.7;
            }
            if (ns < 64 && max_val > 128) max_val /= 2;
            if (max_val > 1024) max_val /= 2;
            if (max_val > 4096) max_val = 4096;
            S.S[i] = (uint16_t)max_val;
            if (max_tot < max_val) max_tot = max_val;
        }
    #pragma unroll
        for (int o = 16; o; o >>= 1) {
            e10 += __shfl_xor_sync(FULL, e10, o);
            e12 += __shfl_xor_sync(FULL, e12, o);
            max_tot = max(max_tot, __shfl_xor_sync(FULL, max_tot, o));
        }
        shift = (e10 / e12 < 1.01 || max_tot <= 1024) ? 10 : 12;

        // ---- rows: normalise to the stored total, measure, serialise, scale, symbols
        int err = 0;
        for (uint32_t i0 = 0; i0 < nsym; i0 += 32) { const uint32_t i = i0 + lane; if (i >= nsym) continue;
            uint32_t *row = H + i * nsym;
            uint32_t Ti = S.T[i];
            if (!Ti) { S.rowlen[i] = 0; continue; }
            uint32_t mv = S.S[i];
            if (shift == 10 && mv > 1024) mv = 1024;
            if (normalise_freq_row(row, nsym, Ti, mv) < 0) err = 1;
            S.S[i] = (uint16_t)mv;
            S.rowlen[i] = put_freq_row(nullptr, row, nsym);
        }
        if (__any_sync(FULL, err)) return 1;
        __syncwarp();
        uint32_t hdr = 0;
        if (lane == 0) {
            // alphabet of the CONTEXTS that have a row, with 0 forced in (:357-361)
            uint32_t *A = S.rowlen + 0;            // reuse not possible: build a private mark array
            (void)A;
            uint8_t *cp = out;
            *cp++ = 0;
            // put_alphabet over symbol space: mark = row total != 0 (or symbol 0)
            int j = 0;
            // listed = occurs in the data (or is 0); every such symbol has a non-zero row total.
            // (rank 255 is a valid rank, so presence is kept separately from rank[])
            auto present = [&](int s) { return (S.pres[s >> 5] >> (s & 31)) & 1; };
            while (j < 256) {
                if (!present(j)) { j++; continue; }
                *cp++ = (uint8_t)j;
                if (j && present(j - 1)) {
                    int k = j + 1;
                    while (k < 256 && present(k)) k++;
                    *cp++ = (uint8_t)(k - (j + 1));
                    j = k;
                } else j++;
            }
            *cp++ = 0;
            hdr = (uint32_t)(cp - out);
            uint32_t off = hdr;                      // exclusive scan of row lengths
            for (uint32_t i = 0; i < nsym; i++) { uint32_t l = S.rowlen[i]; S.rowlen[i] = off; off += l; }
            hdr = off;
        }
        tl = __shfl_sync(FULL, hdr, 0);
        __syncwarp();
        for (uint32_t i0 = 0; i0 < nsym; i0 += 32) { const uint32_t i = i0 + lane; if (i >= nsym) continue;
            if (S.T[i]) put_freq_row(out + S.rowlen[i], H + i * nsym, nsym);
        }
        __threadfence_block();
        __syncwarp();
        out[0] = (uint8_t)(shift << 4);
    }

    // the table, once complete, may itself go through the 4-lane order-0 coder (:396-412)
    int pool_fail = 0;
    auto compress_table = [&]() {
        if (tl <= 1000) return;
        uint32_t usz = tl - 1;
        uint32_t cb = compress_bound(usz, 0) - 20;
        uint8_t *tmp = pool_alloc(pool, cb + 16, lane);
        if (!tmp) { pool_fail = 1; return; }
        uint32_t ctab = 0;
        uint8_t *cptr = nullptr;
        __syncwarp();
        if (enc_o0<4>(out + 1, usz, tmp, tmp + (cb & ~1u), &ctab, &cptr, *o0s, lane) == 0) {
            uint32_t pay = (uint32_t)(tmp + (cb & ~1u) - cptr);
            uint32_t csz = ctab + pay;
            if (csz + 6 < tl) {
                uint32_t h = 1;
                if (lane == 0) {
                    out[0] |= 1;
                    h += var_put_u32(out + h, usz);
                    h += var_put_u32(out + h, csz);
                }
                h = __shfl_sync(FULL, h, 0);
                __syncwarp();
                warp_copy(out + h, tmp, ctab, lane);
                warp_copy(out + h + ctab, cptr, pay, lane);
                tl = h + csz;
            }
        }
        __syncwarp();
    };
    if (h_global || wide) {                       // its scratch aliases the symbol area
        const bool park = staged && tl > 1000;    // (the table coder does nothing for shorter tables)
        if (park) { for (uint32_t j = lane; j < hw; j += 32) Hg[j] = H[j]; __syncwarp(); }
        compress_table();
        if (park) { for (uint32_t j = lane; j < hw; j += 32) H[j] = Hg[j]; __syncwarp(); }
    }
    if (!wide)
    for (uint32_t i0 = 0; i0 < nsym; i0 += 32) { const uint32_t i = i0 + lane; if (i >= nsym) continue;
        uint32_t *row = H + i * nsym;
        if (!S.T[i]) continue;
        uint32_t mv = S.S[i];
        int sh = 0;
        while ((mv << sh) < (1u << shift)) sh++;
        uint32_t x = 0;
        for (uint32_t j = 0; j < nsym; j++) {
            uint32_t f = row[j] << sh;
            symtab[i * nsym + j] = enc_sym_make4(x, f, shift);
            x += f;
        }
    }
    __threadfence_block();
    __syncwarp();
    if (!h_global && !wide) compress_table();     // its scratch aliases the (now dead) pair counts
    if (pool_fail) return 2;
    *tab_len = tl;

    enc_o1_payload<N>(in, n, out, out_end, ptr_out, S.ring, S.rank, symtab, nsym, shift, sym_smem, lane);
    return 0;
}

}  // namespace b200
