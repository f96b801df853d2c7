// rans_decode.cuh -- device side of rans_uncompress_to_4x16: order-0 and order-1
// rANS decode with N = 4 or 32 interleaved states.  One warp per stream.
//
// Reference behaviour restated here (never its code):
//   o0: rANS_static4x16pr.c:234-349, rANS_static32x16pr.c:256-410
//   o1: rANS_static4x16pr.c:524-821, rANS_static32x16pr.c:531-758
//   tables: rANS_static16_int.h:191-272 (alphabet, order-0), :425-536 (order-1)
#pragma once
#include "common.cuh"
#ifdef B200_DEBUG
#include <stdio.h>
#endif

namespace b200 {

// Bump allocator over a scratch pool shared by the jobs of one launch.
struct Pool {
    uint8_t *base;
    unsigned long long cap;
    unsigned long long *used;   // device counter
};
__device__ inline uint8_t *pool_alloc(const Pool &p, uint32_t bytes, int lane) {
    unsigned long long off = 0;
    bytes = (bytes + 255u) & ~255u;
    if (lane == 0) off = atomicAdd(p.used, (unsigned long long)bytes);
    off = __shfl_sync(FULL, off, 0);
    if (off + bytes > p.cap) return nullptr;
    return p.base + off;
}

// ------------------------------------------------------------------------
// Compressed-word staging: the 16-bit renormalisation words of one stream are
// consumed strictly in order by the whole warp.  They are staged through a
// 1 KiB shared-memory ring filled by 16-byte cp.async copies (four 256-byte
// chunks), indexed by the low bits of the global address so that alignment is
// preserved; bytes past the end of the stream are zero-filled, never read.
// ------------------------------------------------------------------------
constexpr uint32_t RING = 1024, CHUNK = 256;      // order-1 decoder: 1 KiB ring

// RS: ring bytes (power of two); refill unit RS/4; `GROUP` steps (64 bytes each at most)
// may be consumed between two advance_group() calls as long as GROUP*64 <= RS/4.
template <uint32_t RS>
struct WordRingT {
    static constexpr uint32_t CH = RS / 4;
    const uint8_t *a0;   // CH-byte aligned global base
    uint8_t *ring;       // shared memory, RS bytes, 16-byte aligned
    uint32_t pos;        // next unread byte, offset from a0
    uint32_t end;        // end of stream, offset from a0
    uint32_t fe;         // ring holds [fe-RS, fe)

    __device__ __forceinline__ void fill_chunk(uint32_t c, int lane) const {
        if ((uint32_t)lane < CH / 16) {
            uint32_t p = c + lane * 16;
            uint32_t nb = p + 16 <= end ? 16u : (p < end ? end - p : 0u);
            cp_async16_zfill(ring + (p & (RS - 1)), a0 + p, nb);
        }
    }
    __device__ __forceinline__ void init(const uint8_t *in, uint32_t start, uint32_t in_size,
                                         uint8_t *ring_, int lane) {
        uintptr_t A = (uintptr_t)in;
        a0 = (const uint8_t *)(A & ~(uintptr_t)(CH - 1));
        uint32_t d = (uint32_t)(A - (uintptr_t)a0);
        ring = ring_;
        pos = d + start;
        end = d + in_size;
        uint32_t c0 = pos & ~(CH - 1);
        __syncwarp();
        for (uint32_t c = 0; c < RS; c += CH) fill_chunk(c0 + c, lane);
        cp_async_commit();
        cp_async_wait_all();
        __syncwarp();
        fe = c0 + RS;
    }
    // Called once per step (a step consumes at most 64 bytes).
    __device__ __forceinline__ void advance(int lane) {
        if (pos + (RS - CH) >= fe) {
            cp_async_wait_all();       // the chunk issued one refill ago
            __syncwarp();
            fill_chunk(fe, lane);
            cp_async_commit();
            fe += CH;
        }
    }
    // Fast-path maintenance, once per group of steps consuming <= CH bytes: keep at least
    // RS-CH bytes ahead; the chunk issued here is not read before the next call, which
    // first waits for it.
    __device__ __forceinline__ void advance_group(int lane) {
        cp_async_wait_all();
        __syncwarp();
        if (pos + (RS - CH) >= fe) {
            fill_chunk(fe, lane);
            cp_async_commit();
            fe += CH;
        }
    }
    __device__ __forceinline__ void advance4(int lane) { advance_group(lane); }
    __device__ __forceinline__ uint32_t word_at(uint32_t p) const {
        // p may be odd when the stream sits at an odd address
        uint32_t o = p & (RS - 1);
        if (p & 1) return ring[o] | ((uint32_t)ring[(o + 1) & (RS - 1)] << 8);
        return *(const uint16_t *)(ring + o);
    }
    __device__ __forceinline__ void drain() const { cp_async_wait_all(); __syncwarp(); }
};
typedef WordRingT<RING> WordRing;                 // order-1 decoder
constexpr uint32_t RING0 = 2048;                  // order-0 decoder: 8 steps per refill check
typedef WordRingT<RING0> WordRing0;

__device__ __forceinline__ void stg_u8(uint8_t *p, uint32_t v) {
    asm volatile("st.global.u8 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(const uint8_t *base, uint32_t off) {
    uint32_t v, a = (uint32_t)__cvta_generic_to_shared(base) + off;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(const uint8_t *base, uint32_t off) {
    uint32_t v, a = (uint32_t)__cvta_generic_to_shared(base) + off;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

__device__ __forceinline__ uint32_t lds_u8a(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16a(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32a(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void stg_u128(uint8_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(__cvta_generic_to_global(p)), "r"(a), "r"(b),
                 "r"(c), "r"(d) : "memory");
}

// One renormalisation round for the warp (rANS_word.h:414-476): lanes whose
// state fell below 2^15 take the next words in lane order.  A lane refills only
// if two more bytes exist, exactly like RansDecRenormSafe.
template <typename WR>
__device__ __forceinline__ uint32_t renorm_step(uint32_t R, bool act, WR &w, int lane,
                                                uint32_t lt) {
    bool need = act && R < RANS_L;
    uint32_t mask = __ballot_sync(FULL, need);
    uint32_t k = __popc(mask & lt);
    uint32_t cnt = __popc(mask);
    if (w.pos + 64 > w.end) {            // near the end of the stream: count what is left
        uint32_t avail = w.end > w.pos ? (w.end - w.pos) >> 1 : 0;
        need = need && k < avail;
        cnt = min(cnt, avail);
    }
    if (need) R = (R << 16) | w.word_at(w.pos + 2 * k);
    w.pos += 2 * cnt;
    w.advance(lane);
    return R;
}

// ------------------------------------------------------------------------
// Alphabet list (rANS_static16_int.h:191-238).  Single-thread, bounds checked.
// Marks F[sym]=1; returns bytes consumed, 0 on failure.
// ------------------------------------------------------------------------
__device__ inline int get_alphabet(const uint8_t *cp, const uint8_t *end, uint32_t *F) {
    const uint8_t *op = cp;
    if (cp >= end) return 0;
    int run = 0, j = *cp++;
    do {
        F[j] = 1;
        if (cp >= end) return 0;
        if (!run && j + 1 == *cp) {
            if (cp + 1 >= end) return 0;
            j = *cp++;
            run = *cp++;
        } else if (run) {
            run--;
            if (++j > 255) return 0;
        } else {
            j = *cp++;
        }
    } while (j && cp < end);
    return (int)(cp - op);
}

// ======================================================================== o0
// 8 KiB per stream.  In dec_kernel<false> the block is 8 KiB aligned in the shared window, so
// the slot look-up address is (state & 4095) | lut and the ring address (pos & 2047) | ring:
// one LOP3 each.  freq and start are split into two 16-bit arrays: two independent loads
// instead of a load plus two bit-field extractions on the state chain.
struct __align__(16) DecO0Smem {
    uint8_t  lut[4096];    // slot -> symbol
    uint16_t f16[256];     // freq (may be 4096: no 12-bit wrap)
    uint16_t b16[256];     // start
    uint32_t tab[256];     // parse scratch: raw counts
    uint8_t  ring[RING0];
};
static_assert(sizeof(DecO0Smem) == 8192, "DecO0Smem layout");

// WIDE (N == 32, `out` 16-byte aligned): the 8 x 32 symbols of a group are collected in shared memory (the
// parse scratch is free by now) and leave as sixteen 16-byte stores.  A template parameter, not a flag: as a
// flag both stores were issued (one of them predicated off) on every step.
template <int N, bool ODD, bool AL, bool WIDE>
__device__ __forceinline__ void dec_o0_fast(uint32_t &R_, uint32_t &i_, uint32_t full, uint8_t *out,
                                            WordRing0 &w, DecO0Smem &S, int lane, uint32_t lt) {
    const bool act = (N == 32) ? true : lane < N;
    uint32_t R = R_, i = i_, pos = w.pos;
    const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(S.lut);
    const uint32_t f_s = (uint32_t)__cvta_generic_to_shared(S.f16);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(w.ring);
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(S.tab);
    uint8_t *o = out + i + (act ? lane : 0);
    while (i + 8 * N <= full && pos + 8 * 64 <= w.end) {
        w.pos = pos;
        w.advance_group(lane);          // also orders the previous group's tile reads before new writes
#pragma unroll
        for (int u = 0; u < 8; u++) {
            uint32_t m = R & 4095;
            uint32_t s = lds_u8a(AL ? (lut_s | m) : (lut_s + m));
            uint32_t fa = f_s + 2 * s;
            uint32_t f = lds_u16a(fa), b = lds_u16a(fa + 512);
            R = f * (R >> 12) + m - b;
            if (WIDE) asm volatile("st.shared.u8 [%0], %1;" ::"r"(tile_s + u * 32 + lane), "r"(s) : "memory");
            else if (act) stg_u8(o + u * N, s);
            if (N == 32 && !ODD) {
                // branch-free refill (rANS_word.h:414-476): every lane reads a word (lanes that do not need one
                // read a valid but unused position), the lanes below 2^15 take theirs.  In PTX so that the
                // position advances by one three-input add per step and one predicate serves ballot and refill.
                static_assert(RING0 == 2048, "the mask below");
                if (AL) asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 m, t, a, v;\n\t"
                                     "setp.lt.u32 p, %0, 0x8000;\n\t"
                                     "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
                                     "and.b32 t, m, %2;\n\t"
                                     "popc.b32 t, t;\n\t"
                                     "add.u32 a, %1, t;\n\t"
                                     "add.u32 a, a, t;\n\t"
                                     "and.b32 a, a, 2046;\n\t"
                                     "or.b32 a, a, %3;\n\t"
                                     "ld.shared.u16 v, [a];\n\t"
                                     "@p mad.lo.u32 %0, %0, 65536, v;\n\t"
                                     "popc.b32 t, m;\n\t"
                                     "add.u32 %1, %1, t;\n\t"
                                     "add.u32 %1, %1, t;\n\t}"
                                     : "+r"(R), "+r"(pos) : "r"(lt), "r"(ring_s) : "memory");
                else asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 m, t, a, v;\n\t"
                                  "setp.lt.u32 p, %0, 0x8000;\n\t"
                                  "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
                                  "and.b32 t, m, %2;\n\t"
                                  "popc.b32 t, t;\n\t"
                                  "add.u32 a, %1, t;\n\t"
                                  "add.u32 a, a, t;\n\t"
                                  "and.b32 a, a, 2046;\n\t"
                                  "add.u32 a, a, %3;\n\t"
                                  "ld.shared.u16 v, [a];\n\t"
                                  "@p mad.lo.u32 %0, %0, 65536, v;\n\t"
                                  "popc.b32 t, m;\n\t"
                                  "add.u32 %1, %1, t;\n\t"
                                  "add.u32 %1, %1, t;\n\t}"
                                  : "+r"(R), "+r"(pos) : "r"(lt), "r"(ring_s) : "memory");
            } else {
                bool need = act && R < RANS_L;
                uint32_t mask = __ballot_sync(FULL, need);
                uint32_t p = pos + 2 * __popc(mask & lt);
                uint32_t wv;
                if (ODD) wv = lds_u8a(ring_s + (p & (RING0 - 1))) | (lds_u8a(ring_s + ((p + 1) & (RING0 - 1))) << 8);
                else wv = lds_u16a(AL ? (ring_s | (p & (RING0 - 1))) : (ring_s + (p & (RING0 - 1))));
                R = need ? ((R << 16) | wv) : R;
                pos += 2 * __popc(mask);
            }
        }
        if (WIDE) {
            __syncwarp();
            if (lane < 16) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(tile_s + 16 * lane));
                stg_u128(out + i + 16 * lane, v.x, v.y, v.z, v.w);
            }
        }
        o += 8 * N;
        i += 8 * N;
    }
    w.pos = pos;
    // hand over to the careful loop with the ring in its per-step regime
    cp_async_wait_all();
    __syncwarp();
    R_ = R;
    i_ = i;
}

// Returns 0 on success.  `out` receives out_sz bytes.
template <int N>
__device__ int dec_o0(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t out_sz,
                      DecO0Smem &S, int lane) {
    if (in_size < 16 || out_sz >= 0x7fffffffu) return 1;
    const uint8_t *end = in + in_size;
    for (int j = lane; j < 256; j += 32) S.tab[j] = 0;
    __syncwarp();

    // --- frequency table: alphabet + one varint per listed symbol (lane 0)
    int hdr = 0;
    if (lane == 0) {
        const uint8_t *tend = (N == 4) ? end - 8 : end;     // rANS_static4x16pr.c:249
        const uint8_t *cp = in;
        int n = get_alphabet(cp, tend, S.tab);
        if (n) {
            cp += n;
            for (int j = 0; j < 256; j++) {
                if (!S.tab[j]) continue;
                uint32_t f;
                cp += var_get_u32(cp, tend, &f);
                S.tab[j] = f;
            }
            hdr = (int)(cp - in);
        }
    }
    hdr = __shfl_sync(FULL, hdr, 0);
    if (!hdr) return 1;
    __syncwarp();

    // --- scale to 4096 (rANS_static16_int.h:151-162) and accumulate starts
    uint32_t f[8], loc = 0;
    bool bad = false;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        f[t] = S.tab[lane * 8 + t];
        bad |= f[t] > 4096;
        loc += f[t];
    }
    if (__any_sync(FULL, bad)) return 1;
    uint32_t incl = warp_incl_scan(loc, lane);
    uint32_t fsum = __shfl_sync(FULL, incl, 31);
    if (fsum == 0 || fsum > 4096 || (fsum & (fsum - 1))) return 1;
    int sh = __clz(fsum) - __clz(4096u);
    uint32_t x = (incl - loc) << sh;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        uint32_t ff = f[t] << sh;
        S.f16[lane * 8 + t] = (uint16_t)ff;
        S.b16[lane * 8 + t] = (uint16_t)x;
        x += ff;
    }
    __syncwarp();
    for (int j = 0; j < 256; j++) {
        uint32_t ff = S.f16[j], b = S.b16[j];
        for (uint32_t y = lane; y < ff; y += 32) S.lut[b + y] = (uint8_t)j;
    }

    // --- N initial states, little endian, lane 0 first (rANS_word.h:121-134)
    if ((uint32_t)(end - (in + hdr)) < (uint32_t)N * 4) return 1;
    const bool act = lane < N;
    uint32_t R = RANS_L;
    if (act) {
        const uint8_t *p = in + hdr + 4 * lane;
        R = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
    }
    if (__any_sync(FULL, R < RANS_L)) return 1;

    WordRing0 w;
    w.init(in, hdr + 4 * N, in_size, S.ring, lane);     // ends with __syncwarp: lut visible
    const uint32_t lt = lanemask_lt();
    const uint32_t full = out_sz - out_sz % N;
    uint32_t i = 0;
    // Hot loop: groups of 8 steps while 8*64 bytes of words are certainly left, so no
    // end-of-stream bookkeeping and one ring check per group.
    const bool al = (((uint32_t)__cvta_generic_to_shared(S.lut)) & 8191) == 0;
    const bool wide = N == 32 && ((uintptr_t)out & 15) == 0;
    if (al && wide) {                                                   // the common case
        if (w.pos & 1) dec_o0_fast<N, true, true, N == 32>(R, i, full, out, w, S, lane, lt);
        else dec_o0_fast<N, false, true, N == 32>(R, i, full, out, w, S, lane, lt);
    } else if (al) {
        if (w.pos & 1) dec_o0_fast<N, true, true, false>(R, i, full, out, w, S, lane, lt);
        else dec_o0_fast<N, false, true, false>(R, i, full, out, w, S, lane, lt);
    } else {
        if (w.pos & 1) dec_o0_fast<N, true, false, false>(R, i, full, out, w, S, lane, lt);
        else dec_o0_fast<N, false, false, false>(R, i, full, out, w, S, lane, lt);
    }
    for (; i < full; i += N) {
        uint32_t m = R & 4095;
        uint32_t s = S.lut[m];
        R = (uint32_t)S.f16[s] * (R >> 12) + m - S.b16[s];
        if (act) out[i + lane] = (uint8_t)s;
        R = renorm_step(R, act, w, lane, lt);
    }
    // the last out_sz % N symbols: table look-up only (rANS_static32x16pr.c:400-401)
    if ((uint32_t)lane < out_sz - full) out[full + lane] = S.lut[R & 4095];
    w.drain();
    return 0;
}

// ======================================================================== o1
// Tables live in rank space: only symbols of the alphabet get rows/columns.
//   cum[ctx][r]  = first slot of the r-th listed symbol in context ctx (r = 0..nsym; 16 bit,
//                  the last entry is the total), so freq = cum[r+1]-cum[r]
//   blut[ctx][b] = rank of the symbol owning slot b<<(shift-6)
// A look-up is blut -> cum[r], cum[r+1], stepping r forward while slot >= cum[r+1].
struct DecO1Tabs {
    uint16_t *cum;      // [nsym][nsym+1]
    uint8_t  *blut;     // [nsym][1 << bb]
    uint8_t  *sym;      // [nsym] rank -> symbol
    uint32_t  nsym;
};
// Buckets per context: 2^bb with bb in {6,7,8}; the finest that fits is used, so that
// a look-up rarely has to step over more than one symbol (the scan is paid by the
// whole warp for its slowest lane).
__host__ __device__ inline uint32_t dec_o1_tab_bytes(uint32_t nsym, uint32_t bb) {
    return ((nsym * (nsym + 1) * 2 + 15) & ~15u) + (nsym << bb) + ((nsym + 15) & ~15u);
}

// one order-1 row (rANS_static16_int.h:425-456), single thread.  A[] lists the
// alphabet (rank -> symbol); writes F[rank].  Returns bytes consumed, 0 = error.
__device__ inline int get_freq_row(const uint8_t *cp, const uint8_t *end, uint32_t nsym,
                                   uint32_t *F, uint32_t *tot) {
    const uint8_t *op = cp;
    if (cp >= end) return 0;
    uint32_t zrun = 0, t = 0;
    for (uint32_t r = 0; r < nsym; r++) {
        uint32_t f = 0;
        if (cp >= end) { F[r] = 0; continue; }    // the reference's loop stops here; rest stay 0
        if (zrun) { zrun--; }
        else {
            cp += var_get_u32(cp, end, &f);
            if (f == 0) {
                if (cp >= end) return 0;
                zrun = *cp++;
            }
        }
        F[r] = f;
        t += f;
    }
    *tot = t;
    return (int)(cp - op);
}

// Order-1 hot loop (32 lanes, tables in shared memory, 16-byte aligned lane
// segments): 16 steps per iteration, each lane collecting its 16 output bytes in
// registers and writing them with one 128-bit store; no end-of-stream checks.
template <bool ODD>
__device__ __forceinline__ void dec_o1_fast(uint32_t &R_, uint32_t &ctx_, uint32_t &k_, uint32_t seg,
                                            uint8_t *o, WordRing &w, uint32_t fs_s, uint32_t blut_s,
                                            uint32_t sym_s, uint32_t ns, uint32_t shift, uint32_t bb,
                                            int lane, uint32_t lt) {
    uint32_t R = R_, ctx = ctx_, k = k_, pos = w.pos;
    const uint32_t mask = (1u << shift) - 1, bw = shift - bb, rs2 = (ns + 1) * 2;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(w.ring);
    while (k + 16 <= seg && pos + 16 * 64 <= w.end) {
        uint32_t acc[4] = {0, 0, 0, 0};
#pragma unroll
        for (int g = 0; g < 4; g++) {
            w.pos = pos;
            w.advance4(lane);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                uint32_t m = R & mask;
                uint32_t r = lds_u8a(blut_s + (ctx << bb) + (m >> bw));
                uint32_t ea = fs_s + ctx * rs2 + r * 2;
                uint32_t c0 = lds_u16a(ea), c1 = lds_u16a(ea + 2);
                while (m >= c1 && r + 1 < ns) { r++; ea += 2; c0 = c1; c1 = lds_u16a(ea + 2); }
                R = (c1 - c0) * (R >> shift) + m - c0;
                ctx = r;
                acc[g] |= lds_u8a(sym_s + r) << (8 * u);
                if (!ODD) {
                    // branch-free refill, as in dec_o0_fast: every lane reads a word, the lanes below 2^15 take theirs
                    static_assert(RING == 1024, "the mask below");
                    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 m, t, a, v;\n\t"
                                 "setp.lt.u32 p, %0, 0x8000;\n\t"
                                 "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
                                 "and.b32 t, m, %2;\n\t"
                                 "popc.b32 t, t;\n\t"
                                 "add.u32 a, %1, t;\n\t"
                                 "add.u32 a, a, t;\n\t"
                                 "and.b32 a, a, 1022;\n\t"
                                 "add.u32 a, a, %3;\n\t"
                                 "ld.shared.u16 v, [a];\n\t"
                                 "@p mad.lo.u32 %0, %0, 65536, v;\n\t"
                                 "popc.b32 t, m;\n\t"
                                 "add.u32 %1, %1, t;\n\t"
                                 "add.u32 %1, %1, t;\n\t}"
                                 : "+r"(R), "+r"(pos) : "r"(lt), "r"(ring_s) : "memory");
                } else {
                    bool need = R < RANS_L;
                    uint32_t bal = __ballot_sync(FULL, need);
                    if (need) {
                        uint32_t p = pos + 2 * __popc(bal & lt);
                        uint32_t wv = lds_u8a(ring_s + (p & (RING - 1))) | (lds_u8a(ring_s + ((p + 1) & (RING - 1))) << 8);
                        R = (R << 16) | wv;
                    }
                    pos += 2 * __popc(bal);
                }
            }
        }
        stg_u128(o + k, acc[0], acc[1], acc[2], acc[3]);
        k += 16;
    }
    w.pos = pos;
    cp_async_wait_all();
    __syncwarp();
    R_ = R; ctx_ = ctx; k_ = k;
}

// Large alphabets (nsym > 64): tables in global memory.  With thousands of streams and up to
// 64 Ki (context, symbol) pairs each, a look-up is a DRAM access, so the layout makes it ONE
// 32-byte sector and no dependent second load:
//   entry        = start << 20 | (freq - 1) << 8 | rank      (start < 4096, freq <= 4096)
//   rec[ctx][b]  = 8 words for bucket b of 64 (slot >> (shift - 6)): the entry owning the
//                  bucket's first slot and the next six that start inside the bucket (padded by
//                  repeating the last one), then an overflow word: 0xffffffff, or
//                  start << 20 | k of the eighth such entry (rare; continues in ent)
//   ent[ctx][k]  = the context's entries in slot order (only symbols with a non-zero
//                  frequency), closed by 0xffffffff
// The symbol is the entry with the largest start <= slot: a max over seven selects.
__device__ __forceinline__ uint32_t ldg_u32d(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(__cvta_generic_to_global(p)));
    return v;
}
struct DecO1Big {
    const uint32_t *rec;    // [nsym][64][8]
    const uint32_t *ent;    // [nsym][nsym+1]
    uint32_t ns1, shift;
    // returns the entry for slot m in context ctx
    __device__ __forceinline__ uint32_t look(uint32_t m, uint32_t ctx) const {
        const uint32_t *r = rec + (((ctx << 6) + (m >> (shift - 6))) << 3);
        uint4 a, c;                                          // one 32-byte load: one request for the record's sector
#ifndef B200_LOOK_LD
#define B200_LOOK_LD "ld.global.v8.u32"
#endif
        asm volatile(B200_LOOK_LD " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w)
                     : "l"(__cvta_generic_to_global(r)));
        const uint32_t T = (m << 20) | 0xfffff;              // entries with start <= m are <= T
        uint32_t best = a.x;                                 // the bucket's first entry always qualifies
        best = max(best, a.y <= T ? a.y : 0u);
        best = max(best, a.z <= T ? a.z : 0u);
        best = max(best, a.w <= T ? a.w : 0u);
        best = max(best, c.x <= T ? c.x : 0u);
        best = max(best, c.y <= T ? c.y : 0u);
        best = max(best, c.z <= T ? c.z : 0u);
        if (c.w <= T && c.w != 0xffffffffu) {                // more than seven entries reach into the bucket
            const uint32_t *e = ent + ctx * ns1 + (c.w & 0xfffff);
            best = ldg_u32d(e);
            for (;;) {
                const uint32_t nx = ldg_u32d(++e);
                if (nx == 0xffffffffu || nx > T) break;
                best = nx;
            }
        }
        return best;
    }
};
template <bool ODD>
__device__ __forceinline__ void dec_o1_fast_big(uint32_t &R_, uint32_t &ctx_, uint32_t &k_, uint32_t seg,
                                                uint8_t *o, WordRing &w, const DecO1Big &T, uint32_t sym_s,
                                                int lane, uint32_t lt) {
    uint32_t R = R_, ctx = ctx_, k = k_, pos = w.pos;
    const uint32_t mask = (1u << T.shift) - 1;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(w.ring);
    while (k + 16 <= seg && pos + 16 * 64 <= w.end) {
        uint32_t acc[4] = {0, 0, 0, 0};
#pragma unroll
        for (int g = 0; g < 4; g++) {
            w.pos = pos;
            w.advance4(lane);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t m = R & mask;
                const uint32_t e = T.look(m, ctx);
                const uint32_t r = e & 0xff;
                R = (((e >> 8) & 0xfff) + 1) * (R >> T.shift) + m - (e >> 20);
                ctx = r;
                acc[g] |= lds_u8a(sym_s + r) << (8 * u);
                if (!ODD) {
                    // branch-free refill, as in dec_o0_fast: every lane reads a word, the lanes below 2^15 take theirs
                    static_assert(RING == 1024, "the mask below");
                    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 m, t, a, v;\n\t"
                                 "setp.lt.u32 p, %0, 0x8000;\n\t"
                                 "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
                                 "and.b32 t, m, %2;\n\t"
                                 "popc.b32 t, t;\n\t"
                                 "add.u32 a, %1, t;\n\t"
                                 "add.u32 a, a, t;\n\t"
                                 "and.b32 a, a, 1022;\n\t"
                                 "add.u32 a, a, %3;\n\t"
                                 "ld.shared.u16 v, [a];\n\t"
                                 "@p mad.lo.u32 %0, %0, 65536, v;\n\t"
                                 "popc.b32 t, m;\n\t"
                                 "add.u32 %1, %1, t;\n\t"
                                 "add.u32 %1, %1, t;\n\t}"
                                 : "+r"(R), "+r"(pos) : "r"(lt), "r"(ring_s) : "memory");
                } else {
                    bool need = R < RANS_L;
                    uint32_t bal = __ballot_sync(FULL, need);
                    if (need) {
                        uint32_t p = pos + 2 * __popc(bal & lt);
                        uint32_t wv = lds_u8a(ring_s + (p & (RING - 1))) | (lds_u8a(ring_s + ((p + 1) & (RING - 1))) << 8);
                        R = (R << 16) | wv;
                    }
                    pos += 2 * __popc(bal);
                }
            }
        }
        stg_u128(o + k, acc[0], acc[1], acc[2], acc[3]);
        k += 16;
    }
    w.pos = pos;
    cp_async_wait_all();
    __syncwarp();
    R_ = R; ctx_ = ctx; k_ = k;
}

// ------------------------------------------------------------------------
// Order-1 table rows (rANS_static16_int.h:425-456, 488-530) parsed by the whole warp.
// The byte stream is a sequence of tokens -- a varint count, or 0x00 followed by a raw
// byte z meaning z further zero counts -- with no delimiters between rows.  Which byte
// starts a token is decided by a three-state machine (token start / inside a varint /
// raw run byte); its per-byte transition functions are composed with a warp scan, 32
// bytes per round, so token starts, their slot numbers (exclusive scan of 1 or 1+z) and
// hence (row, column) are known without walking the stream serially.
// Raw counts are written to cum[row*ns1 + col + 1]; rows must be zeroed beforehand.
// Returns 0 and the end of the table in *end_out, or 1 on malformed input.
// ------------------------------------------------------------------------
__device__ inline int parse_o1_rows(const uint8_t *cp, const uint8_t *tend, uint32_t nsym, uint32_t tot,
                                    uint16_t *cum, uint32_t ns1, int lane, const uint8_t **end_out) {
    const uint32_t total = nsym * nsym;
    const uint32_t avail = (uint32_t)(tend - cp);
    const uint32_t lt = lanemask_lt();
    uint32_t state = 0, slot_base = 0, end_off = 0;     // 0 start, 1 varint, 2 run byte
    int err = 0;
    for (uint32_t base = 0; slot_base < total; base += 32) {
        if (base >= avail) return 1;                    // ran out of bytes before the last row
        const uint32_t i = base + lane;
        const bool in = i < avail;
        const uint32_t b = in ? cp[i] : 0x01;           // padding: harmless one-byte tokens
        // transition function of this byte, 2 bits per source state
        uint32_t f = (b == 0 ? 2u : (b < 128 ? 0u : 1u)) | ((b < 128 ? 0u : 1u) << 2) | (0u << 4);
        // inclusive scan: F_i = f_i o ... o f_0
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t g = __shfl_up_sync(FULL, f, o);    // earlier bytes: applied first
            if (lane >= o) {
                uint32_t a0 = (g >> 0) & 3, a1 = (g >> 2) & 3, a2 = (g >> 4) & 3;
                f = ((f >> (2 * a0)) & 3) | (((f >> (2 * a1)) & 3) << 2) | (((f >> (2 * a2)) & 3) << 4);
            }
        }
        uint32_t fprev = __shfl_up_sync(FULL, f, 1);
        uint32_t before = lane ? ((fprev >> (2 * state)) & 3) : state;
        const bool start = in && before == 0;
        // token at a start byte
        uint32_t val = 0, slots = 0, len = 0;
        bool terr = false;                              // only counts if the token belongs to the table
        if (start) {
            if (b == 0) {
                if (i + 1 >= avail) { terr = true; slots = 1; len = 1; }
                else { slots = 1 + cp[i + 1]; len = 2; }
            } else {
                uint32_t c = b;
                val = c & 0x7f; len = 1; slots = 1;
                while ((c & 0x80) && i + len < avail && len < 6) { c = cp[i + len]; val = (val << 7) | (c & 0x7f); len++; }
                terr = (c & 0x80) || val == 0 || val > tot;     // unterminated, non-canonical zero, too large
            }
        }
        uint32_t incl = warp_incl_scan(slots, lane);
        uint32_t s0 = slot_base + incl - slots;
        if (start && s0 < total) {
            uint32_t row = s0 / nsym, col = s0 - row * nsym;
            if (terr || col + slots > nsym) err = 1;    // a zero run never crosses a row
            else if (val) cum[row * ns1 + col + 1] = (uint16_t)val;
            if (s0 + slots >= total) end_off = i + len; // the token that completes the last row
        }
        uint32_t chunk_slots = __shfl_sync(FULL, incl, 31);
        state = (__shfl_sync(FULL, f, 31) >> (2 * state)) & 3;
        slot_base += chunk_slots;
        if (__any_sync(FULL, err)) return 1;
        (void)lt;
    }
    // exactly one lane saw the closing token
    uint32_t m = __ballot_sync(FULL, end_off != 0);
    if (!m) return 1;
    end_off = __shfl_sync(FULL, end_off, __ffs(m) - 1);
    *end_out = cp + end_off;
    return 0;
}

struct __align__(16) DecO1Smem {
    union {                     // the alphabet marks are dead before the word ring is filled
        uint8_t  ring[RING];
        uint32_t F0[256];
    };
    uint8_t  rank[256];
};                              // followed by dynamic table storage (DecO1Tabs) when it fits

template <int N>
__device__ int dec_o1(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t out_sz,
                      DecO1Smem &S, uint8_t *smem_tabs, uint32_t smem_tab_bytes, DecO0Smem *o0s,
                      const Pool &pool, int lane) {
    if (in_size < (uint32_t)(N == 4 ? 16 : N * 4) || out_sz >= 0x7fffffffu) return 1;
    const uint8_t *end = in + in_size;
    const uint8_t *cp = in, *tend = end, *after = nullptr;
    const uint32_t shift = *cp >> 4;
    if (shift != 10 && shift != 12) return 1;       // the encoder writes nothing else
    const uint32_t tot = 1u << shift;
    const bool comp = *cp++ & 1;
    if (comp) {                                     // table itself rANS-coded (o0, 4 lanes)
        uint32_t usz = 0, csz = 0;
        cp += var_get_u32(cp, end, &usz);
        cp += var_get_u32(cp, end, &csz);
        if (csz > (uint32_t)(end - cp) || usz > 257 * 257 * 3) return 1;
        after = cp + csz;
        uint8_t *tb = pool_alloc(pool, usz + 16, lane);
        if (!tb) return 2;
        if (dec_o0<4>(cp, csz, tb, usz, *o0s, lane)) return 1;
        __threadfence_block();
        __syncwarp();
        cp = tb;
        tend = tb + usz;
    }

    // --- alphabet (single thread), then ranks
    for (int j = lane; j < 256; j += 32) S.F0[j] = 0;
    __syncwarp();
    int n = 0;
    if (lane == 0) n = get_alphabet(cp, tend, S.F0);
    n = __shfl_sync(FULL, n, 0);
    if (!n) return 1;
    cp += n;
    if (cp >= tend) return 1;
    __syncwarp();
    uint32_t nsym = 0;
    {
        uint32_t loc = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) loc += S.F0[lane * 8 + t] ? 1 : 0;
        uint32_t incl = warp_incl_scan(loc, lane);
        nsym = __shfl_sync(FULL, incl, 31);
        uint32_t r = incl - loc;
        for (int t = 0; t < 8; t++) {
            int j = lane * 8 + t;
            S.rank[j] = S.F0[j] ? (uint8_t)r++ : 0xff;
        }
    }
    __syncwarp();

    // --- table storage: shared memory when it fits, else the scratch pool
    DecO1Tabs T;
    T.nsym = nsym;
    const bool big = nsym > 64;
    const uint32_t ns1 = nsym + 1;
    uint32_t bb = 8;
    while (bb > 6 && dec_o1_tab_bytes(nsym, bb) > smem_tab_bytes) bb--;
    const uint32_t ent_bytes = (nsym * ns1 * 4 + 31) & ~31u;      // records start on a sector boundary
    uint32_t need = big ? ent_bytes + (nsym << 11) + 256 : dec_o1_tab_bytes(nsym, bb);
    uint8_t *tb;
    const bool in_smem = !big && need <= smem_tab_bytes;
    if (in_smem) tb = smem_tabs;
    else { tb = pool_alloc(pool, need, lane); if (!tb) return 2; }
    uint32_t *ent = (uint32_t *)tb;               // big: compact rows (DecO1Big)
    uint32_t craw_stride = ns1;                   // raw counts: row stride in 16-bit units
    if (big) {
        // the raw 16-bit counts of row i are parsed into the upper half of ent row i and are
        // in registers before the compact row overwrites them
        T.cum = (uint16_t *)tb + ns1;
        craw_stride = 2 * ns1;
        T.blut = tb + ent_bytes;                      // big: the bucket records (DecO1Big::rec)
        T.sym = T.blut + (nsym << 11);
    } else {
        T.cum = (uint16_t *)tb;
        T.blut = tb + ((nsym * ns1 * 2 + 15) & ~15u);
        T.sym = T.blut + (nsym << bb);
    }
    for (int j = lane; j < 256; j += 32)          // presence from F0: rank 255 is a valid rank
        if (S.F0[j]) T.sym[S.rank[j]] = (uint8_t)j;

    // --- rows, in alphabet order: raw counts by parse_o1_rows, then scaling, cumulative
    // starts and validation one lane per row
    int err = 0;
    if (big) { for (uint32_t j = lane; j < nsym * ns1; j += 32) ent[j] = 0; }
    else for (uint32_t j = lane; j < nsym * ns1; j += 32) T.cum[j] = 0;
    __syncwarp();
    if (cp >= tend) return 1;
    {
        const uint8_t *table_end = nullptr;
        if (parse_o1_rows(cp, tend, nsym, tot, T.cum, craw_stride, lane, &table_end)) return 1;
        cp = table_end;
    }
    __syncwarp();
    const uint32_t bw = shift - bb, nb = 1u << bb;
    if (nsym <= 64) {
        // small alphabets (tables in shared memory): one lane per row
        for (uint32_t i0 = 0; i0 < nsym; i0 += 32) {
            const uint32_t i = i0 + lane;
            if (i < nsym) {
                uint16_t *row = T.cum + i * ns1;
                uint32_t tsum = 0;
                for (uint32_t r = 0; r < nsym; r++) tsum += row[r + 1];
                int sh = 0;
                if (tsum) { uint32_t z = tsum; while (z < tot) { z *= 2; sh++; } }   // normalise_freq_shift
                uint32_t x = 0;
                for (uint32_t r = 0; r < nsym; r++) {          // in place: raw count of r sits at [r+1]
                    uint32_t f = (uint32_t)row[r + 1] << sh;
                    if (f > tot - x) { err = 1; break; }
                    row[r] = (uint16_t)x;
                    x += f;
                }
                row[nsym] = (uint16_t)x;
                if (!err && tsum && x != tot) err = 1;
                // bucket index: rank of the symbol owning the first slot of each bucket
                uint8_t *bl = T.blut + (i << bb);
                uint32_t r = 0;
                for (uint32_t bk = 0; bk < nb && !err; bk++) {
                    uint32_t m = bk << bw;
                    while (r + 1 < nsym && m >= row[r + 1]) r++;
                    bl[bk] = (uint8_t)r;
                }
            }
        }
    } else {
        // large alphabets (tables in the L2-resident pool): the warp walks the rows together,
        // lanes over columns, so that global accesses coalesce
        // (lane l holds columns 8l..8l+7: one prefix over the lane's eight and one warp scan)
        const uint32_t bwb = shift - 6, B = 1u << bwb;        // 64 buckets of B slots
        uint32_t *cnt = (uint32_t *)smem_tabs;                // scratch: the tables are in the pool
        uint32_t *crow = cnt + 64;                            // the row's entries, <= 257 words
        uint32_t *recs = (uint32_t *)T.blut;
        for (uint32_t i = 0; i < nsym; i++) {
            const uint16_t *raw = T.cum + i * craw_stride;
            uint32_t *row = ent + i * ns1;
            uint32_t f[8], tsum = 0, nz = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                uint32_t c = lane * 8 + t;
                f[t] = c < nsym ? raw[c + 1] : 0;
                tsum += f[t];
                nz += f[t] ? 1 : 0;
            }
            if (warp_sum(tsum) > tot) { err = 1; continue; }            // (uniform) keeps the packed scan exact
            uint32_t pk = warp_incl_scan(tsum | (nz << 16), lane);      // sum <= 4096, <= 256 non-zero
            const uint32_t all = __shfl_sync(FULL, pk, 31);
            const uint32_t rsum = all & 0xffff, nnz = all >> 16;
            pk -= tsum | (nz << 16);
            int sh = 0;
            if (rsum) { uint32_t z = rsum; while (z < tot) { z *= 2; sh++; } }
            if (rsum && (rsum << sh) != tot) { err = 1; continue; }     // (uniform)
            cnt[lane] = 0; cnt[lane + 32] = 0;
            __syncwarp();                                // every lane has read its raw counts
            uint32_t x = (pk & 0xffff) << sh, k = pk >> 16;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const uint32_t ff = f[t] << sh;
                if (ff) {
                    const uint32_t e = (x << 20) | ((ff - 1) << 8) | (lane * 8 + t);
                    crow[k] = e;
                    row[k++] = e;
                    // an entry belongs to the bucket records from ceil(start / B) on
                    const uint32_t fb = (x + B - 1) >> bwb;
                    if (fb < 64) atomicAdd(&cnt[fb], 1u);
                }
                x += ff;
            }
            if (lane == 0) row[nnz] = 0xffffffffu;       // closes the row (an empty row is only this)
            __syncwarp();
            // bucket b's first entry: k(b) = #{entries with ceil(start / B) <= b} - 1
            const uint32_t c0 = cnt[2 * lane], c1 = cnt[2 * lane + 1];
            const uint32_t upto = warp_incl_scan(c0 + c1, lane);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t b = 2 * lane + h;
                const uint32_t kin = h ? upto : upto - c1;              // entries starting at or before b*B
                const uint32_t bend = (b + 1) << bwb;
                uint32_t wv[8];
                uint32_t lastv = 0;
#pragma unroll
                for (int q = 0; q < 7; q++) {
                    const uint32_t kk = kin - 1 + q;
                    uint32_t e = lastv;
                    if (kin && kk < nnz) {
                        const uint32_t ce = crow[kk];
                        if (q == 0 || (ce >> 20) < bend) e = ce;
                    }
                    wv[q] = e;
                    lastv = e;
                }
                wv[7] = 0xffffffffu;
                if (kin && kin + 6 < nnz) {
                    const uint32_t ce = crow[kin + 6];
                    if ((ce >> 20) < bend) wv[7] = (ce & 0xfff00000u) | (kin + 6);
                }
                uint4 *dst = (uint4 *)(recs + ((size_t)((i << 6) + b) << 3));
                dst[0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                dst[1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
            }
            __syncwarp();
        }
    }
    err = __any_sync(FULL, err);
    if (err) return 1;
    __threadfence_block();
    __syncwarp();
    if (after) cp = after;
    if ((uint32_t)(end - cp) < (uint32_t)N * 4) return 1;
    const bool act = lane < N;
    uint32_t R = RANS_L;
    if (act) {
        const uint8_t *p = cp + 4 * lane;
        R = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
    }
    if (__any_sync(FULL, R < RANS_L)) return 1;

#ifdef B200_DEBUG
    if (lane == 0) {
        printf("dec_o1<%d> in_size=%u out_sz=%u shift=%u comp=%d nsym=%u in_smem=%d cpoff=%ld R0=%x\n", N, in_size, out_sz, shift, (int)comp, nsym, (int)in_smem, (long)(cp - in), R);
        for (uint32_t i = 0; i < nsym && i < 4; i++) { printf(" row%u:", i); for (uint32_t r = 0; r < nsym && r < 8; r++) printf(" %u", T.cum[i*ns1+r]); printf(" | blut"); for (int b = 0; b < 8; b++) printf(" %u", T.blut[i*64+b*8]); printf(" sym %u\n", T.sym[i]); }
    }
#endif
    WordRing w;
    if (!S.F0[0]) return 1;                 // symbol 0 is always listed by the encoder
    __syncwarp();
    w.init(in, (uint32_t)(cp - in) + 4 * N, in_size, S.ring, lane);      // overwrites F0
    __threadfence_block();
    __syncwarp();
    const uint32_t lt = lanemask_lt();
    const uint32_t seg = out_sz / N, mask = tot - 1;
    uint8_t *o = out + (size_t)lane * seg;
    uint32_t ctx = S.rank[0];               // every lane starts in context 0
    const uint16_t *cumt = T.cum;
    const uint8_t *blut = T.blut, *symtab = T.sym;
    const uint32_t ns = nsym;

    DecO1Big B{(const uint32_t *)T.blut, ent, ns1, shift};
    if (big) {                                // rank -> symbol moves to shared memory (ranks are dead)
        uint32_t v[8];
#pragma unroll
        for (int t = 0; t < 8; t++) v[t] = (uint32_t)lane + 32 * t < ns ? symtab[lane + 32 * t] : 0;
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; t++) S.rank[lane + 32 * t] = (uint8_t)v[t];
        __syncwarp();
        symtab = S.rank;
    }
    auto step = [&](bool on) {
        uint32_t m = R & mask, r, c0, c1;
        if (big) {
            const uint32_t e = B.look(m, ctx);
            r = e & 0xff; c0 = e >> 20; c1 = c0 + ((e >> 8) & 0xfff) + 1;
        } else {
            r = blut[(ctx << bb) + (m >> bw)];
            const uint16_t *row = cumt + ctx * (ns + 1);
            c0 = row[r]; c1 = row[r + 1];
            while (m >= c1 && r + 1 < ns) { r++; c0 = c1; c1 = row[r + 1]; }
        }
        if (on) {
            R = (c1 - c0) * (R >> shift) + m - c0;
            ctx = r;
        }
        return (uint8_t)symtab[r];
    };

    uint32_t k = 0;
    if (N == 32 && big && ((((uintptr_t)out) | seg) & 15) == 0) {
        const uint32_t sy_s = (uint32_t)__cvta_generic_to_shared(S.rank);
        if (w.pos & 1) dec_o1_fast_big<true>(R, ctx, k, seg, o, w, B, sy_s, lane, lt);
        else dec_o1_fast_big<false>(R, ctx, k, seg, o, w, B, sy_s, lane, lt);
    } else if (N == 32 && in_smem && ((((uintptr_t)out) | seg) & 15) == 0) {
        const uint32_t fs_s = (uint32_t)__cvta_generic_to_shared(T.cum);
        const uint32_t bl_s = (uint32_t)__cvta_generic_to_shared(T.blut);
        const uint32_t sy_s = (uint32_t)__cvta_generic_to_shared(T.sym);
        __syncwarp();
        if (w.pos & 1) dec_o1_fast<true>(R, ctx, k, seg, o, w, fs_s, bl_s, sy_s, ns, shift, bb, lane, lt);
        else dec_o1_fast<false>(R, ctx, k, seg, o, w, fs_s, bl_s, sy_s, ns, shift, bb, lane, lt);
    }
    for (; k < seg; k++) {
        uint8_t s = step(act);
        if (act) o[k] = s;
        R = renorm_step(R, act, w, lane, lt);
    }
    // remainder: the last lane alone (rANS_static32x16pr.c:676-684)
    const bool last = lane == N - 1;
    for (uint32_t k = seg * N; k < out_sz; k++) {
        uint8_t s = step(last);
        if (last) out[k] = s;
        R = renorm_step(R, last, w, lane, lt);
    }
    w.drain();
    return 0;
}

}  // namespace b200
