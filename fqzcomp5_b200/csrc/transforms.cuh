// transforms.cuh -- PACK and RLE pre/post passes of the rANS Nx16 container,
// warp-cooperative (one warp per stream, like the coders they feed).
//
// Reference behaviour restated here (never its code):
//   pack:   pack.c:56-147 (hts_pack), :161-194 (hts_unpack_meta), :207-344 (hts_unpack)
//   rle:    rle.c:48-98 (symbol choice), :100-138 (encode), :142-189 (decode)
#pragma once
#include "common.cuh"

namespace b200 {

// ------------------------------------------------------------------ PACK
// meta = [nsym][symbols ascending]; codes are ranks; the first symbol of a byte
// sits in the low bits; 8 / 4 / 2 codes per byte for nsym <= 2 / 4 / 16, none
// for a single symbol.  Returns false when more than 16 symbols occur.
__device__ inline bool warp_pack(const uint8_t *in, uint32_t n, uint8_t *meta, uint32_t *meta_len,
                                 uint8_t *out, uint32_t *out_len, uint8_t *smem, int lane) {
    uint8_t *code = smem;                 // 256 bytes: presence, then code numbers
    for (int j = lane; j < 64; j += 32) ((uint32_t *)code)[j] = 0;
    __syncwarp();
    {   // presence flags (benign write races)
        uint32_t head = (uint32_t)((16 - ((uintptr_t)in & 15)) & 15);
        if (head > n) head = n;
        if ((uint32_t)lane < head) code[in[lane]] = 1;
        const uint8_t *p = in + head;
        uint32_t rest = n - head, nv = rest >> 4;
        const uint4 *v = (const uint4 *)p;
        auto mark = [&](const uint4 &q) {
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) code[(w[a] >> (8 * b)) & 0xff] = 1;
        };
        uint32_t i = lane;
        for (; i + 96 < nv; i += 128) {           // four loads in flight per lane
            uint4 q0 = __ldg(v + i), q1 = __ldg(v + i + 32), q2 = __ldg(v + i + 64), q3 = __ldg(v + i + 96);
            mark(q0); mark(q1); mark(q2); mark(q3);
        }
        for (; i < nv; i += 32) mark(__ldg(v + i));
        for (uint32_t i = (nv << 4) + lane; i < rest; i += 32) code[p[i]] = 1;
    }
    __syncwarp();
    uint32_t nsym;
    {
        uint32_t loc = 0;
        uint8_t pr[8];
#pragma unroll
        for (int t = 0; t < 8; t++) { pr[t] = code[lane * 8 + t]; loc += pr[t]; }
        uint32_t incl = warp_incl_scan(loc, lane);
        nsym = __shfl_sync(FULL, incl, 31);
        uint32_t r = incl - loc;
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; t++) {
            int j = lane * 8 + t;
            if (pr[t]) {
                if (r < 256) meta[1 + r] = (uint8_t)j;     // pack.c:65-70 writes every listed symbol
                code[j] = (uint8_t)r++;
            }
        }
        if (lane == 0) meta[0] = (uint8_t)nsym;                // 256 wraps to 0
    }
    __syncwarp();
    if (nsym > 16) return false;
    *meta_len = nsym + 1;
    const uint32_t per = nsym > 4 ? 2 : nsym > 2 ? 4 : nsym > 1 ? 8 : 0;
    if (!per) { *out_len = 0; return true; }
    const uint32_t bits = 8 / per, olen = (n + per - 1) / per;
    // each lane builds whole output bytes; `out` is 16-byte aligned scratch
    uint32_t jdone = 0;
    if ((((uintptr_t)in) & 15) == 0) {
        // 16 input bytes per lane and load, four loads in flight; 16/per output bytes per load
        const uint4 *v = (const uint4 *)in;
        const uint32_t nv = n >> 4, ob = 16 / per;            // ob = 2, 4 or 8 output bytes
        auto pack16 = [&](const uint4 &q, uint32_t vi) {
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            uint32_t c[16];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) c[4 * a + b] = code[(w[a] >> (8 * b)) & 0xff];
            if (per == 4) {
                uint32_t r = 0;
#pragma unroll
                for (int k = 0; k < 16; k++) r |= c[k] << (2 * k);
                *(uint32_t *)(out + vi * 4) = r;
            } else if (per == 2) {
                uint32_t r0 = 0, r1 = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) { r0 |= c[k] << (4 * k); r1 |= c[8 + k] << (4 * k); }
                *(uint2 *)(out + vi * 8) = make_uint2(r0, r1);
            } else {
                uint32_t r = 0;
#pragma unroll
                for (int k = 0; k < 16; k++) r |= c[k] << k;
                *(uint16_t *)(out + vi * 2) = (uint16_t)r;
            }
        };
        uint32_t i = lane;
        for (; i + 96 < nv; i += 128) {
            uint4 q0 = __ldg(v + i), q1 = __ldg(v + i + 32), q2 = __ldg(v + i + 64), q3 = __ldg(v + i + 96);
            pack16(q0, i); pack16(q1, i + 32); pack16(q2, i + 64); pack16(q3, i + 96);
        }
        for (; i < nv; i += 32) pack16(__ldg(v + i), i);
        jdone = nv * ob;
    }
    for (uint32_t j = jdone + lane; j < olen; j += 32) {
        uint32_t v = 0, base = j * per;
        for (uint32_t q = 0; q < per && base + q < n; q++) v |= (uint32_t)code[in[base + q]] << (q * bits);
        out[j] = (uint8_t)v;
    }
    *out_len = olen;
    __syncwarp();
    return true;
}

struct PackMap {
    uint8_t map[16];
    int per;            // symbols per byte: 0 (single symbol), 8, 4, 2, or 1 (not packed)
};

// pack.c:161-194.  Returns bytes consumed, 0 on failure.
__device__ inline int unpack_meta(const uint8_t *d, uint32_t dlen, PackMap &pm) {
    for (int i = 0; i < 16; i++) pm.map[i] = 0;
    pm.per = 0;
    if (dlen == 0) return 0;
    uint32_t n = d[0] ? d[0] : 256;
    if (n <= 1) pm.per = 0; else if (n <= 2) pm.per = 8; else if (n <= 4) pm.per = 4;
    else if (n <= 16) pm.per = 2; else { pm.per = 1; return 1; }
    if (dlen <= 1) return 0;
    uint32_t c = 0;
    while (c < n && 1 + c < dlen) { pm.map[c] = d[1 + c]; c++; }
    return c < n ? 0 : (int)(1 + c);
}

// pack.c:207-344
__device__ inline bool warp_unpack(const uint8_t *src, uint32_t len, uint8_t *dst, uint32_t out_len,
                                   const PackMap &pm, uint8_t *smem, int lane) {
    const int per = pm.per;
    if (per == 1) { warp_copy(dst, src, len, lane); __syncwarp(); return true; }
    if (per == 0) {
        for (uint32_t i = lane; i < out_len; i += 32) dst[i] = pm.map[0];
        __syncwarp();
        return true;
    }
    if ((out_len + per - 1) / per > len) return false;
    const uint32_t bits = 8 / per, cm = (1u << bits) - 1;
    const uint32_t whole = out_len / per;
    if (per == 4 && (((uintptr_t)dst) & 3) == 0) {
        uint32_t *lut = (uint32_t *)smem;                      // byte -> 4 symbols
        for (int x = lane; x < 256; x += 32)
            lut[x] = pm.map[x & 3] | (pm.map[(x >> 2) & 3] << 8) | (pm.map[(x >> 4) & 3] << 16) |
                     ((uint32_t)pm.map[(x >> 6) & 3] << 24);
        __syncwarp();
        uint32_t *d4 = (uint32_t *)dst;
        uint32_t j0 = 0;
        if (((((uintptr_t)src) & 3) | (((uintptr_t)dst) & 15)) == 0) {
            // 4 packed bytes per lane and load -> one 16-byte store; four loads in flight
            const uint32_t *s4 = (const uint32_t *)src;
            uint4 *d16 = (uint4 *)dst;
            const uint32_t nq = whole >> 2;
            auto expand = [&](uint32_t w) {
                return make_uint4(lut[w & 0xff], lut[(w >> 8) & 0xff], lut[(w >> 16) & 0xff], lut[w >> 24]);
            };
            uint32_t i = lane;
            for (; i + 96 < nq; i += 128) {
                uint32_t w0 = s4[i], w1 = s4[i + 32], w2 = s4[i + 64], w3 = s4[i + 96];   // (written by this kernel: no ld.nc)
                d16[i] = expand(w0); d16[i + 32] = expand(w1); d16[i + 64] = expand(w2); d16[i + 96] = expand(w3);
            }
            for (; i < nq; i += 32) d16[i] = expand(s4[i]);
            j0 = nq << 2;
        }
        for (uint32_t j = j0 + lane; j < whole; j += 32) d4[j] = lut[src[j]];
    } else if (per == 2 && (((uintptr_t)dst) & 1) == 0) {
        uint16_t *lut = (uint16_t *)smem;
        for (int x = lane; x < 256; x += 32) lut[x] = (uint16_t)(pm.map[x & 15] | (pm.map[x >> 4] << 8));
        __syncwarp();
        uint16_t *d2 = (uint16_t *)dst;
        for (uint32_t j = lane; j < whole; j += 32) d2[j] = lut[src[j]];
    } else {
        for (uint32_t j = lane; j < whole; j += 32) {
            uint32_t c = src[j];
            for (int q = 0; q < per; q++) { dst[j * per + q] = pm.map[c & cm]; c >>= bits; }
        }
    }
    // trailing partial byte
    uint32_t done = whole * per;
    if (done < out_len && lane == 0) {
        uint32_t c = src[whole];
        for (uint32_t i = done; i < out_len; i++) { dst[i] = pm.map[c & cm]; c >>= bits; }
    }
    __syncwarp();
    return true;
}

// ------------------------------------------------------------------- RLE
// Encode.  A symbol is run-length coded iff it repeats its predecessor more
// often than not (score = sum of +1 / -1 over its occurrences, rle.c:48-98).
// Literal stream: one byte per maximal run of such a symbol, one byte per
// occurrence otherwise; (run length - 1) goes to a varint stream in run order.
//   meta = [nsyms][syms ascending][varints...]
__device__ inline void warp_rle_encode(const uint8_t *in, uint32_t n, uint8_t *lits, uint32_t *lits_len,
                                       uint8_t *meta, uint32_t *meta_len, uint8_t *smem, int lane) {
    int32_t *score = (int32_t *)smem;                     // 1 KiB
    uint8_t *stage = smem + 1024;                         // 512 B: one block of input, see below
    for (int j = lane; j < 256; j += 32) score[j] = 0;
    __syncwarp();
    const bool al16 = (((uintptr_t)in) & 15) == 0;
    {
        // score[c] += (in[p-1] == c) ? +1 : -1 over all positions.  Aligned inputs: 16 bytes per
        // lane and load (two loads in flight); a stretch of equal bytes inside a lane's 16 is
        // one atomic.  (`in` may have been written by this kernel: plain loads.)
        uint32_t done = 0;
        if (al16) {
            const uint4 *v = (const uint4 *)in;
            const uint32_t nv = n >> 4;
            uint32_t carry = 0x100;                       // byte before the round's first byte (none at p = 0)
            auto tally = [&](const uint4 &q, bool on, uint32_t prev) {
                if (!on) return;
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
                uint32_t cur = w[0] & 0xff;
                int acc = cur == prev ? 1 : -1;
#pragma unroll
                for (int k = 1; k < 16; k++) {
                    const uint32_t c = (w[k >> 2] >> (8 * (k & 3))) & 0xff;
                    if (c == cur) acc++;
                    else { atomicAdd(&score[cur], acc); cur = c; acc = -1; }
                }
                atomicAdd(&score[cur], acc);
            };
            for (uint32_t base = 0; base < nv; base += 64) {
                const uint32_t i0 = base + lane, i1 = base + 32 + lane;
                const bool on0 = i0 < nv, on1 = i1 < nv;
                uint4 q0 = on0 ? v[i0] : make_uint4(0, 0, 0, 0), q1 = on1 ? v[i1] : make_uint4(0, 0, 0, 0);
                uint32_t l0 = q0.w >> 24, l1 = q1.w >> 24;
                uint32_t p0 = __shfl_up_sync(FULL, l0, 1), p1 = __shfl_up_sync(FULL, l1, 1);
                const uint32_t e0 = __shfl_sync(FULL, l0, 31), e1 = __shfl_sync(FULL, l1, 31);
                if (lane == 0) { p0 = carry; p1 = e0; }
                carry = e1;
                tally(q0, on0, p0);
                tally(q1, on1, p1);
            }
            done = nv << 4;
        }
        for (uint32_t base = done; base < n; base += 32) {
            uint32_t p = base + lane;
            if (p < n) {
                uint32_t c = in[p];
                bool same = p && in[p - 1] == c;
                atomicAdd(&score[c], same ? 1 : -1);
            }
        }
    }
    __syncwarp();
    uint32_t nsyms;
    {
        uint32_t loc = 0;
        bool pr[8];
#pragma unroll
        for (int t = 0; t < 8; t++) { pr[t] = score[lane * 8 + t] > 0; loc += pr[t]; }
        uint32_t incl = warp_incl_scan(loc, lane);
        nsyms = __shfl_sync(FULL, incl, 31);
        uint32_t r = incl - loc;
#pragma unroll
        for (int t = 0; t < 8; t++)
            if (pr[t]) meta[1 + r++] = (uint8_t)(lane * 8 + t);
        if (lane == 0) meta[0] = (uint8_t)nsyms;
    }
    uint8_t *runs = meta + 1 + nsyms;
    const uint32_t lt = lanemask_lt();
    uint32_t nl = 0, nr = 0;             // literals / run bytes written so far
    uint32_t open_start = 0;             // start of the run still open from earlier chunks
    bool open = false;
    // one round: positions base..base+31; c = byte at p (0x100 at and past n), pc = byte at p-1.
    // Position n acts as a terminating "emitter" that only closes an open run.
    auto round = [&](uint32_t base, uint32_t c, uint32_t pc) {
        uint32_t p = base + lane;
        bool isr = p < n && score[c & 0xff] > 0;
        bool cont = isr && p && pc == c;                  // continues a run: emits nothing
        bool emit = (p < n && !cont) || p == n;
        uint32_t E = __ballot_sync(FULL, emit);
        uint32_t R = __ballot_sync(FULL, emit && isr);     // emitters that open a run
        // previous emitter of each emitter lane
        uint32_t below = E & lt;
        bool prev_in_chunk = below != 0;
        int pl = 31 - __clz(below);
        bool prev_open = prev_in_chunk ? ((R >> pl) & 1) : open;
        uint32_t prev_pos = prev_in_chunk ? base + pl : open_start;
        bool closes = emit && prev_open;
        uint32_t rl = closes ? p - prev_pos - 1 : 0;
        uint32_t C = __ballot_sync(FULL, closes);
        if (C) {                                           // (uniform) most rounds close no run
            uint32_t vs = closes ? var_size_u32(rl) : 0;
            uint32_t vincl = warp_incl_scan(vs, lane);
            if (closes) var_put_u32(runs + nr + vincl - vs, rl);
            nr += __shfl_sync(FULL, vincl, 31);
        }
        bool lit = emit && p < n;
        uint32_t L = __ballot_sync(FULL, lit);
        if (lit) lits[nl + __popc(L & lt)] = (uint8_t)c;
        nl += __popc(L);
        if (E) {
            int hl = 31 - __clz(E);
            open = (R >> hl) & 1;
            open_start = base + hl;
        }
    };
    uint32_t base = 0;
    if (al16 && n >= 512) {
        // 512-byte blocks staged through shared memory; the next block is already on its way
        // while the sixteen rounds of the current one run.
        const uint4 *v = (const uint4 *)in;
        const uint32_t nblk = n >> 9;
        const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
        uint4 nxt = v[lane];
        uint32_t carry = 0;
        for (uint32_t blk = 0; blk < nblk; blk++) {
            __syncwarp();
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stage_s + 16 * lane), "r"(nxt.x), "r"(nxt.y),
                         "r"(nxt.z), "r"(nxt.w) : "memory");
            if (blk + 1 < nblk) nxt = v[(blk + 1) * 32 + lane];
            __syncwarp();
#pragma unroll 4
            for (int k = 0; k < 16; k++) {
                const uint32_t o = 32 * k + lane;
                const uint32_t c = stage[o];
                const uint32_t pc = o ? stage[o - 1] : carry;
                round(base + 32 * k, c, pc);
            }
            carry = stage[511];
            base += 512;
        }
    }
    for (; base < n + 32; base += 32) {
        uint32_t p = base + lane;
        uint32_t c = p < n ? in[p] : 0x100;
        uint32_t pc = (p && p < n) ? in[p - 1] : 0x200;
        round(base, c, pc);
        if (base >= n) break;
    }
    *lits_len = nl;
    *meta_len = 1 + nsyms + nr;
    __syncwarp();
}

// Decode (rle.c:142-189).  Returns false where the reference returns NULL.
__device__ inline bool warp_rle_decode(const uint8_t *lit, uint32_t lit_len, const uint8_t *run,
                                       uint32_t run_len, const uint8_t *syms, uint32_t nsyms,
                                       uint8_t *out, uint32_t *out_len, uint8_t *smem, int lane) {
    uint8_t *isr = smem;                                  // 256 flags
    for (int j = lane; j < 64; j += 32) ((uint32_t *)isr)[j] = 0;
    __syncwarp();
    for (uint32_t j = lane; j < nsyms; j += 32) isr[syms[j]] = 1;
    __syncwarp();
    const uint32_t lt = lanemask_lt();
    const uint32_t cap = *out_len;
    uint32_t li = 0, rp = 0, op = 0;
    while (li < lit_len) {
        uint32_t c = li + lane < lit_len ? lit[li + lane] : 0;
        bool r = li + lane < lit_len && isr[c];
        uint32_t M = __ballot_sync(FULL, r);
        uint32_t b = rp + lane < run_len ? run[rp + lane] : 0x80;
        bool term = !(b & 0x80);
        uint32_t Tm = __ballot_sync(FULL, term);
        uint32_t nv = __popc(Tm), nm = __popc(M);
        uint32_t take = min(32u, lit_len - li);
        if (rp >= run_len) {
            // run stream exhausted: the reference reads 0 and does not advance
            Tm = 0; nv = 32;
        } else if (nm > nv) {
            if (nv == 0) return false;                    // a varint longer than the window: corrupt
            take = __fns(M, 0, nv + 1);                   // stop before the first unserved run literal
        }
        bool mine = (uint32_t)lane < take;
        uint32_t k = __popc(M & lt);                      // which varint serves this literal
        uint32_t val = 0;
        uint32_t used = 0;
        if (Tm) {
            uint32_t served = __popc(M & ((take >= 32) ? FULL : ((1u << take) - 1)));
            used = served ? __fns(Tm, 0, served) + 1 : 0;
            // value of varint k: bytes (prev terminator+1 .. terminator)
            uint32_t e = __fns(Tm, 0, k + 1);
            uint32_t s = k ? __fns(Tm, 0, k) + 1 : 0;
            bool want = mine && r && e != 0xffffffffu;
#pragma unroll
            for (int q = 0; q < 5; q++) {
                uint32_t src = s + q;
                uint32_t bb = __shfl_sync(FULL, b, src & 31);
                if (want && src <= e) val = (val << 7) | (bb & 0x7f);
            }
        }
        uint32_t len = mine ? ((r ? val : 0) + 1) : 0;
        // the reference's checks: outp >= out_end before each literal, outp+rlen >= out_end for runs
        uint32_t incl = warp_incl_scan(len, lane);
        uint32_t o = op + incl - len;
        bool bad = mine && (o >= cap || (r && val && (uint64_t)o + val >= cap));
        if (__any_sync(FULL, bad)) return false;
        if (mine && len <= 8) for (uint32_t q = 0; q < len; q++) out[o + q] = (uint8_t)c;
        uint32_t big = __ballot_sync(FULL, mine && len > 8);
        while (big) {
            int l = __ffs(big) - 1;
            big &= big - 1;
            uint32_t oo = __shfl_sync(FULL, o, l), ll = __shfl_sync(FULL, len, l), cc = __shfl_sync(FULL, c, l);
            for (uint32_t q = lane; q < ll; q += 32) out[oo + q] = (uint8_t)cc;
        }
        op += __shfl_sync(FULL, incl, 31);
        li += take;
        rp += used;
    }
    *out_len = op;
    __syncwarp();
    return true;
}

}  // namespace b200
