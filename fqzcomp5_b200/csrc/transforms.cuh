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
        for (uint32_t i = lane; i < nv; i += 32) {
            uint4 q = v[i];
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) code[(w[a] >> (8 * b)) & 0xff] = 1;
        }
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
    for (uint32_t j = lane; j < olen; j += 32) {
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
        for (uint32_t j = lane; j < whole; j += 32) d4[j] = lut[src[j]];
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
    for (int j = lane; j < 256; j += 32) score[j] = 0;
    __syncwarp();
    for (uint32_t base = 0; base < n; base += 32) {
        uint32_t p = base + lane;
        if (p < n) {
            uint32_t c = in[p];
            bool same = p && in[p - 1] == c;
            atomicAdd(&score[c], same ? 1 : -1);
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
    for (uint32_t base = 0; base < n + 32; base += 32) {
        // position n acts as a terminating "emitter" that only closes an open run
        uint32_t p = base + lane;
        uint32_t c = p < n ? in[p] : 0x100;
        bool isr = p < n && score[c] > 0;
        bool cont = isr && p && in[p - 1] == c;           // continues a run: emits nothing
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
        uint32_t vs = closes ? var_size_u32(rl) : 0;
        uint32_t vincl = warp_incl_scan(vs, lane);
        if (closes) var_put_u32(runs + nr + vincl - vs, rl);
        nr += __shfl_sync(FULL, vincl, 31);
        bool lit = emit && p < n;
        uint32_t L = __ballot_sync(FULL, lit);
        if (lit) lits[nl + __popc(L & lt)] = (uint8_t)c;
        nl += __popc(L);
        if (E) {
            int hl = 31 - __clz(E);
            open = (R >> hl) & 1;
            open_start = base + hl;
        }
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
