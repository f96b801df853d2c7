// dec_staged.cuh -- order-1 streams behind PACK / RLE whose alphabet is large (the packed bases of a
// sequence block: up to 256 x 256 table entries for 64 KiB of packed data) decoded in four launches
// instead of by one warp from end to end:
//   head   one warp per stream: container header, the run-length meta-data and the table text through the
//          order-0 decoder (serial state chains: nothing else can be done with them), alphabet
//   table  one CTA per stream: the table text parsed by 256 threads (token boundaries by a scan over the
//          per-byte transition functions), rows scaled and turned into the one-sector bucket records
//   chain  one warp per stream, nothing but the N state chains (28 warps per SM: a 1 GB block is one wave)
//   post   one CTA per stream: run-length expansion and unpack
// A stream the head stage does not take (small alphabet, anything unusual) goes to the general kernel
// (dec_kernel<true>) untouched, so every error path keeps the behaviour of dec_stream.
//
// Reference behaviour restated here (never its code): rANS_static4x16pr.c:1696-1887 (container),
// rANS_static16_int.h:425-536 (order-1 table), rANS_static32x16pr.c:531-758 / rANS_static4x16pr.c:524-821
// (state chains), rle.c:142-189, pack.c:161-344.
#pragma once
#include <stddef.h>
#include <type_traits>
#include "common.cuh"
#include "rans_decode.cuh"
#include "transforms.cuh"

namespace b200 {

struct __align__(16) DecPrep {
    uint32_t state;             // 0: left to the general kernel, 1: staged, 2: failed (status holds the result)
    int32_t  status;
    uint32_t flag, N;
    uint32_t out_size;          // bytes the stream decodes to
    uint32_t t1_size;           // bytes the entropy decoder produces (before RLE expansion / unpack)
    uint32_t unpacked_sz, u_meta;
    uint32_t shift, nsym;
    uint32_t pad_[2];
    const uint8_t *meta;        // run-length meta-data (decoded or stored raw), or null
    const uint8_t *rows, *rows_end;     // table text behind the alphabet
    const uint8_t *pay;         // the N initial states (known after the table when it is stored raw)
    const uint8_t *end;         // end of the compressed stream
    uint8_t *tb;                // compact rows | bucket records (DecO1Big), from the pool
    uint8_t *t1, *t2, *t3;
    uint8_t *scr;               // scratch of the post stage (terminator bitmap and its prefix counts)
    uint8_t sym[256];           // rank -> symbol (16-byte aligned)
    PackMap pm;
};
static_assert(offsetof(DecPrep, sym) % 16 == 0, "DecPrep::sym alignment");

// what became of the streams routed to the head stage, per reason (diagnostics: b200rans_dec_staged_stats)
//   0 taken   1 failed after being taken   2.. handed back: 2 header  3 pack meta  4 rle header  5 no payload
//   6 order-1 header  7 table peek  8 alphabet  9 small alphabet
__device__ unsigned long long g_dec_stats[16];

// ------------------------------------------------------------------------ head (one warp)
// Mirrors dec_stream (kernels.cu) up to the payload and dec_o1 (rans_decode.cuh) up to the alphabet.
// smem: one DecO0Smem.
__device__ inline void dec_head(DecJob &J, DecPrep &P, uint8_t *smem, const Pool &pool, int lane) {
    DecO0Smem &S0 = *(DecO0Smem *)smem;
    const uint8_t *in = J.in, *in_end = J.in + J.in_size;
    uint32_t in_size = J.in_size;
    uint32_t state = 0;                 // what the stage decides: 0 general kernel, 1 staged, 2 failed
    int32_t status = ST_FAIL;
    uint32_t why = 2;
    do {
        if (in_size == 0 || !J.tmp) break;
        const int flag = *in++; in_size--;
        if ((flag & (X_STRIPE | X_CAT)) || !(flag & 1) || !(flag & (X_PACK | X_RLE))) break;
        const int do_pack = flag & X_PACK, do_rle = flag & X_RLE, no_size = flag & X_NOSZ, do_simd = flag & X_32;
        const int N = do_simd ? 32 : 4;
        uint32_t osz = J.out_cap;
        if (!no_size) { int sz = var_get_u32(in, in_end, &osz); in += sz; in_size -= sz; }
        if (J.out_cap < osz) break;
        uint8_t *out = J.out, *tmp = J.tmp;
        uint8_t *t1, *t2, *t3;                                          // rANS_static4x16pr.c:1760-1782
        if (do_pack && do_rle) { t1 = out; t2 = tmp; t3 = out; }
        else if (do_pack)      { t1 = tmp; t2 = tmp; t3 = out; }
        else                   { t1 = tmp; t2 = out; t3 = out; }
        uint32_t t1_size = osz, unpacked_sz = 0;
        PackMap pm;
        pm.per = 1;
        for (int i = 0; i < 16; i++) pm.map[i] = 0;
        why = 3;
        if (do_pack) {
            int c = unpack_meta(in, in_size, pm);
            if (!c) break;
            unpacked_sz = osz;
            in += c; in_size -= c;
            uint32_t psz;
            int sz = var_get_u32(in, in_end, &psz);
            in += sz; in_size -= sz;
            if (psz > t1_size) break;
            t1_size = psz;
        }
        why = 4;
        const uint8_t *meta = nullptr, *cmeta = nullptr;
        uint32_t u_meta = 0, cmeta_size = 0;
        if (do_rle) {
            uint32_t c_meta, rle_len;
            uint32_t sz = var_get_u32(in, in_end, &u_meta);
            sz += var_get_u32(in + sz, in_end, &rle_len);
            if (rle_len > t1_size) break;
            if (u_meta & 1) {
                meta = in + sz;
                uint32_t left = (uint32_t)(in_end - meta);
                u_meta = u_meta / 2 > left ? left : u_meta / 2;
                c_meta = u_meta;
                if (u_meta > J.out_cap + 1024) break;                   // (the post stage's scratch is sized by this)
            } else {
                sz += var_get_u32(in + sz, in_end, &c_meta);
                u_meta /= 2;
                if (u_meta > J.out_cap + 1024 || in_size < sz) break;
                cmeta = in + sz; cmeta_size = in_size - sz;             // decoded below, once the stream is taken
            }
            if ((uint64_t)c_meta + sz > in_size) break;
            in += c_meta + sz; in_size -= c_meta + sz;
            t1_size = rle_len;
        }
        why = 5;
        if (!in_size) break;
        why = 6;
        // ---- the order-1 stream (dec_o1)
        if (in_size < (uint32_t)(N == 4 ? 16 : N * 4) || t1_size >= 0x7fffffffu) break;
        if (N == 32 && t1_size < 32) break;
        const uint8_t *end = in + in_size, *cp = in, *tend = end, *after = nullptr;
        const uint32_t shift = *cp >> 4;
        if (shift != 10 && shift != 12) break;
        const bool comp = *cp++ & 1;
        uint32_t usz = 0, csz = 0;
        const uint8_t *ctab = nullptr;
        if (comp) {
            cp += var_get_u32(cp, end, &usz);
            cp += var_get_u32(cp, end, &csz);
            if (csz > (uint32_t)(end - cp) || usz > 257 * 257 * 3) break;
            after = cp + csz;
            ctab = cp;
            // the alphabet heads the table text: a look at its first bytes says whether the stream is taken
            const uint32_t pk = usz <= 768 ? usz : 768;
            why = 7;
            if (dec_o0<4>(ctab, csz, tmp, pk, S0, lane)) break;
            __threadfence_block();
            __syncwarp();
            cp = tmp; tend = tmp + pk;
        }
        why = 8;
        for (int j = lane; j < 256; j += 32) S0.tab[j] = 0;
        __syncwarp();
        int n = 0;
        if (lane == 0) n = get_alphabet(cp, tend, S0.tab);
        n = __shfl_sync(FULL, n, 0);
        if (!n || cp + n >= tend) break;
        __syncwarp();
        uint32_t nsym = 0, myrank = 0, pres8 = 0;
        {
            uint32_t loc = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) if (S0.tab[lane * 8 + t]) { loc++; pres8 |= 1u << t; }
            const uint32_t incl = warp_incl_scan(loc, lane);
            nsym = __shfl_sync(FULL, incl, 31);
            myrank = incl - loc;
        }
        why = 9;
        if (nsym <= 64 || !(__shfl_sync(FULL, pres8, 0) & 1)) break;    // small alphabets: the general kernel
        // ---- taken: from here on a failure is this stream's result
        state = 2;
        {
            uint32_t r = myrank;
#pragma unroll
            for (int t = 0; t < 8; t++) if ((pres8 >> t) & 1) P.sym[r++] = (uint8_t)(lane * 8 + t);
        }
        if (comp) {
            uint8_t *tb = pool_alloc(pool, usz + 64, lane);
            if (!tb) { status = ST_UNSUPPORTED; break; }
            if (dec_o0<4>(ctab, csz, tb, usz, S0, lane)) break;
            __threadfence_block();
            __syncwarp();
            cp = tb; tend = tb + usz;
        }
        cp += n;
        if (cp >= tend) break;
        const uint32_t ns1 = nsym + 1;
        const uint32_t ent_bytes = (nsym * ns1 * 4 + 31) & ~31u;
        uint8_t *tb2 = pool_alloc(pool, ent_bytes + (nsym << 11) + 256, lane);
        if (!tb2) { status = ST_UNSUPPORTED; break; }
        uint8_t *mbuf = tmp + ((J.out_cap + 15) & ~15u);
        if (cmeta) {
            int e = do_simd ? dec_o0<32>(cmeta, cmeta_size, mbuf, u_meta, S0, lane)
                            : dec_o0<4>(cmeta, cmeta_size, mbuf, u_meta, S0, lane);
            if (e) break;
            __syncwarp();
            meta = mbuf;
        }
        if (do_rle) {
            if (u_meta == 0) break;
            const uint32_t nsyms = *meta ? *meta : 256;
            if (u_meta < 1 + nsyms) break;
        }
        if (lane == 0) {
            P.flag = (uint32_t)flag; P.N = (uint32_t)N;
            P.out_size = osz; P.t1_size = t1_size; P.unpacked_sz = unpacked_sz; P.u_meta = u_meta;
            P.shift = shift; P.nsym = nsym;
            P.meta = meta; P.rows = cp; P.rows_end = tend; P.pay = after; P.end = end;
            P.tb = tb2; P.t1 = t1; P.t2 = t2; P.t3 = t3;
            P.scr = tmp + (((size_t)((J.out_cap + 15) & ~15u) + u_meta + 16 + 255) & ~(size_t)255);
            P.pm = pm;
        }
        state = 1;
        status = ST_OK;
    } while (0);
    __syncwarp();
    if (lane == 0) {
        P.state = state;
        P.status = status;
        if (state == 0) J.route = 1;
        atomicAdd(&g_dec_stats[state == 1 ? 0 : state == 2 ? 1 : why], 1ull);
    }
}

// ------------------------------------------------------------------------ table (one CTA of 256 threads)
constexpr int DT_THREADS = 256, DT_WARPS = 8;
constexpr uint32_t DT_SPAN = 32, DT_TILE = DT_THREADS * DT_SPAN;
struct __align__(16) DecTabSmem {
    uint32_t wF[DT_WARPS], wc[DT_WARPS][4];
    uint32_t end_off, err;
    uint32_t cnt[DT_WARPS][64];
    uint32_t crow[DT_WARPS][260];
};

// The order-1 rows (rANS_static16_int.h:425-456, 488-530): a stream of tokens -- a varint count, or 0x00 and a
// raw byte z meaning z further zero counts -- without delimiters between rows.  Which byte starts a token is a
// three-state machine (0 token start, 1 inside a varint, 2 the raw byte of a zero run).  Every thread takes 32
// bytes: a first walk yields the span's transition function and the table slots it covers for each of the
// three entry states; a scan over the CTA composes them; a second walk, now from the known state and slot,
// stores the counts.  Raw counts go to raw[row * stride + col + 1] (zeroed beforehand).  Same error rules as
// parse_o1_rows (rans_decode.cuh).  Returns 0 and the end of the table, or 1.
__device__ inline int cta_parse_o1_rows(const uint8_t *cp, const uint8_t *tend, uint32_t nsym, uint32_t tot,
                                        uint16_t *raw, uint32_t stride, const uint8_t **end_out, DecTabSmem &S) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t total = nsym * nsym, avail = (uint32_t)(tend - cp);
    uint32_t state = 0, slot_base = 0, my_end = 0;
    int err = 0;
    if (tid == 0) S.end_off = 0;
    for (uint32_t base = 0; slot_base < total; base += DT_TILE) {
        if (base >= avail) return 1;                        // ran out of bytes before the last row (uniform)
        const uint32_t i0 = base + DT_SPAN * (uint32_t)tid;
        // bytes [i0, i0 + 40) of the text: aligned words that hold at least one byte of it, funnel-shifted
        uint32_t v[10];
        {
            const uint8_t *A = cp + i0;
            const uint32_t a = (uint32_t)((uintptr_t)A & 3);
            const uint32_t *wp = (const uint32_t *)(A - a);
            uint32_t w[11];
#pragma unroll
            for (int j = 0; j < 11; j++) w[j] = ((uint64_t)i0 + 4 * j < (uint64_t)avail + a) ? wp[j] : 0u;
#pragma unroll
            for (int j = 0; j < 10; j++) v[j] = __funnelshift_r(w[j], w[j + 1], 8 * a);
        }
        auto byte_at = [&](int k) { return (v[k >> 2] >> (8 * (k & 3))) & 0xffu; };
        // ---- walk 1: exit state and slots for each entry state
        uint32_t s0 = 0, s1 = 1, s2 = 2, c0 = 0, c1 = 0, c2 = 0;
        const bool whole = (uint64_t)i0 + DT_SPAN + 8 <= avail;     // the span and its look-ahead lie inside the text
        auto walk1 = [&](auto all) {
#pragma unroll
            for (int k = 0; k < (int)DT_SPAN; k++) {
                const uint32_t b = byte_at(k);
                if (decltype(all)::value || i0 + k < avail) {
                    const uint32_t from0 = b == 0 ? 2u : (b < 128 ? 0u : 1u), from1 = b < 128 ? 0u : 1u;
                    c0 += s0 == 0 ? 1u : s0 == 2 ? b : 0u;
                    c1 += s1 == 0 ? 1u : s1 == 2 ? b : 0u;
                    c2 += s2 == 0 ? 1u : s2 == 2 ? b : 0u;
                    s0 = s0 == 0 ? from0 : s0 == 1 ? from1 : 0u;
                    s1 = s1 == 0 ? from0 : s1 == 1 ? from1 : 0u;
                    s2 = s2 == 0 ? from0 : s2 == 1 ? from1 : 0u;
                }
            }
        };
        if (whole) walk1(std::true_type{}); else walk1(std::false_type{});
        uint32_t F = s0 | (s1 << 2) | (s2 << 4);
        // ---- inclusive scan over the warp: (earlier) then (this)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t gF = __shfl_up_sync(FULL, F, o), g0 = __shfl_up_sync(FULL, c0, o),
                           g1 = __shfl_up_sync(FULL, c1, o), g2 = __shfl_up_sync(FULL, c2, o);
            if (lane >= o) {
                const uint32_t a0 = gF & 3, a1 = (gF >> 2) & 3, a2 = (gF >> 4) & 3;
                const uint32_t n0 = g0 + (a0 == 0 ? c0 : a0 == 1 ? c1 : c2);
                const uint32_t n1 = g1 + (a1 == 0 ? c0 : a1 == 1 ? c1 : c2);
                const uint32_t n2 = g2 + (a2 == 0 ? c0 : a2 == 1 ? c1 : c2);
                F = ((F >> (2 * a0)) & 3) | (((F >> (2 * a1)) & 3) << 2) | (((F >> (2 * a2)) & 3) << 4);
                c0 = n0; c1 = n1; c2 = n2;
            }
        }
        if (lane == 31) { S.wF[wid] = F; S.wc[wid][0] = c0; S.wc[wid][1] = c1; S.wc[wid][2] = c2; }
        __syncthreads();
        // entry of this warp, then of this thread
        uint32_t x = state, sl = slot_base;
        for (int w = 0; w < wid; w++) { sl += S.wc[w][x]; x = (S.wF[w] >> (2 * x)) & 3; }
        {
            uint32_t eF = __shfl_up_sync(FULL, F, 1), e0 = __shfl_up_sync(FULL, c0, 1),
                     e1 = __shfl_up_sync(FULL, c1, 1), e2 = __shfl_up_sync(FULL, c2, 1);
            if (lane == 0) { eF = 0 | (1 << 2) | (2 << 4); e0 = e1 = e2 = 0; }
            sl += x == 0 ? e0 : x == 1 ? e1 : e2;
            x = (eF >> (2 * x)) & 3;
        }
        // the tile's exit becomes the next tile's entry (every thread computes the same)
        for (int w = 0; w < DT_WARPS; w++) { slot_base += S.wc[w][state]; state = (S.wF[w] >> (2 * state)) & 3; }
        // ---- walk 2: tokens that start in this span
        bool rc_valid = false;
        uint32_t row = 0, col = 0;
        auto walk2 = [&](auto all) {
            constexpr bool ALL = decltype(all)::value;
#pragma unroll
            for (int k = 0; k < (int)DT_SPAN; k++) {
                const uint32_t i = i0 + k;
                if (!ALL && i >= avail) continue;
                const uint32_t b = byte_at(k);
                if (x == 1) { x = b < 128 ? 0u : 1u; continue; }
                if (x == 2) { sl += b; x = 0; continue; }
                uint32_t val = 0, slots = 1, len = 1;
                bool terr = false;
                if (b == 0) {
                    if (!ALL && i + 1 >= avail) terr = true;
                    else { slots = 1 + byte_at(k + 1); len = 2; }
                    x = 2;
                } else {
                    uint32_t c = b;
                    val = c & 0x7f;
#pragma unroll
                    for (int q = 1; q < 6; q++)
                        if ((c & 0x80) && len == (uint32_t)q && (ALL || i + q < avail)) { c = byte_at(k + q); val = (val << 7) | (c & 0x7f); len = q + 1; }
                    terr = (c & 0x80) || val == 0 || val > tot;
                    x = b < 128 ? 0u : 1u;
                }
                if (sl < total) {
                    if (!rc_valid) { row = sl / nsym; col = sl - row * nsym; rc_valid = true; }
                    if (terr || col + slots > nsym) err = 1;                // a zero run never crosses a row
                    else if (val) raw[row * stride + col + 1] = (uint16_t)val;
                    if (sl + slots >= total) my_end = i + len;              // the token that completes the last row
                    col += slots;
                    if (col >= nsym) { col -= nsym; row++; }
                }
                sl += 1;
            }
        };
        if (whole) walk2(std::true_type{}); else walk2(std::false_type{});
        if (__syncthreads_or(err)) return 1;
    }
    if (my_end) S.end_off = my_end;                         // exactly one thread saw the closing token
    __syncthreads();
    const uint32_t eo = S.end_off;
    if (!eo) return 1;
    *end_out = cp + eo;
    return 0;
}

// One row of the large-alphabet tables (the `big` branch of dec_o1): raw counts -> scaled entries in slot order
// (ent row, closed by 0xffffffff) and the 64 one-sector bucket records.  cnt (64 words) and crow (257 words) are
// the warp's scratch.  Returns non-zero when the row is invalid.
__device__ __forceinline__ int dec_big_row(const uint16_t *rawrow, uint32_t *row, uint32_t *recs_row, uint32_t nsym,
                                           uint32_t shift, uint32_t *cnt, uint32_t *crow, int lane) {
    const uint32_t tot = 1u << shift, bwb = shift - 6, B = 1u << bwb;
    uint32_t f[8], tsum = 0, nz = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const uint32_t c = lane * 8 + t;
        f[t] = c < nsym ? rawrow[c + 1] : 0;
        tsum += f[t];
        nz += f[t] ? 1 : 0;
    }
    if (warp_sum(tsum) > tot) return 1;                                 // (uniform) keeps the packed scan exact
    uint32_t pk = warp_incl_scan(tsum | (nz << 16), lane);
    const uint32_t all = __shfl_sync(FULL, pk, 31);
    const uint32_t rsum = all & 0xffff, nnz = all >> 16;
    pk -= tsum | (nz << 16);
    int sh = 0;
    if (rsum) { uint32_t z = rsum; while (z < tot) { z *= 2; sh++; } }
    if (rsum && (rsum << sh) != tot) return 1;                          // (uniform)
    cnt[lane] = 0; cnt[lane + 32] = 0;
    __syncwarp();                                                       // every lane has read its raw counts
    uint32_t x = (pk & 0xffff) << sh, k = pk >> 16;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const uint32_t ff = f[t] << sh;
        if (ff) {
            const uint32_t e = (x << 20) | ((ff - 1) << 8) | (lane * 8 + t);
            crow[k] = e;
            row[k++] = e;
            const uint32_t fb = (x + B - 1) >> bwb;                     // first bucket whose records list the entry
            if (fb < 64) atomicAdd(&cnt[fb], 1u);
        }
        x += ff;
    }
    if (lane == 0) row[nnz] = 0xffffffffu;
    __syncwarp();
    const uint32_t q0 = cnt[2 * lane], q1 = cnt[2 * lane + 1];
    const uint32_t upto = warp_incl_scan(q0 + q1, lane);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t b = 2 * lane + h;
        const uint32_t kin = h ? upto : upto - q1;                      // entries starting at or before b * B
        const uint32_t bend = (b + 1) << bwb;
        uint32_t wv[8], lastv = 0;
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
        uint4 *dst = (uint4 *)(recs_row + ((size_t)b << 3));
        dst[0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        dst[1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
    }
    __syncwarp();
    return 0;
}

__device__ inline void dec_table_stage(DecPrep &P, DecTabSmem &S) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t nsym = P.nsym, ns1 = nsym + 1, shift = P.shift;
    const uint32_t ent_bytes = (nsym * ns1 * 4 + 31) & ~31u;
    uint32_t *ent = (uint32_t *)P.tb;
    uint32_t *recs = (uint32_t *)(P.tb + ent_bytes);
    {   // the raw 16-bit counts of row i are parsed into the upper half of compact row i
        uint4 *z = (uint4 *)ent;
        const uint32_t nz = ent_bytes >> 4;
        for (uint32_t j = tid; j < nz; j += DT_THREADS) z[j] = make_uint4(0, 0, 0, 0);
    }
    __threadfence_block();
    __syncthreads();
    const uint8_t *table_end = nullptr;
    int err = cta_parse_o1_rows(P.rows, P.rows_end, nsym, 1u << shift, (uint16_t *)ent + ns1, 2 * ns1, &table_end, S);
    __threadfence_block();
    __syncthreads();
    if (!err) {
        int e2 = 0;
        for (uint32_t i = wid; i < nsym; i += DT_WARPS)
            e2 |= dec_big_row((const uint16_t *)(ent + (size_t)i * ns1) + ns1, ent + (size_t)i * ns1,
                              recs + ((size_t)i << 9), nsym, shift, S.cnt[wid], S.crow[wid], lane);
        err = __syncthreads_or(e2);
    }
    if (!err) {
        const uint8_t *pay = P.pay ? P.pay : table_end;
        if ((uint32_t)(P.end - pay) < P.N * 4) err = 1;
        else if (tid == 0) P.pay = pay;
    }
    if (err && tid == 0) { P.state = 2; P.status = ST_FAIL; }
}

// ------------------------------------------------------------------------ chain (one warp)
struct __align__(16) DecChainSmem {
    uint8_t ring[RING];
    uint8_t sym[256];
};
template <int N>
__device__ inline int dec_chain_big(const DecPrep &P, DecChainSmem &S, int lane) {
    const uint8_t *cp = P.pay, *end = P.end;
    const uint32_t shift = P.shift, nsym = P.nsym, ns1 = nsym + 1, out_sz = P.t1_size;
    const uint32_t ent_bytes = (nsym * ns1 * 4 + 31) & ~31u;
    ((uint2 *)S.sym)[lane] = ((const uint2 *)P.sym)[lane];
    const bool act = lane < N;
    uint32_t R = RANS_L;
    if (act) {
        const uint8_t *p = cp + 4 * lane;
        R = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
    }
    if (__any_sync(FULL, R < RANS_L)) return 1;
    WordRing w;
    __syncwarp();
    w.init(cp, 4 * N, (uint32_t)(end - cp), S.ring, lane);
    __syncwarp();
    const uint32_t lt = lanemask_lt();
    const uint32_t seg = out_sz / N, mask = (1u << shift) - 1;
    uint8_t *out = P.t1;
    uint8_t *o = out + (size_t)lane * seg;
    uint32_t ctx = 0;                       // symbol 0 is listed, so its rank is 0: every lane starts there
    const DecO1Big B{(const uint32_t *)(P.tb + ent_bytes), (const uint32_t *)P.tb, ns1, shift};
    auto step = [&](bool on) {
        const uint32_t m = R & mask;
        const uint32_t e = B.look(m, ctx);
        const uint32_t r = e & 0xff, c0 = e >> 20, f = ((e >> 8) & 0xfff) + 1;
        if (on) {
            R = f * (R >> shift) + m - c0;
            ctx = r;
        }
        return S.sym[r];
    };
    uint32_t k = 0;
    if (N == 32 && ((((uintptr_t)out) | seg) & 15) == 0) {
        const uint32_t sy_s = (uint32_t)__cvta_generic_to_shared(S.sym);
        if (w.pos & 1) dec_o1_fast_big<true>(R, ctx, k, seg, o, w, B, sy_s, lane, lt);
        else dec_o1_fast_big<false>(R, ctx, k, seg, o, w, B, sy_s, lane, lt);
    }
    for (; k < seg; k++) {
        const uint8_t s = step(act);
        if (act) o[k] = s;
        R = renorm_step(R, act, w, lane, lt);
    }
    const bool last = lane == N - 1;
    for (uint32_t k2 = seg * N; k2 < out_sz; k2++) {
        const uint8_t s = step(last);
        if (last) out[k2] = s;
        R = renorm_step(R, last, w, lane, lt);
    }
    w.drain();
    return 0;
}

// ------------------------------------------------------------------------ post (one CTA of 256 threads)
constexpr int DP_THREADS = 256, DP_WARPS = 8;
constexpr uint32_t DP_PER = 16, DP_TILE = DP_THREADS * DP_PER, DP_STAGE = 8192;
struct __align__(16) DecPostSmem {
    uint32_t lut[256];          // unpack: byte -> symbols
    uint8_t  isr[256];          // RLE: symbol is run-length coded
    uint32_t wsum[DP_WARPS], wsum2[DP_WARPS];
    uint32_t carry[4];
    uint32_t bad;
    uint32_t pad_[3];
    uint8_t  stage[DP_STAGE + 16];      // RLE: the output of one round of literals (16-byte aligned)
};

// exclusive scan of one value per thread over the CTA; returns the thread's offset, *total the sum
__device__ __forceinline__ uint32_t cta_excl_scan(uint32_t v, uint32_t *wsum, uint32_t *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t incl = warp_incl_scan(v, lane);
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < DP_WARPS; w++) { const uint32_t c = wsum[w]; if (w < wid) before += c; tot += c; }
    __syncthreads();
    *total = tot;
    return before + incl - v;
}

// Run-length expansion (rle.c:142-189) by a CTA.  Pass 1 over the run-length bytes: a bitmap of the bytes that
// end a varint and, per 32 bytes, how many varints end before them -- so that "the k-th varint" is a search
// and a bit select instead of a walk.  Pass 2 over the literals, 4096 per round: every thread takes 16, counts
// those that carry a run, finds its first varint, decodes forward from there; a scan over the lengths gives the
// output offsets.  Fails where warp_rle_decode fails (output overflow, an unterminated varint that a literal asks
// for) and, unlike it, on a varint of more than five bytes (no encoder writes one).
__device__ inline bool cta_rle_decode(const uint8_t *lit, uint32_t lit_len, const uint8_t *run, uint32_t run_len,
                                      const uint8_t *syms, uint32_t nsyms, uint8_t *out, uint32_t *out_len,
                                      uint8_t *scr, DecPostSmem &S) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t lt = lanemask_lt();
    if (tid < 64) ((uint32_t *)S.isr)[tid] = 0;
    if (tid == 0) S.bad = 0;
    __syncthreads();
    for (uint32_t j = tid; j < nsyms; j += DP_THREADS) S.isr[syms[j]] = 1;
    const uint32_t cap = *out_len;
    const uint32_t G = (run_len + 31) >> 5;
    uint32_t *bm = (uint32_t *)scr, *gp = bm + G + 1;           // gp[g] = varints ending before byte 32 g; gp[G] = all
    // ---- pass 1
    int bad = 0;
    for (uint32_t g0 = 0; g0 < G; g0 += DP_THREADS) {           // a warp per 32 bytes, 256 groups per round
        uint32_t mine = 0;                                      // thread t keeps the count of group g0 + t
        for (uint32_t q = 0; q < 32; q++) {
            const uint32_t g = g0 + 32 * wid + q;               // warp w: groups g0 + 32 w .. + 31
            if (g >= G) break;                                  // (uniform per warp)
            const uint32_t p = 32 * g + lane;
            const uint32_t b = p < run_len ? run[p] : 0x80u;
            const uint32_t T = __ballot_sync(FULL, p < run_len && !(b & 0x80));
            // five continuation bytes in a row (looking back over the group boundary) cannot be a varint
            const uint32_t Cn = __ballot_sync(FULL, p < run_len && (b & 0x80));
            uint32_t prev4 = 0;
            if (g && lane < 4) prev4 = run[32 * g - 4 + lane] & 0x80 ? 1u : 0u;
            const uint32_t P4 = __ballot_sync(FULL, prev4 != 0) & 15;           // bit j: byte 32g-4+j continues
            const uint64_t C64 = ((uint64_t)Cn << 4) | P4;
            if (C64 & (C64 >> 1) & (C64 >> 2) & (C64 >> 3) & (C64 >> 4)) bad = 1;
            if (lane == 0) bm[g] = T;
            if ((uint32_t)lane == q) mine = __popc(T);
        }
        __syncthreads();
        uint32_t tot;
        const uint32_t ex = cta_excl_scan(mine, S.wsum, &tot);
        const uint32_t carry = g0 ? S.carry[0] : 0;
        if (g0 + tid < G) gp[g0 + tid] = carry + ex;
        __syncthreads();
        if (tid == 0) S.carry[0] = carry + tot;
        __syncthreads();
    }
    const uint32_t V = G ? S.carry[0] : 0;                      // varints in the stream
    if (tid == 0) { gp[G] = V; bm[G] = 0; }
    __threadfence_block();
    if (__syncthreads_or(bad)) return false;
    // bytes behind the last terminator: an unfinished varint
    uint32_t last_end = 0;                                      // one past the last terminator
    if (V) {
        // (uniform) the last group with a terminator: from the back
        uint32_t g = G;
        while (g-- > 0) if (bm[g]) { last_end = 32 * g + (32 - __clz(bm[g])); break; }
    }
    const bool partial = last_end < run_len;
    // ---- pass 2
    const bool lit_al = (((uintptr_t)lit) & 15) == 0;
    uint32_t K = 0, op = 0;                                     // varints consumed / bytes written before this round
    for (uint32_t base = 0; base < lit_len; base += DP_TILE) {
        const uint32_t i0 = base + DP_PER * (uint32_t)tid;
        const uint32_t nlit = i0 >= lit_len ? 0u : (lit_len - i0 < DP_PER ? lit_len - i0 : DP_PER);
        uint32_t c[DP_PER];
        if (lit_al && nlit == DP_PER) {
            const uint4 q = *(const uint4 *)(lit + i0);         // (written by the chain launch: plain load)
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < (int)DP_PER; k++) c[k] = (w4[k >> 2] >> (8 * (k & 3))) & 0xff;
        } else {
#pragma unroll
            for (int k = 0; k < (int)DP_PER; k++) c[k] = (uint32_t)k < nlit ? lit[i0 + k] : 0;
        }
        uint32_t rm = 0;                                        // bit k: literal k carries a run
#pragma unroll
        for (int k = 0; k < (int)DP_PER; k++) if ((uint32_t)k < nlit && S.isr[c[k]]) rm |= 1u << k;
        uint32_t nrun_tot;
        const uint32_t kfirst = K + cta_excl_scan(__popc(rm), S.wsum, &nrun_tot);
        // the thread's varints: kfirst, kfirst + 1, ... ; values
        uint32_t val[DP_PER];
        uint32_t lensum = nlit;
        int tbad = 0;
#pragma unroll
        for (int k = 0; k < (int)DP_PER; k++) val[k] = 0;
        if (rm) {
            uint32_t pos = 0;                                   // next byte of the run stream for this thread
            bool located = false;
            uint32_t kk = kfirst;
#pragma unroll
            for (int k = 0; k < (int)DP_PER; k++) {
                uint32_t v = 0;
                if ((rm >> k) & 1) {
                    if (kk >= V) {                              // stream exhausted: reads as 0 ...
                        if (partial) tbad = 1;                  // ... unless an unfinished varint is left
                    } else {
                        if (!located) {
                            // group holding the end of varint kk: gp[g] <= kk < gp[g + 1]
                            uint32_t lo = 0, hi = G;            // invariant: gp[lo] <= kk < gp[hi]
                            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (gp[mid] <= kk) lo = mid; else hi = mid; }
                            const uint32_t e = 32 * lo + __fns(bm[lo], 0, kk - gp[lo] + 1);
                            uint32_t s2 = e;
                            while (s2 > 0 && e - s2 < 4 && (run[s2 - 1] & 0x80)) s2--;
                            pos = s2;
                            located = true;
                        }
                        uint32_t b;
                        do { b = run[pos++]; v = (v << 7) | (b & 0x7f); } while (b & 0x80);
                        kk++;
                    }
                }
                val[k] = v;
                lensum += v;
            }
        }
        uint32_t round_out;
        const uint32_t o0 = op + cta_excl_scan(lensum, S.wsum2, &round_out);
        // the reference's checks: outp >= out_end before each literal, outp + rlen >= out_end for runs
        if (nlit) {
            if (!rm) { if ((uint64_t)o0 + nlit > cap) tbad = 1; }
            else {
                uint32_t o = o0;
#pragma unroll
                for (int k = 0; k < (int)DP_PER; k++)
                    if ((uint32_t)k < nlit) {
                        if (o >= cap || (val[k] && (uint64_t)o + val[k] >= cap)) tbad = 1;
                        o += val[k] + 1;
                    }
            }
        }
        if (__syncthreads_or(tbad)) return false;
        // A round that expands to at most DP_STAGE bytes is put together in shared memory and leaves as 16-byte
        // stores; a larger one (long runs) is written in place.
        const bool staged = round_out <= DP_STAGE;
        const uint32_t sa = (uint32_t)((uintptr_t)(out + op) & 15);   // the stage keeps the output's alignment mod 16
        uint8_t *dst = staged ? S.stage + sa - op : out;        // byte x of the output goes to dst[x]
        if (nlit) {
            if (!rm) {
#pragma unroll
                for (int k = 0; k < (int)DP_PER; k++) if ((uint32_t)k < nlit) dst[o0 + k] = (uint8_t)c[k];
            } else {
                uint32_t o = o0;
#pragma unroll
                for (int k = 0; k < (int)DP_PER; k++)
                    if ((uint32_t)k < nlit) {
                        const uint32_t len = val[k] + 1;
                        dst[o] = (uint8_t)c[k];
                        if (len <= 32) for (uint32_t q = 1; q < len; q++) dst[o + q] = (uint8_t)c[k];
                        o += len;
                    }
            }
        }
        // long runs: the warp writes them together
        if (__any_sync(FULL, rm != 0)) {
            uint32_t o2 = o0;
#pragma unroll
            for (int k = 0; k < (int)DP_PER; k++) {
                const uint32_t len = (uint32_t)k < nlit ? val[k] + 1 : 0;
                uint32_t big = __ballot_sync(FULL, len > 32);
                while (big) {
                    const int l = __ffs(big) - 1;
                    big &= big - 1;
                    const uint32_t oo = __shfl_sync(FULL, o2, l), ll = __shfl_sync(FULL, len, l), cc = __shfl_sync(FULL, c[k], l);
                    for (uint32_t q = 1 + lane; q < ll; q += 32) dst[oo + q] = (uint8_t)cc;
                }
                o2 += len;
            }
        }
        if (staged) {
            __syncthreads();
            uint8_t *g = out + op;
            const uint8_t *sp = S.stage + sa;
            uint32_t head = (16 - sa) & 15;
            if (head > round_out) head = round_out;
            if ((uint32_t)tid < head) g[tid] = sp[tid];
            const uint32_t nv = (round_out - head) >> 4;
            for (uint32_t i = tid; i < nv; i += DP_THREADS) *(uint4 *)(g + head + 16 * i) = *(const uint4 *)(sp + head + 16 * i);
            for (uint32_t i = head + (nv << 4) + tid; i < round_out; i += DP_THREADS) g[i] = sp[i];
            __syncthreads();
        }
        K += nrun_tot;
        op += round_out;
    }
    *out_len = op;
    __threadfence_block();
    __syncthreads();
    return true;
}

// pack.c:207-344 by a CTA (same cases as warp_unpack)
__device__ inline bool cta_unpack(const uint8_t *src, uint32_t len, uint8_t *dst, uint32_t out_len, const PackMap &pm,
                                  DecPostSmem &S) {
    const int tid = threadIdx.x;
    const int per = pm.per;
    if (per == 1) {
        for (uint32_t i = tid; i < len; i += DP_THREADS) dst[i] = src[i];
        return true;
    }
    if (per == 0) {
        for (uint32_t i = tid; i < out_len; i += DP_THREADS) dst[i] = pm.map[0];
        return true;
    }
    if ((out_len + per - 1) / per > len) return false;
    const uint32_t bits = 8 / per, cm = (1u << bits) - 1;
    const uint32_t whole = out_len / per;
    if (per == 4 && (((uintptr_t)dst) & 3) == 0) {
        S.lut[tid] = pm.map[tid & 3] | (pm.map[(tid >> 2) & 3] << 8) | (pm.map[(tid >> 4) & 3] << 16) |
                     ((uint32_t)pm.map[(tid >> 6) & 3] << 24);
        __syncthreads();
        const uint32_t *lut = S.lut;
        uint32_t *d4 = (uint32_t *)dst;
        uint32_t j0 = 0;
        if (((((uintptr_t)src) & 3) | (((uintptr_t)dst) & 15)) == 0) {
            const uint32_t *s4 = (const uint32_t *)src;
            uint4 *d16 = (uint4 *)dst;
            const uint32_t nq = whole >> 2;
            auto expand = [&](uint32_t w) {
                return make_uint4(lut[w & 0xff], lut[(w >> 8) & 0xff], lut[(w >> 16) & 0xff], lut[w >> 24]);
            };
            uint32_t i = tid;
            for (; i + 3 * DP_THREADS < nq; i += 4 * DP_THREADS) {          // four loads in flight per thread
                const uint32_t w0 = s4[i], w1 = s4[i + DP_THREADS], w2 = s4[i + 2 * DP_THREADS], w3 = s4[i + 3 * DP_THREADS];
                d16[i] = expand(w0); d16[i + DP_THREADS] = expand(w1);
                d16[i + 2 * DP_THREADS] = expand(w2); d16[i + 3 * DP_THREADS] = expand(w3);
            }
            for (; i < nq; i += DP_THREADS) d16[i] = expand(s4[i]);
            j0 = nq << 2;
        }
        for (uint32_t j = j0 + tid; j < whole; j += DP_THREADS) d4[j] = lut[src[j]];
    } else if (per == 2 && (((uintptr_t)dst) & 1) == 0) {
        uint16_t *lut = (uint16_t *)S.lut;
        lut[tid] = (uint16_t)(pm.map[tid & 15] | (pm.map[tid >> 4] << 8));
        __syncthreads();
        uint16_t *d2 = (uint16_t *)dst;
        uint32_t j0 = 0;
        if (((((uintptr_t)src) & 3) | (((uintptr_t)dst) & 7)) == 0) {
            const uint32_t *s4 = (const uint32_t *)src;
            uint2 *d8 = (uint2 *)dst;
            const uint32_t nq = whole >> 2;
            for (uint32_t i = tid; i < nq; i += DP_THREADS) {
                const uint32_t w = s4[i];
                d8[i] = make_uint2(lut[w & 0xff] | ((uint32_t)lut[(w >> 8) & 0xff] << 16),
                                   lut[(w >> 16) & 0xff] | ((uint32_t)lut[w >> 24] << 16));
            }
            j0 = nq << 2;
        }
        for (uint32_t j = j0 + tid; j < whole; j += DP_THREADS) d2[j] = lut[src[j]];
    } else {
        for (uint32_t j = tid; j < whole; j += DP_THREADS) {
            uint32_t c = src[j];
            for (int q = 0; q < per; q++) { dst[j * per + q] = pm.map[c & cm]; c >>= bits; }
        }
    }
    const uint32_t done = whole * per;
    if (done < out_len && tid == 0) {
        uint32_t c = src[whole];
        for (uint32_t i = done; i < out_len; i++) { dst[i] = pm.map[c & cm]; c >>= bits; }
    }
    return true;
}

__device__ inline void dec_post_stage(DecJob &J, DecPrep &P, DecPostSmem &S) {
    const int tid = threadIdx.x;
    int status = ST_OK;
    uint32_t out_size = P.out_size;
    do {
        const uint32_t flag = P.flag;
        uint32_t t2_size = P.t1_size, t3_size = P.t1_size;
        if (flag & X_RLE) {                                             // rANS_static4x16pr.c:1856-1871
            const uint8_t *meta = P.meta;
            const uint32_t nsyms = *meta ? *meta : 256;
            uint32_t unrle = out_size;
            if (!cta_rle_decode(P.t1, P.t1_size, meta + 1 + nsyms, P.u_meta - (1 + nsyms), meta + 1, nsyms, P.t2,
                                &unrle, P.scr, S)) { status = ST_FAIL; break; }
            t3_size = t2_size = unrle;
        }
        if (flag & X_PACK) {                                            // :1872-1881
            uint32_t unpacked_sz = P.unpacked_sz;
            if (P.pm.per == 1) unpacked_sz = t2_size;
            if (!cta_unpack(P.t2, t2_size, P.t3, unpacked_sz, P.pm, S)) { status = ST_FAIL; break; }
            t3_size = unpacked_sz;
        }
        out_size = t3_size;
    } while (0);
    if (tid == 0) { J.status = status; J.out_size = status == ST_OK ? out_size : 0; }
}

}  // namespace b200
