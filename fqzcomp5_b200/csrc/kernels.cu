// kernels.cu -- the __global__ entry points: container logic of
// rans_compress_to_4x16 / rans_uncompress_to_4x16 (rANS_static4x16pr.c:1224-1894)
// around the coders in rans_encode.cuh / rans_decode.cuh, plus the size scan and
// the gather that packs finished streams back to back.
#include <stdlib.h>
#include <atomic>
#include "kernels.h"
#include "rans_decode.cuh"
#include "rans_encode.cuh"
#include "transforms.cuh"
#include "prep.cuh"

namespace b200 {

// ------------------------------------------------------------------------
// Encode one stream (everything except STRIPE, which the host expands into
// NOSZ sub-streams).  Mirrors the decisions of rANS_static4x16pr.c:1256-1579.
// Capacity checks of the reference are evaluated against J.cap; need_cap
// records the smallest capacity for which this call succeeds.
// ------------------------------------------------------------------------
struct CapCheck {
    uint32_t cap, need;
    bool ok;
    __device__ void require(uint64_t v) {
        if (v > 0xfffffff0ull) v = 0xfffffff0ull;
        if (v > need) need = (uint32_t)v;
        if (v > cap) ok = false;
    }
};

// inslot: leave the finished stream contiguous inside its slot (head moved up against the payload,
// as the reference's memmove at rANS_static32x16pr.c:249-251 does the other way round) and report it
// as tail / tail_len with head_len 0, so that no packing pass is needed.
// PREPPED: the stream went through prep_kernel (prep.cuh); the in-warp transforms and model building are
// compiled out, which leaves a kernel with fewer registers and less shared memory per stream.
template <bool O1, bool PREPPED = false>
__device__ void enc_stream(EncJob &J, uint8_t *smem, uint32_t smem_bytes, const Pool &pool, int lane, bool inslot) {
    int order = J.order;
    const uint8_t *in = J.in;
    uint32_t in_size = J.in_size;
    uint8_t *out = J.slot;
    // the order-0 kernel gives every warp a multiple of 1 KiB of shared memory: its output ring is aligned (OutRingT)
    constexpr bool AL = !O1;
    // PACK / RLE streams: transforms, counts and the order-1 model may have been done by prep_kernel (prep.cuh)
    const Prep *P = (J.prep && ((const Prep *)J.prep)->state == 1) ? (const Prep *)J.prep : nullptr;
    CapCheck cc{J.cap, 1, J.cap != 0};            // (out && *out_size == 0) -> NULL (:1227)
    uint32_t status = (PREPPED && !P) ? ST_UNSUPPORTED : ST_OK;      // routed here without a prep area: host bug
    uint32_t head_len = 0, tail_len = 0;
    const uint8_t *tail = nullptr;

    if ((order & ORDER_SIMD_AUTO) && in_size >= 50000 && !(order & X_STRIPE)) order |= X_32;
    if (in_size <= 20) order &= ~X_STRIPE;
    if (in_size <= 1000) order &= ~X_32;

    if (in_size > 0x7fffffffu || (order & X_STRIPE)) {
        status = (order & X_STRIPE) ? ST_UNSUPPORTED : ST_FAIL;
    } else if (order & X_CAT) {                                           // :1395-1409
        uint32_t m = 1;
        if (lane == 0) { out[0] = X_CAT; m += var_put_u32(out + 1, in_size); }
        m = __shfl_sync(FULL, m, 0);
        cc.require((uint64_t)m + in_size);
        head_len = m; tail = in; tail_len = in_size;
    } else {
        const int do_pack = order & X_PACK, no_size = order & X_NOSZ;
        int do_rle = order & X_RLE, do_simd = order & X_32;
        uint32_t flag = order & 0xff;
        uint32_t meta = 1, szq = 0;                 // szq: "*out_size -= sz" at :1453
        if (!no_size) {
            if (lane == 0) meta += var_put_u32(out + 1, in_size);
            meta = __shfl_sync(FULL, meta, 0);
        }
        int o1 = order & 1;
        uint8_t *work = J.work;
        // counts from hist_kernel describe the caller's bytes: unusable once PACK/RLE rewrote them
        const uint32_t *model = (J.model && in == J.in) ? J.model : nullptr;

        if ((do_pack || do_rle) && in_size && !work) status = ST_UNSUPPORTED;   // host sizes work for these
        if (status != ST_OK) {
        } else if (do_pack && in_size) {                                  // :1429-1459
            cc.require((uint64_t)meta + 256);
            uint32_t pmeta = 0, plen = 0;
            bool packed;
            if (PREPPED || P) { packed = P->packed != 0; pmeta = P->pmeta; plen = P->plen; }
            else if constexpr (!PREPPED) packed = work && warp_pack(in, in_size, out + meta, &pmeta, work, &plen, smem, lane);
            if (!packed) flag &= ~X_PACK;
            else {
                in = work; work += (plen + 15) & ~15u;
                in_size = plen;
                meta += pmeta;
                uint32_t sz = 0;
                if (lane == 0) sz = var_put_u32(out + meta, in_size);
                sz = __shfl_sync(FULL, sz, 0);
                meta += sz; szq = sz;
                if (do_simd && in_size < 32) { do_simd = 0; flag &= ~X_32; }
            }
        } else if (do_pack) flag &= ~X_PACK;

        if (status != ST_OK) {
        } else if (do_rle && in_size) {                                   // :1464-1533
            // work: [literals in_size][meta in_size+257+16]
            uint8_t *lits = work, *rmeta = work + ((in_size + 15) & ~15u);
            uint32_t rle_len = 0, rmeta_len = 0;
            if (PREPPED || P) { rle_len = P->rle_len; rmeta_len = P->rmeta_len; }
            else if constexpr (!PREPPED) warp_rle_encode(in, in_size, lits, &rle_len, rmeta, &rmeta_len, smem, lane);
            if ((double)((uint64_t)rle_len + rmeta_len) >= .99 * (double)in_size) {
                flag &= ~X_RLE; do_rle = 0;
            } else {
                uint32_t sz = 0;
                if (lane == 0) {
                    sz = var_put_u32(out + meta, rmeta_len * 2);
                    sz += var_put_u32(out + meta + sz, rle_len);
                }
                sz = __shfl_sync(FULL, sz, 0);
                cc.require((uint64_t)meta + sz + 5 + szq);
                if (do_simd && (rmeta_len < 32 || rle_len < 32)) { do_simd = 0; flag &= ~X_32; }
                // the run-length bytes go through the order-0 coder (same lane count)
                cc.require((uint64_t)compress_bound(rmeta_len, 0) - 20 + meta + sz + 5 + szq);
                uint8_t *tmp = rmeta + ((rmeta_len + 15) & ~15u);        // scratch for the coded meta
                uint32_t cb = (compress_bound(rmeta_len, 0) - 20) & ~1u, ctab = 0;
                uint8_t *cptr = nullptr;
                int e = do_simd ? enc_o0<32, AL>(rmeta, rmeta_len, tmp, tmp + cb, &ctab, &cptr, *(EncO0Smem *)smem, lane)
                                : enc_o0<4, AL>(rmeta, rmeta_len, tmp, tmp + cb, &ctab, &cptr, *(EncO0Smem *)smem, lane);
                if (e) status = ST_FAIL;
                uint32_t pay = (uint32_t)(tmp + cb - cptr), c_rmeta = ctab + pay, sz2 = 0;
                if (!e && c_rmeta < rmeta_len) {
                    if (lane == 0) sz2 = var_put_u32(out + meta + sz, c_rmeta);
                    sz2 = __shfl_sync(FULL, sz2, 0);
                    __syncwarp();
                    warp_copy(out + meta + sz + sz2, tmp, ctab, lane);
                    warp_copy(out + meta + sz + sz2 + ctab, cptr, pay, lane);
                } else if (!e) {                                          // raw: too small to pay off
                    if (lane == 0) {
                        sz = var_put_u32(out + meta, rmeta_len * 2 + 1);
                        sz2 = var_put_u32(out + meta + sz, rle_len);
                    }
                    sz = __shfl_sync(FULL, sz, 0); sz2 = __shfl_sync(FULL, sz2, 0);
                    __syncwarp();
                    warp_copy(out + meta + sz + sz2, rmeta, rmeta_len, lane);
                    c_rmeta = rmeta_len;
                }
                meta += sz + sz2 + c_rmeta;
                in = lits; in_size = rle_len;
            }
        } else if (do_rle) flag &= ~X_RLE;

        if (in != J.in) model = nullptr;
        if (P && P->model) model = P->F;                                  // counts of the data as it is now
        cc.require((uint64_t)meta + szq);                                 // :1538
        if (o1 && in_size < 8) { flag &= ~1u; o1 = 0; }                   // :1547
        cc.require((uint64_t)compress_bound(in_size, o1) - 20 + meta + szq);   // bound > *out_size in the coder

        uint32_t tab = 0;
        uint8_t *ptr = nullptr, *oend = out + (J.slot_cap & ~1u);
        int e = 0;
        if (status == ST_OK) {
            __syncwarp();
            if (O1 && o1) {
                EncO1Smem &S = *(EncO1Smem *)smem;
                uint8_t *dyn = smem + sizeof(EncO1Smem);
                uint32_t dynb = smem_bytes - (uint32_t)sizeof(EncO1Smem);
                if (P && P->model == 2)
                    e = do_simd ? enc_o1_prepped<32>(in, in_size, out + meta, oend, &tab, &ptr, smem, smem_bytes, *P, J.prep, J.in_size, lane)
                                : enc_o1_prepped<4>(in, in_size, out + meta, oend, &tab, &ptr, smem, smem_bytes, *P, J.prep, J.in_size, lane);
                else if constexpr (PREPPED) e = 1;          // N == 32 with fewer than 32 bytes: the reference fails too
                else {
                    if (P) model = nullptr;       // the order-1 coder's own model layout differs from Prep::F
                    e = do_simd ? enc_o1<32>(in, in_size, out + meta, oend, &tab, &ptr, S, dyn, dynb, pool, lane, model)
                                : enc_o1<4>(in, in_size, out + meta, oend, &tab, &ptr, S, dyn, dynb, pool, lane, model);
                }
            } else if (o1) {
                e = 3;      // order-1 stream routed to the order-0-only kernel: host bug
            } else {
                EncO0Smem &S = *(EncO0Smem *)smem;
                e = do_simd ? enc_o0<32, AL>(in, in_size, out + meta, oend, &tab, &ptr, S, lane, model)
                            : enc_o0<4, AL>(in, in_size, out + meta, oend, &tab, &ptr, S, lane, model);
            }
            if (e) status = (e == 2 || e == 3) ? ST_UNSUPPORTED : ST_FAIL;
        }
        if (status == ST_OK) {
            uint32_t pay = (uint32_t)(oend - ptr);
            if (tab + pay >= in_size) {                                   // :1560-1574 store raw instead
                flag &= ~3u;
                flag |= X_CAT | no_size;
                cc.require((uint64_t)meta + in_size);
                head_len = meta; tail = in; tail_len = in_size;
            } else {
                head_len = meta + tab; tail = ptr; tail_len = pay;
            }
            if (lane == 0) out[0] = (uint8_t)flag;
        }
    }
    if (!cc.ok && status == ST_OK) status = ST_FAIL;
    if (inslot && status == ST_OK) {
        __syncwarp();
        if (tail >= out && tail <= out + J.slot_cap) {          // coded payload at the slot's end
            uint8_t *start = const_cast<uint8_t *>(tail) - head_len;
            warp_move_up(start, out, head_len, lane);
            tail = start;
        } else {                                                // stored raw: the bytes are still the input's
            warp_copy(out + head_len, tail, tail_len, lane);
            tail = out;
        }
        tail_len += head_len; head_len = 0;
    }
    if (lane == 0) {
        J.tail = tail; J.head_len = head_len; J.tail_len = tail_len;
        J.status = status; J.need_cap = cc.need;
    }
    __syncwarp();
}

// O1: 22 warps per SM (<= 92 registers); O0: 28 warps (<= 72) so that the 3815 streams of a 1 GB block are one
// wave (unbounded the compiler takes 179 registers and residency collapses); order-1 streams that went through
// prep_kernel: 28 warps (<= 72 registers, 7.5 KiB of shared memory each), one wave as well
template <bool O1, bool PREPPED>
__global__ void __launch_bounds__((O1 ? ENC_WARPS_O1 : ENC_WARPS) * 32, PREPPED ? 14 : O1 ? 11 : 7)
enc_kernel(EncJob *jobs, uint32_t njobs, uint32_t warp_smem, Pool pool, uint32_t route, uint32_t inslot) {
    extern __shared__ __align__(1024) uint8_t smem_all[];
    static_assert(ENC_SMEM_O0 % ORING == 0, "the order-0 kernel's output rings are aligned to their size");
    if (!O1 && ((warp_smem | (uint32_t)__cvta_generic_to_shared(smem_all)) & (ORING - 1))) __trap();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t j = blockIdx.x * (O1 ? ENC_WARPS_O1 : ENC_WARPS) + wid;
    if (j >= njobs || jobs[j].route != route) return;   // another launch's stream, or a STRIPE parent
    enc_stream<O1, PREPPED>(jobs[j], smem_all + (size_t)wid * warp_smem, warp_smem, pool, lane, inslot != 0);
}

// ------------------------------------------------------------------------
// Decode one stream (everything except STRIPE: the host splits those into
// their NOSZ sub-streams and un-stripes afterwards).  rANS_static4x16pr.c:1696-1887.
// ------------------------------------------------------------------------
template <bool O1>
__device__ void dec_stream(DecJob &J, uint8_t *smem, uint32_t smem_bytes, const Pool &pool, int lane) {
    const uint8_t *in = J.in, *in_end = J.in + J.in_size;
    uint32_t in_size = J.in_size;
    int status = ST_OK;
    uint32_t out_size = 0;
    do {
        if (in_size == 0) { status = ST_FAIL; break; }
        const int flag = *in++; in_size--;
        if (flag & X_STRIPE) { status = ST_UNSUPPORTED; break; }
        const int do_pack = flag & X_PACK, do_rle = flag & X_RLE, do_cat = flag & X_CAT,
                  no_size = flag & X_NOSZ, do_simd = flag & X_32, o1 = flag & 1;
        uint32_t osz = J.out_cap;
        if (!no_size) { int sz = var_get_u32(in, in_end, &osz); in += sz; in_size -= sz; }
        if (J.out_cap < osz) { status = ST_FAIL; break; }
        out_size = osz;
        uint8_t *out = J.out, *tmp = J.tmp;
        if ((do_pack || do_rle) && !tmp) { status = ST_UNSUPPORTED; break; }
        uint8_t *t1, *t2, *t3;                                          // :1760-1782
        if (do_pack && do_rle) { t1 = out; t2 = tmp; t3 = out; }
        else if (do_pack)      { t1 = tmp; t2 = tmp; t3 = out; }
        else if (do_rle)       { t1 = tmp; t2 = out; t3 = out; }
        else                   { t1 = t2 = t3 = out; }
        uint32_t t1_size = out_size;

        PackMap pm;                                                      // :1788-1806
        uint32_t unpacked_sz = 0;
        if (do_pack) {
            int c = unpack_meta(in, in_size, pm);
            if (!c) { status = ST_FAIL; break; }
            unpacked_sz = osz;
            in += c; in_size -= c;
            uint32_t psz;
            int sz = var_get_u32(in, in_end, &psz);
            in += sz; in_size -= sz;
            if (psz > t1_size) { status = ST_FAIL; break; }
            t1_size = psz;
        }
        const uint8_t *meta = nullptr;
        uint32_t u_meta = 0;
        if (do_rle) {                                                    // :1810-1834
            uint32_t c_meta, rle_len;
            uint32_t sz = var_get_u32(in, in_end, &u_meta);
            sz += var_get_u32(in + sz, in_end, &rle_len);
            if (rle_len > t1_size) { status = ST_FAIL; break; }
            if (u_meta & 1) {
                meta = in + sz;
                uint32_t left = (uint32_t)(in_end - meta);
                u_meta = u_meta / 2 > left ? left : u_meta / 2;
                c_meta = u_meta;
            } else {
                sz += var_get_u32(in + sz, in_end, &c_meta);
                u_meta /= 2;
                // run lengths never outgrow the data they describe
                if (u_meta > J.out_cap + 1024 || in_size < sz) { status = ST_FAIL; break; }
                uint8_t *mbuf = tmp + ((J.out_cap + 15) & ~15u);
                int e = do_simd ? dec_o0<32>(in + sz, in_size - sz, mbuf, u_meta, *(DecO0Smem *)smem, lane)
                                : dec_o0<4>(in + sz, in_size - sz, mbuf, u_meta, *(DecO0Smem *)smem, lane);
                if (e) { status = ST_FAIL; break; }
                __syncwarp();
                meta = mbuf;
            }
            if ((uint64_t)c_meta + sz > in_size) { status = ST_FAIL; break; }
            in += c_meta + sz; in_size -= c_meta + sz;
            t1_size = rle_len;
        }
        if (in_size) {                                                   // :1838-1853
            if (do_cat) {
                if (t1_size > in_size || t1_size > out_size) { status = ST_FAIL; break; }
                warp_copy(t1, in, t1_size, lane);
            } else {
                int e;
                if (O1 && o1) {
                    DecO1Smem &S = *(DecO1Smem *)smem;
                    uint8_t *dyn = smem + sizeof(DecO1Smem);
                    uint32_t tb = smem_bytes - (uint32_t)sizeof(DecO1Smem);
                    e = do_simd ? dec_o1<32>(in, in_size, t1, t1_size, S, dyn, tb, (DecO0Smem *)smem, pool, lane)
                                : dec_o1<4>(in, in_size, t1, t1_size, S, dyn, tb, (DecO0Smem *)smem, pool, lane);
                } else if (o1) {
                    e = 2;
                } else {
                    DecO0Smem &S = *(DecO0Smem *)smem;
                    e = do_simd ? dec_o0<32>(in, in_size, t1, t1_size, S, lane)
                                : dec_o0<4>(in, in_size, t1, t1_size, S, lane);
                }
                if (e) { status = e == 2 ? ST_UNSUPPORTED : ST_FAIL; break; }
            }
        } else t1_size = 0;
        __syncwarp();
        uint32_t t2_size = t1_size, t3_size = t1_size;
        if (do_rle) {                                                    // :1856-1871
            if (u_meta == 0) { status = ST_FAIL; break; }
            uint32_t nsyms = *meta ? *meta : 256;
            if (u_meta < 1 + nsyms) { status = ST_FAIL; break; }
            uint32_t unrle = out_size;
            if (!warp_rle_decode(t1, t1_size, meta + 1 + nsyms, u_meta - (1 + nsyms), meta + 1, nsyms,
                                 t2, &unrle, smem, lane)) { status = ST_FAIL; break; }
            t3_size = t2_size = unrle;
        }
        if (do_pack) {                                                   // :1872-1881
            if (pm.per == 1) unpacked_sz = t2_size;
            if (!warp_unpack(t2, t2_size, t3, unpacked_sz, pm, smem, lane)) { status = ST_FAIL; break; }
            t3_size = unpacked_sz;
        }
        out_size = t3_size;
    } while (0);
    __syncwarp();
    if (lane == 0) { J.status = status; J.out_size = status == ST_OK ? out_size : 0; }
}

template <bool O1>
__global__ void __launch_bounds__((O1 ? DEC_WARPS_O1 : DEC_WARPS) * 32)
dec_kernel(DecJob *jobs, uint32_t njobs, uint32_t warp_smem, Pool pool) {
    extern __shared__ __align__(8192) uint8_t smem_dec[];
    uint8_t *smem_all = smem_dec;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t j = blockIdx.x * (O1 ? DEC_WARPS_O1 : DEC_WARPS) + wid;
    if (j >= njobs || jobs[j].route != (O1 ? 1u : 0u)) return;
    // order-0 blocks are 8 KiB each: when the dynamic shared memory itself starts 8 KiB aligned
    // (checked at run time in dec_o0) the decoder uses OR-addressing into its tables
    dec_stream<O1>(jobs[j], smem_all + (size_t)wid * warp_smem, warp_smem, pool, lane);
}

// ------------------------------------------------------------------------
// Packing: exclusive scan of the finished stream sizes (one CTA), then one CTA
// per stream copies head and tail to their final place.
// ------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
scan_kernel(const EncJob *jobs, uint32_t njobs, uint64_t *out_off, uint32_t *out_size, uint64_t *total,
            uint32_t align) {
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < njobs; base += 1024) {
        uint32_t k = base + threadIdx.x;
        uint32_t sz = 0;
        const bool item = k < njobs && jobs[k].item != 0xffffffffu;
        if (item && jobs[k].status == ST_OK) sz = jobs[k].head_len + jobs[k].tail_len;
        uint64_t v = ((uint64_t)sz + align - 1) & ~(uint64_t)(align - 1);   // streams start `align`-byte aligned (a power of two)
        uint64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t t = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += t;
        }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint64_t y = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint64_t t = __shfl_up_sync(FULL, y, o);
                if (lane >= o) y += t;
            }
            wsum[lane] = y;
        }
        __syncthreads();
        uint64_t excl = carry + (wid ? wsum[wid - 1] : 0) + x - v;
        if (item) { out_off[jobs[k].item] = excl; out_size[jobs[k].item] = sz; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256)
gather_kernel(const EncJob *jobs, uint32_t njobs, const uint64_t *out_off, uint32_t *out_size,
              uint8_t *out, uint64_t out_cap) {
    const uint32_t k = blockIdx.x;
    if (k >= njobs) return;
    const EncJob &J = jobs[k];
    if (J.status != ST_OK || J.item == 0xffffffffu) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint64_t off = out_off[J.item];
    const uint32_t hl = J.head_len, tl = J.tail_len;
    if (off + hl + tl > out_cap) { if (threadIdx.x == 0) out_size[J.item] = 0; return; }
    uint8_t *dst = out + off;
    if (wid == 0) warp_copy(dst, J.slot, hl, lane);
    if (J.stripe_n) {
        // STRIPE parent: the tail is the chosen sub-streams, back to back
        const uint32_t *list = (const uint32_t *)(J.slot + J.slot_cap);
        uint64_t o = hl;
        for (uint32_t i = 0; i < J.stripe_n; i++) {
            const EncJob &S = jobs[list[i]];
            if (wid == (int)(i % nw)) {
                warp_copy(dst + o, S.slot, S.head_len, lane);
                warp_copy(dst + o + S.head_len, S.tail, S.tail_len, lane);
            }
            o += S.head_len + S.tail_len;
        }
        return;
    }
    // tail in 16 KiB pieces spread over the warps; piece boundaries keep dst+hl alignment mod 16
    const uint32_t P = 16384;
    for (uint32_t p = wid * P; p < tl; p += nw * P) {
        uint32_t len = tl - p < P ? tl - p : P;
        warp_copy(dst + hl + p, J.tail + p, len, lane);
    }
}


// In-slot output: a STRIPE parent's chosen sub-streams are appended to its header inside its own
// slot (one CTA per parent), every other stream already sits contiguously in its slot.
__global__ void __launch_bounds__(256)
assemble_parents_kernel(EncJob *jobs, const uint32_t *parents, uint32_t nparents) {
    if (blockIdx.x >= nparents) return;
    EncJob &J = jobs[parents[blockIdx.x]];
    if (J.status != ST_OK || !J.stripe_n) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t hl = J.head_len, tl = J.tail_len, N = J.stripe_n;
    const uint32_t *list = (const uint32_t *)(J.slot + J.slot_cap);
    __syncthreads();
    if ((uint64_t)hl + tl > J.slot_cap) { if (threadIdx.x == 0) J.status = ST_FAIL; return; }
    uint64_t o = hl;
    for (uint32_t i = 0; i < N; i++) {
        const EncJob &S = jobs[list[i]];
        if (wid == (int)(i % nw)) {
            warp_copy(J.slot + o, S.slot, S.head_len, lane);
            warp_copy(J.slot + o + S.head_len, S.tail, S.tail_len, lane);
        }
        o += S.head_len + S.tail_len;
    }
    __syncthreads();
    if (threadIdx.x == 0) { J.tail = J.slot; J.tail_len = hl + tl; J.head_len = 0; }
}

// where each caller item's stream was left (in-slot output): offset from `base`, size (0 = failed)
__global__ void inslot_results_kernel(const EncJob *jobs, uint32_t njobs, const uint8_t *base, uint64_t *out_off,
                                      uint32_t *out_size) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    const EncJob &J = jobs[j];
    if (J.item == 0xffffffffu) return;
    const bool ok = J.status == ST_OK;
    out_off[J.item] = ok ? (uint64_t)(J.tail - base) : 0;
    out_size[J.item] = ok ? J.head_len + J.tail_len : 0;
}

cudaError_t launch_inslot_results(EncJob *d_jobs, uint32_t njobs, const uint32_t *d_parents, uint32_t nparents,
                                  const uint8_t *base, uint64_t *d_off, uint32_t *d_size, cudaStream_t st) {
    if (!njobs) return cudaSuccess;
    if (nparents) assemble_parents_kernel<<<nparents, 256, 0, st>>>(d_jobs, d_parents, nparents);
    inslot_results_kernel<<<(njobs + 255) / 256, 256, 0, st>>>(d_jobs, njobs, base, d_off, d_size);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------
// Histograms at full occupancy: one CTA of 8 warps per stream.  Order 0: a table of 256 symbols x 32
// columns in shared memory, lane l of every warp counting into column l -- the 32 shared-memory
// atomics of a warp instruction fall into 32 different banks whatever the bytes are, so a byte costs
// three instructions (shift, mask + base, ATOMS.POPC.INC) and the pass runs at the speed the stream
// arrives from DRAM (0.19 ms per GB against 0.43 ms for warp-private bins with equal neighbours
// merged, which is bound by its 13 instructions per byte: scripts/microbench/hist_variants.cu).
// Order-1 streams then count their (previous, current) pairs into a rank-space matrix, merging
// equal neighbours before touching shared memory.  The coder kernels start from the counts instead
// of reading the stream a second time with one warp.
// ------------------------------------------------------------------------
constexpr int HIST_THREADS = 256;
constexpr uint32_t HIST_COLS = 32;                  // one column per lane
constexpr uint32_t HIST_SMEM_PAIRS = 256 * HIST_COLS;   // words: the pair matrix reuses the order-0 table (up to 90 x 90 symbols)

// the four bytes of w, each into its row of the lane's column; base = table + 4 * lane (shared space)
__device__ __forceinline__ void hist_word_cols(uint32_t w, uint32_t base) {
    const uint32_t a0 = ((w << 7) & 0x7f80u) + base, a1 = ((w >> 1) & 0x7f80u) + base,
                   a2 = ((w >> 9) & 0x7f80u) + base, a3 = ((w >> 17) & 0x7f80u) + base;
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a0) : "memory");
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a1) : "memory");
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a2) : "memory");
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a3) : "memory");
}

__global__ void __launch_bounds__(HIST_THREADS)
hist_kernel(EncJob *jobs, uint32_t njobs) {
    const uint32_t jn = blockIdx.x;
    if (jn >= njobs) return;
    EncJob &J = jobs[jn];
    uint32_t *model = J.model;
    if (!model) return;
    __shared__ __align__(16) uint32_t Hs[HIST_SMEM_PAIRS];       // order 0: [symbol][column]; then the pair matrix
    __shared__ uint8_t rank[256], sym_of[256];
    __shared__ uint32_t wtot[8], wpres[8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint8_t *in = J.in;
    const uint32_t n = J.in_size;
    for (int j = tid; j < (int)(HIST_SMEM_PAIRS / 4); j += HIST_THREADS) ((uint4 *)Hs)[j] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    uint32_t head = (uint32_t)((16 - ((uintptr_t)in & 15)) & 15);
    if (head > n) head = n;
    const uint8_t *p = in + head;
    const uint32_t rest = n - head, nv = rest >> 4;
    const uint4 *v = (const uint4 *)p;
    {
        const uint32_t base = (uint32_t)__cvta_generic_to_shared(Hs) + 4 * lane;
        if ((uint32_t)tid < head) atomicAdd(&Hs[in[tid] * HIST_COLS + lane], 1u);
        uint32_t i = tid;
        for (; i + 7 * HIST_THREADS < nv; i += 8 * HIST_THREADS) {
            uint4 q[8];
#pragma unroll
            for (int u = 0; u < 8; u++) q[u] = ldg_u128(v + i + u * HIST_THREADS);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                hist_word_cols(q[u].x, base); hist_word_cols(q[u].y, base);
                hist_word_cols(q[u].z, base); hist_word_cols(q[u].w, base);
            }
        }
        for (; i < nv; i += HIST_THREADS) {
            const uint4 q = ldg_u128(v + i);
            hist_word_cols(q.x, base); hist_word_cols(q.y, base); hist_word_cols(q.z, base); hist_word_cols(q.w, base);
        }
        for (uint32_t t = (nv << 4) + tid; t < rest; t += HIST_THREADS) atomicAdd(&Hs[p[t] * HIST_COLS + lane], 1u);
    }
    __syncthreads();
    uint32_t f = 0;                                   // row tid, columns rotated so that a warp reads 32 banks
#pragma unroll 8
    for (int k = 0; k < (int)HIST_COLS; k++) f += Hs[tid * HIST_COLS + ((k + tid) & (HIST_COLS - 1))];
    model[tid] = f;
    const bool o1 = (J.order & 1) && n >= 8;
    if (!o1) return;                                  // uniform per CTA
    // ---- alphabet ranks (symbol 0 is always a member)
    const bool pres = f != 0 || tid == 0;
    uint32_t bal = __ballot_sync(FULL, pres);
    const uint32_t bal_data = __ballot_sync(FULL, f != 0);
    if (lane == 0) { wtot[wid] = __popc(bal); wpres[wid] = bal_data; }
    __syncthreads();
    uint32_t before = 0, nsym = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { uint32_t c = wtot[w]; if (w < wid) before += c; nsym += c; }
    rank[tid] = (uint8_t)(before + __popc(bal & lanemask_lt()));
    if (tid == 0) model[256] = nsym;
    const uint32_t hw = nsym * nsym;
    uint32_t *gH = model + MODEL_HDR_WORDS;
    // ---- pairs, small symbol range: the matrix is indexed by the bytes themselves, M[prev - lo][cur - lo] with an
    // even row length (equal neighbours -- most pairs of a quality stream -- then walk the diagonal with an odd
    // stride, i.e. over all 32 banks; with 39 symbols in rows of 39 they met in 4 banks and the pass took 2.4 times
    // as long).  One shared-memory atomic per byte, no rank look-up, no run bookkeeping: 0.32 ms per GB against
    // 0.46 ms for the rank-space form below (scripts/microbench/pairs_variants.cu).  Ranks are applied when the
    // matrix is written out.  The stream's first byte follows symbol 0 (utils.h:279-357): it is counted as its own
    // successor, taken back, and added to row 0 on the way out.
    uint32_t lo = 0, hi = 255;
#pragma unroll
    for (int w = 7; w >= 0; w--) if (wpres[w]) lo = 32 * w + __ffs(wpres[w]) - 1;
#pragma unroll
    for (int w = 0; w < 8; w++) if (wpres[w]) hi = 32 * w + 31 - __clz(wpres[w]);
    const uint32_t span = hi - lo + 1, rowlen = (span + 1) & ~1u;
    if (span * rowlen <= HIST_SMEM_PAIRS) {
        if (pres) sym_of[rank[tid]] = (uint8_t)tid;
        for (int j = tid; j < (int)(HIST_SMEM_PAIRS / 4); j += HIST_THREADS) ((uint4 *)Hs)[j] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        const uint32_t S4 = rowlen * 4;
        const uint32_t K = (uint32_t)__cvta_generic_to_shared(Hs) - lo * S4 - lo * 4;
        auto pair1 = [&](uint32_t cp, uint32_t c) {
            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(cp * S4 + K + c * 4) : "memory");
        };
        if (tid == 0) atomicSub(&Hs[(in[0] - lo) * rowlen + in[0] - lo], 1u);
        if ((uint32_t)tid < head) pair1(in[tid ? tid - 1 : 0], in[tid]);
        auto word16 = [&](uint4 q, uint32_t cp) {
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint32_t c = (w4[a] >> (8 * b)) & 0xff;
                    pair1(cp, c);
                    cp = c;
                }
        };
        uint32_t i = tid;
        for (; i + 7 * HIST_THREADS < nv; i += 8 * HIST_THREADS) {
            uint4 q[8];
            uint32_t pb[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t k = i + u * HIST_THREADS;
                q[u] = ldg_u128(v + k);
                pb[u] = (k || head) ? ldg_u8(p + 16 * (size_t)k - 1) : (q[u].x & 0xff);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) word16(q[u], pb[u]);
        }
        for (; i < nv; i += HIST_THREADS) {
            const uint4 q = ldg_u128(v + i);
            word16(q, (i || head) ? ldg_u8(p + 16 * (size_t)i - 1) : (q.x & 0xff));
        }
        for (uint32_t t = (nv << 4) + tid; t < rest; t += HIST_THREADS) {
            const uint32_t pos = head + t;
            pair1(in[pos ? pos - 1 : 0], in[pos]);
        }
        __syncthreads();
        const uint32_t r_first = rank[in[0]];
        for (uint32_t j = tid; j < hw; j += HIST_THREADS) {
            const uint32_t ri = j / nsym, rj = j - ri * nsym;
            const uint32_t si = sym_of[ri], sj = sym_of[rj];
            uint32_t c = (si >= lo && sj >= lo) ? Hs[(si - lo) * rowlen + sj - lo] : 0;
            if (ri == 0 && rj == r_first) c++;
            gH[j] = c;
        }
        return;
    }
    const bool in_smem = hw <= HIST_SMEM_PAIRS;
    uint32_t *H = in_smem ? Hs : gH;
    const uint32_t Hs_s = (uint32_t)__cvta_generic_to_shared(Hs);
    for (uint32_t j = tid; j < hw; j += HIST_THREADS) H[j] = 0;
    __syncthreads();
    // ---- pairs: H[rank(prev)][rank(cur)], the first byte follows symbol 0 (utils.h:279-357)
    if ((uint32_t)tid < head) {
        uint32_t prev = tid ? in[tid - 1] : 0;
        atomicAdd(&H[rank[prev] * nsym + rank[in[tid]]], 1u);
    }
    for (uint32_t i = tid; i < nv; i += HIST_THREADS) {
        uint4 q = ldg_u128(v + i);
        uint32_t pb = (i || head) ? p[16 * (size_t)i - 1] : 0;
        uint32_t w4[4] = {q.x, q.y, q.z, q.w};
        uint32_t rp = rank[pb], last = 0, cnt = 0;          // cnt == 0: adding it is harmless
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                uint32_t rc = rank[(w4[a] >> (8 * b)) & 0xff];
                uint32_t idx = rp * nsym + rc;
                if (in_smem) {      // branch-free: predicated reduction when the pair changes
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, %1;\n\t@p red.shared.add.u32 [%2], %3;\n\t}"
                                 ::"r"(idx), "r"(last), "r"(Hs_s + last * 4), "r"(cnt) : "memory");
                } else if (idx != last && cnt) atomicAdd(&H[last], cnt);
                cnt = (idx == last) ? cnt + 1 : 1;
                last = idx;
                rp = rc;
            }
        atomicAdd(&H[last], cnt);
    }
    for (uint32_t t = (nv << 4) + tid; t < rest; t += HIST_THREADS) {
        uint32_t pos = head + t;
        uint32_t prev = pos ? in[pos - 1] : 0;
        atomicAdd(&H[rank[prev] * nsym + rank[in[pos]]], 1u);
    }
    if (in_smem) {
        __syncthreads();
        for (uint32_t j = tid; j < hw; j += HIST_THREADS) gH[j] = Hs[j];
    }
}

// PACK / RLE streams: one CTA per stream does everything in front of the state chains (prep.cuh)
__global__ void __launch_bounds__(PREP_THREADS, 4)
prep_kernel(EncJob *jobs, uint32_t njobs) {
    __shared__ PrepSmem S;
    const uint32_t j = blockIdx.x;
    if (j >= njobs || !jobs[j].prep) return;
    prep_stream(jobs[j], S);
}

// the shared reciprocal table of the order-1 encoder, filled once per device
static cudaError_t ensure_rcp_table(cudaStream_t st) {
    static std::atomic<bool> done[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) dev = 63;
    if (dev == 63 || !done[dev].load(std::memory_order_acquire)) {
        rcp_table_kernel<<<17, 256, 0, st>>>();      // idempotent; later work on `st` is ordered after it
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(st);               // other streams of this device may encode next
        if (e != cudaSuccess) return e;
        done[dev].store(true, std::memory_order_release);
    }
    return cudaSuccess;
}

#ifdef B200_PREP_PROF
extern "C" __attribute__((visibility("default"))) int b200rans_prep_prof(unsigned long long *out16) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out16, g_prep_prof, sizeof(g_prep_prof));
}
#endif
size_t prep_area_bytes(uint32_t in_size) { return prep_plan(in_size).total; }

cudaError_t launch_prep(EncJob *d_jobs, uint32_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    // the CTA stores reciprocals beside the per-position encoder symbols (cta_o1_model): the table must be there
    { cudaError_t e = ensure_rcp_table(st); if (e != cudaSuccess) return e; }
    // tuning knob: unused dynamic shared memory per CTA, i.e. fewer streams in flight per SM (their tables share L2)
    static const char *e = getenv("B200RANS_PREP_PAD_KB");
    const size_t pad = e ? (size_t)atoi(e) << 10 : 0;
    if (pad) cudaFuncSetAttribute(prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
    prep_kernel<<<n, PREP_THREADS, pad, st>>>(d_jobs, n);
    return cudaGetLastError();
}

cudaError_t launch_hist(EncJob *d_jobs, uint32_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    hist_kernel<<<n, HIST_THREADS, 0, st>>>(d_jobs, n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------ launchers
static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

cudaError_t launch_enc(EncJob *d_jobs, uint32_t n, uint32_t route, Pool pool, cudaStream_t st, bool inslot) {
    if (!n) return cudaSuccess;
    const bool o1 = route != ROUTE_O0;
    if (o1) { cudaError_t e = ensure_rcp_table(st); if (e != cudaSuccess) return e; }
    uint32_t ws = route == ROUTE_O1_WIDE ? ENC_SMEM_O1_WIDE : route == ROUTE_O1_PREP ? ENC_SMEM_O1_PREP
                : o1 ? ENC_SMEM_O1 : ENC_SMEM_O0;
    if (o1 && route != ROUTE_O1_PREP) {    // tuning knob: shared memory per order-1 stream
        static const char *e = getenv("B200RANS_ENC_O1_SMEM");
        static const char *ew = getenv("B200RANS_ENC_O1_WIDE_SMEM");
        const char *k = route == ROUTE_O1_WIDE ? ew : e;
        if (k && atoi(k) >= 8192 && atoi(k) <= 100000) ws = (uint32_t)atoi(k) & ~15u;
    }
    size_t sm = (size_t)ws * (o1 ? ENC_WARPS_O1 : ENC_WARPS);
    if (route == ROUTE_O1_PREP) {
        cudaFuncSetAttribute(enc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        enc_kernel<true, true><<<cdiv(n, ENC_WARPS_O1), ENC_WARPS_O1 * 32, sm, st>>>(d_jobs, n, ws, pool, route, inslot ? 1u : 0u);
    } else if (o1) {
        cudaFuncSetAttribute(enc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        enc_kernel<true, false><<<cdiv(n, ENC_WARPS_O1), ENC_WARPS_O1 * 32, sm, st>>>(d_jobs, n, ws, pool, route, inslot ? 1u : 0u);
    } else {
        cudaFuncSetAttribute(enc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        enc_kernel<false, false><<<cdiv(n, ENC_WARPS), ENC_WARPS * 32, sm, st>>>(d_jobs, n, ws, pool, route, inslot ? 1u : 0u);
    }
    return cudaGetLastError();
}

cudaError_t launch_dec(DecJob *d_jobs, uint32_t n, bool o1, Pool pool, cudaStream_t st) {
    if (!n) return cudaSuccess;
    uint32_t ws = o1 ? DEC_SMEM_O1 : DEC_SMEM_O0;
    if (o1) {                              // tuning knob: shared memory per order-1 stream
        static const char *e = getenv("B200RANS_DEC_O1_SMEM");
        if (e && atoi(e) >= 8192 && atoi(e) <= 100000) ws = (uint32_t)atoi(e) & ~15u;
    }
    size_t sm = (size_t)ws * (o1 ? DEC_WARPS_O1 : DEC_WARPS);
    if (o1) {
        cudaFuncSetAttribute(dec_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        dec_kernel<true><<<cdiv(n, DEC_WARPS_O1), DEC_WARPS_O1 * 32, sm, st>>>(d_jobs, n, ws, pool);
    } else {
        cudaFuncSetAttribute(dec_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        dec_kernel<false><<<cdiv(n, DEC_WARPS), DEC_WARPS * 32, sm, st>>>(d_jobs, n, ws, pool);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------
// Method trial (fqzcomp5.c:1979-2119, the rANS members of compress_with_methods):
// the caller's items come in groups of candidate encodings of the same input, items
// first[k] .. first[k+1]-1 belonging to input k.  trial_sizes records every candidate's size,
// trial_pick keeps the first smallest of each group (fqzcomp5.c:2097 `best_sz >
// out_len`, methods tried in list order) and renumbers it to item = input so that
// the packing kernels deliver exactly one stream per input.
// ------------------------------------------------------------------------
__global__ void trial_sizes_kernel(const EncJob *jobs, uint32_t njobs, uint32_t *csize, uint32_t *jobidx) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    const EncJob &J = jobs[j];
    if (J.item == 0xffffffffu) return;
    csize[J.item] = J.status == ST_OK ? J.head_len + J.tail_len : 0u;
    jobidx[J.item] = j;
}

__global__ void trial_pick_kernel(EncJob *jobs, uint32_t ninputs, const uint32_t *first, const uint32_t *csize,
                                  const uint32_t *jobidx, int32_t *best) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ninputs) return;
    const uint32_t f0 = first[k], f1 = first[k + 1];
    uint32_t best_sz = 0xffffffffu;
    int32_t b = -1;
    for (uint32_t c = f0; c < f1; c++) {
        uint32_t sz = csize[c];
        if (sz && best_sz > sz) { best_sz = sz; b = (int32_t)(c - f0); }
    }
    best[k] = b;
    for (uint32_t c = f0; c < f1; c++) {
        EncJob &J = jobs[jobidx[c]];
        if ((int32_t)(c - f0) == b) J.item = k;
        else J.item = 0xffffffffu;
    }
    if (b < 0 && f1 > f0) {  // every candidate failed: report a failed stream for this input
        EncJob &J = jobs[jobidx[f0]];
        J.item = k; J.status = ST_FAIL;
    }
}

cudaError_t launch_trial_select(EncJob *d_jobs, uint32_t njobs, uint32_t ninputs, const uint32_t *d_first,
                                uint32_t *d_csize, uint32_t *d_jobidx, int32_t *d_best, cudaStream_t st) {
    if (!njobs || !ninputs) return cudaSuccess;
    trial_sizes_kernel<<<cdiv(njobs, 256), 256, 0, st>>>(d_jobs, njobs, d_csize, d_jobidx);
    trial_pick_kernel<<<cdiv(ninputs, 128), 128, 0, st>>>(d_jobs, ninputs, d_first, d_csize, d_jobidx, d_best);
    return cudaGetLastError();
}

cudaError_t launch_pack(const EncJob *d_jobs, uint32_t n, uint64_t *d_off, uint32_t *d_size,
                        uint64_t *d_total, uint8_t *d_out, uint64_t out_cap, uint32_t align, cudaStream_t st) {
    if (!n) return cudaSuccess;
    scan_kernel<<<1, 1024, 0, st>>>(d_jobs, n, d_off, d_size, d_total, align);
    gather_kernel<<<n, 256, 0, st>>>(d_jobs, n, d_off, d_size, d_out, out_cap);
    return cudaGetLastError();
}

}  // namespace b200
