// fastq.cu -- FASTQ block split and join on the device: the step either side of the codec.
//
// Split restates load_seqs (fqzcomp5.c:279-410) as data-parallel passes over a block of
// 4-line FASTQ text already in HBM:
//   1. count_kernel      newlines per 16 KiB tile                        (reads the text)
//   2. scan_small        exclusive scan of the tile counts
//   3. mark_kernel       position of every newline, in order            (reads the n/8 bytes of match
//                        masks pass 1 left behind)
//   4. records_kernel    one thread per record: field lengths, '@' / '+' / length checks,
//                        the trailing partial record's rules, fixed_len
//   5. scan (3 kernels)  offsets of every record in the name and seq/qual buffers
//   6. scatter_kernel    one CTA per 32 records, their text staged in shared memory: name + NUL,
//                        seq, qual - 33, READ2 flag, '@' / '+' checks
//                                                                       (reads the text, writes the buffers)
//   7. finalize_kernel   the block's totals
// Join is output_fastq (fqzcomp5.c:3440-3480) the other way round, with the +33 of
// fqzcomp5.c:2532-2533 folded into the copy.
//
// Algorithmic bytes: split reads n and writes ~n (names, bases, qualities, 8 bytes per
// record); join the same.  The text is read twice (passes 1 and 6).
#include "fastq.h"
#include "common.cuh"

namespace b200 {
namespace {

constexpr uint32_t TILE = 16384;         // bytes per CTA in the byte-search passes
constexpr uint32_t TPB = 512;            // 32 bytes per thread
constexpr uint32_t GREC = 32;            // records per CTA in the copy kernels
constexpr uint32_t SPAN_CAP = 24576;     // bytes of a group's text staged in shared memory
constexpr uint32_t STILE = 2048;         // elements per CTA in the offset scan (8 per thread)
constexpr uint32_t FREAD2 = 128;         // FQZ_FREAD2, htscodecs/fqzcomp_qual.h:45

struct FqWork {                          // zeroed before every call
    uint32_t total;                      // newlines (split) / NULs (join) in the block
    uint32_t err;                        // bit 0 malformed, bit 1 capacity
    uint32_t drop_last;                  // the last complete record is held back (fqzcomp5.c:382-384)
    uint32_t not_minlen;                 // ~min(len) over the records that reach fixed_len's update
    uint32_t maxlen;
    uint32_t have_len;
    uint32_t tot[2];                     // totals of the two scanned arrays
    uint32_t cut;                        // kseq mode: records taken by the block-size rule
    uint32_t pad[7];
};

__device__ __forceinline__ uint32_t eq_mask4(uint32_t w, uint32_t pat) {    // 4 bits: byte j of w == pattern
    uint32_t x = __vcmpeq4(w, pat) & 0x01010101u;
    return (x | (x >> 7) | (x >> 14) | (x >> 21)) & 0xfu;
}

// bitmask (bit j <=> text[base + j] == B) of the 32 bytes this thread owns; bytes past n never match
template <bool CHECK_NUL>
__device__ __forceinline__ uint32_t match32(const uint8_t *text, uint32_t n, uint32_t base, uint32_t B, bool *nul) {
    uint32_t m = 0;
    if (base >= n) return 0;
    const uint32_t pat = B * 0x01010101u;
    if (base + 32 <= n) {
        const uint4 *p = (const uint4 *)(text + base);
        uint4 a = __ldg(p), b = __ldg(p + 1);
        uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            m |= eq_mask4(w[k], pat) << (4 * k);
            if (CHECK_NUL && __vcmpeq4(w[k], 0u)) *nul = true;
        }
    } else {
        for (uint32_t j = 0; base + j < n; j++) {
            uint32_t c = text[base + j];
            if (c == B) m |= 1u << j;
            if (CHECK_NUL && c == 0) *nul = true;
        }
    }
    return m;
}

template <bool CHECK_NUL>
__global__ void __launch_bounds__(TPB)
count_kernel(const uint8_t *__restrict__ text, uint32_t n, uint32_t B, uint32_t *__restrict__ tile_count,
             uint32_t *__restrict__ masks, FqWork *W) {
    __shared__ uint32_t ws[TPB / 32];
    const uint32_t base = blockIdx.x * TILE + threadIdx.x * 32;
    bool nul = false;
    const uint32_t m = match32<CHECK_NUL>(text, n, base, B, &nul);
    masks[blockIdx.x * TPB + threadIdx.x] = m;       // mark_kernel reads n/8 bytes of masks instead of the text again
    uint32_t c = __popc(m);
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    if (CHECK_NUL && __any_sync(FULL, nul) && (threadIdx.x & 31) == 0) atomicOr(&W->err, 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (uint32_t k = 0; k < TPB / 32; k++) t += ws[k];
        tile_count[blockIdx.x] = t;
    }
}

// Exclusive scan of a short array by one CTA.  The element count is n_static, or
// ceil(*n_dev / div) when n_dev is given.  Up to two arrays (in[a] -> out[a], total[a]).
__global__ void __launch_bounds__(1024)
scan_small(const uint32_t *in0, uint32_t *out0, uint32_t *tot0, const uint32_t *in1, uint32_t *out1,
           uint32_t *tot1, uint32_t n_static, const uint32_t *n_dev, uint32_t n_cap, uint32_t div) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    uint32_t n = n_static;
    if (n_dev) { uint32_t v = *n_dev; if (v > n_cap) v = n_cap; n = (v + div - 1) / div; }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int a = 0; a < 2; a++) {
        const uint32_t *in = a ? in1 : in0;
        uint32_t *out = a ? out1 : out0, *tot = a ? tot1 : tot0;
        if (!in) continue;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (uint32_t base = 0; base < n; base += 1024) {
            uint32_t k = base + threadIdx.x;
            uint32_t v = k < n ? in[k] : 0;
            uint32_t x = warp_incl_scan(v, lane);
            if (lane == 31) wsum[wid] = x;
            __syncthreads();
            if (wid == 0) wsum[lane] = warp_incl_scan(wsum[lane], lane);
            __syncthreads();
            uint32_t excl = carry + (wid ? wsum[wid - 1] : 0) + x - v;
            if (k < n) out[k] = excl;
            __syncthreads();
            if (threadIdx.x == 1023) carry = excl + v;
            __syncthreads();
        }
        if (threadIdx.x == 0 && tot) *tot = carry;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(TPB)
mark_kernel(const uint32_t *__restrict__ masks, const uint32_t *__restrict__ tile_off,
            uint32_t *__restrict__ pos, uint32_t cap) {
    __shared__ uint32_t ws[TPB / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * TILE + threadIdx.x * 32;
    uint32_t m = masks[blockIdx.x * TPB + threadIdx.x];
    uint32_t c = __popc(m);
    uint32_t incl = warp_incl_scan(c, lane);
    if (lane == 31) ws[wid] = incl;
    __syncthreads();
    uint32_t o = tile_off[blockIdx.x] + incl - c;
    for (int k = 0; k < wid; k++) o += ws[k];
    while (m) {
        uint32_t j = __ffs(m) - 1;
        m &= m - 1;
        if (o < cap) pos[o] = base + j;
        o++;
    }
}

// ---------------------------------------------------------------- split: per-record pass
// nl[] = newline positions.  Record r owns newlines 4r .. 4r+3.
__global__ void __launch_bounds__(256)
records_kernel(const uint8_t *__restrict__ text, uint32_t n, const uint32_t *__restrict__ nl, uint32_t cap_nl,
               uint32_t max_records, uint32_t *__restrict__ nlen1, uint32_t *__restrict__ len, FqWork *W) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t total = W->total;
    bool capped = false;
    if (total > cap_nl) { capped = true; total = cap_nl; }
    uint32_t R = total >> 2;
    if (R > max_records) { capped = true; R = max_records; }
    if (capped && r == 0) atomicOr(&W->err, 2u);
    // length statistics for fixed_len (fqzcomp5.c:344-348), reduced per warp before the atomics
    uint32_t not_mn = 0, mx = 0, have = 0, e = 0;
    auto note_len = [&](uint32_t l) { not_mn = max(not_mn, ~l); mx = max(mx, l); have = 1; };
    if (r < R) {
        const uint32_t a = r ? nl[4 * r - 1] + 1 : 0;
        const uint32_t n0 = nl[4 * r], n1 = nl[4 * r + 1], n2 = nl[4 * r + 2], n3 = nl[4 * r + 3];
        const uint32_t sl = n1 - n0 - 1, ql = n3 - n2 - 1;
        nlen1[r] = n0 - a;                       // name without '@', plus its NUL
        len[r] = sl;
        note_len(sl);
        if (sl != ql) {                          // :382-386
            if (r == R - 1 && n3 == n - 1) {
                // held back, not an error; scatter_kernel will not see it, so its '@' and '+'
                // (:302-304, :351-352) are judged here
                W->drop_last = 1;
                if (text[a] != '@' || text[n1 + 1] != '+') e = 1;
            } else e = 1;
        }
    } else if (r == R && !capped) {
        // the record the block ends in (no fourth newline): the checks load_seqs makes
        // before it notices the end of the block.  (A held-back last record ends on the
        // block's last byte, so nothing follows it and start == n.)
        const uint32_t start = R ? nl[4 * R - 1] + 1 : 0;
        const uint32_t k = total - 4 * R;        // complete lines of the partial record
        if (start < n) {
            if (text[start] != '@') e = 1;       // :302-304
            // the sequence line counts once its newline is inside the block and not its last
            // byte (:339-340 breaks first otherwise): length noted (:344-348), '+' checked (:351)
            if (k >= 2 && nl[4 * R + 1] != n - 1) {
                note_len(nl[4 * R + 1] - nl[4 * R] - 1);
                if (text[nl[4 * R + 1] + 1] != '+') e = 1;
            }
        }
    }
    not_mn = __reduce_max_sync(FULL, not_mn);
    mx = __reduce_max_sync(FULL, mx);
    have = __reduce_max_sync(FULL, have);
    e = __reduce_max_sync(FULL, e);
    if ((threadIdx.x & 31) == 0) {
        // the statistics only grow: look before touching them (same-address atomics serialise in L2,
        // and with fixed-length reads every warp would send the same two values)
        volatile FqWork *V = W;
        if (have) {
            if (not_mn > V->not_minlen) atomicMax(&W->not_minlen, not_mn);
            if (mx > V->maxlen) atomicMax(&W->maxlen, mx);
            if (!V->have_len) W->have_len = 1;
        }
        if (e) atomicOr(&W->err, 1u);
    }
}

// ---------------------------------------------------------------- split, kseq mode (load_seqs_kseq, fqzcomp5.c:423-623)
// The live loader reads kseq records (kseq.h:176-218).  For strict 4-line FASTQ a record is
//   name    = header up to its first isspace() byte,   comment = the rest of the header line
//   stored  = name [+ ' ' + comment when the comment is not empty] + NUL          (:485-510)
//   size    = name.l + 1 + seq.l + qual.l, and a block takes records while the running total
//             stays <= blk_size, one at least (:468-476)
// hdr_info(): stored length and the position of the separator inside the header (0xffffffff: none).
__device__ __forceinline__ bool is_space(uint32_t c) { return c == ' ' || (c >= 9 && c <= 13); }

__global__ void __launch_bounds__(256)
records_kseq_kernel(const uint8_t *__restrict__ text, uint32_t n, const uint32_t *__restrict__ nl, uint32_t cap_nl,
                    uint32_t max_records, uint32_t *__restrict__ nlen1, uint32_t *__restrict__ len,
                    uint32_t *__restrict__ rsize, uint32_t *__restrict__ wsp, uint8_t *__restrict__ rerr, FqWork *W) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t total = W->total;
    bool capped = false;
    if (total > cap_nl) { capped = true; total = cap_nl; }
    uint32_t R = total >> 2;
    if (R > max_records) { capped = true; R = max_records; }
    if (capped && r == 0) atomicOr(&W->err, 2u);
    if (r >= R) return;
    const uint32_t a = r ? nl[4 * r - 1] + 1 : 0;
    const uint32_t n0 = nl[4 * r], n1 = nl[4 * r + 1], n2 = nl[4 * r + 2], n3 = nl[4 * r + 3];
    const uint32_t sl = n1 - n0 - 1, ql = n3 - n2 - 1;
    const uint32_t L = n0 - a - 1;                   // header without '@'
    uint32_t ws = 0xffffffffu;
    for (uint32_t i = 0; i < L; i++) if (is_space(text[a + 1 + i])) { ws = i; break; }
    const uint32_t name_l = ws == 0xffffffffu ? L : ws;
    const uint32_t stored = (ws != 0xffffffffu && ws + 1 == L) ? L - 1 : L;      // an empty comment leaves no separator
    nlen1[r] = stored + 1;
    len[r] = sl;
    rsize[r] = name_l + 1 + sl + ql;
    wsp[r] = ws;
    uint32_t e = 0;
    if (text[a] != '@' || text[n1 + 1] != '+' || sl != ql) e = 1;                // kseq: FASTA / "-2 truncated quality"
    if (sl) {
        const uint32_t c0 = text[n0 + 1];
        if (c0 == '@' || c0 == '+' || c0 == '>') e = 1;                          // would end kseq's sequence loop
        if (text[n1 - 1] == '\r' || text[n3 - 1] == '\r') e = 1;                // CRLF text is not handled here
    }
    rerr[r] = (uint8_t)e;
}

// records taken: the longest prefix whose sizes sum to <= blk_size, one at least.  roff = exclusive
// prefix of rsize.
__global__ void __launch_bounds__(256)
cut_kseq_kernel(const uint32_t *__restrict__ rsize, const uint32_t *__restrict__ roff, uint32_t cap_nl,
                uint32_t max_records, uint32_t blk_size, FqWork *W) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t R = min(min(W->total, cap_nl) >> 2, max_records);
    if (r >= R) return;
    const uint64_t incl = (uint64_t)roff[r] + rsize[r];
    const bool fits = r == 0 || incl <= blk_size;
    const bool next_fits = r + 1 < R && (uint64_t)roff[r + 1] + rsize[r + 1] <= blk_size;
    if (fits && !next_fits) W->cut = r + 1;          // the sums grow with r: exactly one thread
}

// statistics and checks over the records taken (and the one behind them, which the reference has
// parsed before it decides to keep it for the next block)
__global__ void __launch_bounds__(256)
stats_kseq_kernel(const uint32_t *__restrict__ len, const uint8_t *__restrict__ rerr, uint32_t cap_nl,
                  uint32_t max_records, FqWork *W) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t R = min(min(W->total, cap_nl) >> 2, max_records), cut = W->cut;
    uint32_t not_mn = 0, mx = 0, have = 0, e = 0;
    if (r < cut) { not_mn = ~len[r]; mx = len[r]; have = 1; }
    if (r <= cut && r < R) e = rerr[r];
    not_mn = __reduce_max_sync(FULL, not_mn);
    mx = __reduce_max_sync(FULL, mx);
    have = __reduce_max_sync(FULL, have);
    e = __reduce_max_sync(FULL, e);
    if ((threadIdx.x & 31) == 0) {
        volatile FqWork *V = W;
        if (have) {
            if (not_mn > V->not_minlen) atomicMax(&W->not_minlen, not_mn);
            if (mx > V->maxlen) atomicMax(&W->maxlen, mx);
            if (!V->have_len) W->have_len = 1;
        }
        if (e) atomicOr(&W->err, 1u);
    }
}

__global__ void set_count(FqWork *W, uint32_t cap_nl, uint32_t max_records) {
    uint32_t t = min(W->total, cap_nl) >> 2;
    W->tot[0] = min(t, max_records);
}

// ---------------------------------------------------------------- offsets: tiled exclusive scan
__global__ void __launch_bounds__(256)
scan_reduce(const uint32_t *__restrict__ v0, const uint32_t *__restrict__ v1, const uint32_t *n_dev, uint32_t n_cap,
            uint32_t *__restrict__ sums0, uint32_t *__restrict__ sums1) {
    __shared__ uint32_t ws[2][8];
    uint32_t n = *n_dev; if (n > n_cap) n = n_cap;
    const uint32_t base = blockIdx.x * STILE;
    if (base >= n) return;
    uint32_t s0 = 0, s1 = 0;
    for (uint32_t k = threadIdx.x; k < STILE; k += 256) {
        uint32_t i = base + k;
        if (i < n) { s0 += v0[i]; if (v1) s1 += v1[i]; }
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    if ((threadIdx.x & 31) == 0) { ws[0][threadIdx.x >> 5] = s0; ws[1][threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t0 = 0, t1 = 0;
        for (int k = 0; k < 8; k++) { t0 += ws[0][k]; t1 += ws[1][k]; }
        sums0[blockIdx.x] = t0;
        if (v1) sums1[blockIdx.x] = t1;
    }
}

__global__ void __launch_bounds__(256)
scan_apply(const uint32_t *__restrict__ v0, const uint32_t *__restrict__ v1, const uint32_t *n_dev, uint32_t n_cap,
           const uint32_t *__restrict__ toff0, const uint32_t *__restrict__ toff1, uint32_t *__restrict__ o0,
           uint32_t *__restrict__ o1) {
    __shared__ uint32_t ws[2][8];
    uint32_t n = *n_dev; if (n > n_cap) n = n_cap;
    const uint32_t base = blockIdx.x * STILE;
    if (base >= n) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // thread t owns elements base + 8t .. base + 8t + 7
    uint32_t a[8], b[8], sa = 0, sb = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t i = base + threadIdx.x * 8 + k;
        a[k] = i < n ? v0[i] : 0; b[k] = (v1 && i < n) ? v1[i] : 0;
        sa += a[k]; sb += b[k];
    }
    uint32_t ia = warp_incl_scan(sa, lane), ib = warp_incl_scan(sb, lane);
    if (lane == 31) { ws[0][wid] = ia; ws[1][wid] = ib; }
    __syncthreads();
    uint32_t ea = toff0[blockIdx.x] + ia - sa, eb = (v1 ? toff1[blockIdx.x] : 0) + ib - sb;
    for (int k = 0; k < wid; k++) { ea += ws[0][k]; eb += ws[1][k]; }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t i = base + threadIdx.x * 8 + k;
        if (i < n) { o0[i] = ea; if (v1) o1[i] = eb; }
        ea += a[k]; eb += b[k];
    }
}

// ---------------------------------------------------------------- copies
// warp-cooperative copy with a byte-wise add (0, -33 or +33 mod 256), any alignment:
// 4-byte stores assembled from aligned 4-byte loads.  Reads stay inside the aligned
// words that hold src[0 .. n).  src may be global or shared memory.
__device__ __forceinline__ void warp_copy_add(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, uint32_t n,
                                              uint32_t add4, int lane) {
    if (!n) return;
    const uint32_t add = add4 & 0xff;
    uint32_t head = (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3);
    if (head > n) head = n;
    if ((uint32_t)lane < head) dst[lane] = (uint8_t)(src[lane] + add);
    dst += head; src += head; n -= head;
    const uint32_t nw = n >> 2;
    const uint32_t sh = ((uintptr_t)src & 3) * 8;
    const uint32_t *sw = (const uint32_t *)((uintptr_t)src & ~(uintptr_t)3);
    uint32_t *dw = (uint32_t *)dst;
    if (sh == 0) {
#pragma unroll 4
        for (uint32_t i = lane; i < nw; i += 32) dw[i] = __vadd4(sw[i], add4);
    } else {
#pragma unroll 4
        for (uint32_t i = lane; i < nw; i += 32)
            dw[i] = __vadd4(__funnelshift_r(sw[i], sw[i + 1], sh), add4);
    }
    for (uint32_t i = (nw << 2) + lane; i < n; i += 32) dst[i] = (uint8_t)(src[i] + add);
}

// The staged fast path: source bytes in shared memory (32-bit shared addresses), one byte per
// lane and round.  Far fewer instructions than the word path for the 25-150 byte fields of short
// reads; the 32 bytes a warp stores per round are contiguous and merge in L2.
__device__ __forceinline__ uint32_t lds_b(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void smem_copy_add(uint8_t *__restrict__ dst, uint32_t saddr, uint32_t n, uint32_t add,
                                              int lane) {
    uint8_t *d = dst + lane;
    saddr += lane;
    for (uint32_t i = lane; i < n; i += 32, saddr += 32, d += 32) *d = (uint8_t)(lds_b(saddr) + add);
}

// cooperative copy of text[lo, hi) into shared memory at its position relative to `base`
// (base <= lo, 16-byte aligned); 16-byte loads where the text allows
__device__ __forceinline__ void stage_span(uint8_t *s, const uint8_t *__restrict__ g, uint32_t base, uint32_t hi,
                                           uint32_t n) {
    const uint32_t nv = (hi - base + 15) >> 4;
    for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) {
        const uint32_t p = base + 16 * i;
        if (p + 16 <= n) ((uint4 *)s)[i] = __ldg((const uint4 *)(g + p));
        else for (uint32_t q = p; q < n; q++) s[q - base] = g[q];
    }
}

// One CTA per GREC consecutive records: their text is one contiguous span, staged in shared
// memory with coalesced 16-byte loads when it fits (reads of 150-byte fields straight from
// HBM are latency bound); each warp then writes whole records.  Groups whose span exceeds
// SPAN_CAP (long reads) copy from global memory.
__global__ void __launch_bounds__(256)
scatter_kernel(const uint8_t *__restrict__ text, uint32_t n, const uint32_t *__restrict__ nl, uint32_t cap_nl,
               uint32_t max_records, const uint32_t *__restrict__ name_off, const uint32_t *__restrict__ seq_off,
               uint8_t *__restrict__ name, uint8_t *__restrict__ seq, uint8_t *__restrict__ qual, uint32_t name_cap,
               uint32_t seq_cap, uint32_t *__restrict__ flag, FqWork *W, const uint32_t *__restrict__ wsp) {
    // wsp != null: kseq mode -- the records taken are W->cut, names are stored in kseq's form (separator
    // -> ' ', dropped before an empty comment) and READ2 follows fqzcomp5.c:512-520
    __shared__ uint32_t s_nl[4 * GREC + 1];      // newline in front of the group, then the group's own
    __shared__ uint32_t s_no[GREC], s_so[GREC];
    __shared__ __align__(16) uint8_t s_text[SPAN_CAP + 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t total = min(W->total, cap_nl);
    const uint32_t R = wsp ? W->cut : min(total >> 2, max_records) - W->drop_last;
    const uint32_t r0 = blockIdx.x * GREC;
    if (r0 >= R) return;
    const uint32_t g = min(GREC, R - r0);
    for (uint32_t i = threadIdx.x; i < 4 * g + 1; i += blockDim.x)
        s_nl[i] = (r0 == 0 && i == 0) ? 0xffffffffu : nl[4 * r0 - 1 + i];
    for (uint32_t i = threadIdx.x; i < g; i += blockDim.x) { s_no[i] = name_off[r0 + i]; s_so[i] = seq_off[r0 + i]; }
    __syncthreads();
    const uint32_t lo = s_nl[0] + 1, hi = s_nl[4 * g] + 1, base = lo & ~15u;
    const bool staged = hi - base <= SPAN_CAP;
    if (staged) stage_span(s_text, text, base, hi, n);
    __syncthreads();
    // position p of the text: shared address sa0 + p when staged
    const uint32_t sa0 = (uint32_t)__cvta_generic_to_shared(s_text) - base;
    auto byte_at = [&](uint32_t p) -> uint32_t { return staged ? lds_b(sa0 + p) : (uint32_t)text[p]; };
    auto copy = [&](uint8_t *dst, uint32_t p, uint32_t len, uint32_t add4) {
        if (staged) smem_copy_add(dst, sa0 + p, len, add4 & 0xff, lane);
        else warp_copy_add(dst, text + p, len, add4, lane);
    };
    uint32_t err = 0;
    for (uint32_t j = wid; j < g; j += blockDim.x >> 5) {
        const uint32_t r = r0 + j;
        const uint32_t a = s_nl[4 * j] + 1, n0 = s_nl[4 * j + 1], n1 = s_nl[4 * j + 2], n2 = s_nl[4 * j + 3];
        const uint32_t L = n0 - a - 1, sl = n1 - n0 - 1;
        const uint32_t no = s_no[j], so = s_so[j];
        if (byte_at(a) != '@' || byte_at(n1 + 1) != '+') err |= 1u;           // fqzcomp5.c:302-304, :351-352
        if ((uint64_t)no + L + 1 > name_cap || (uint64_t)so + sl > seq_cap) { err |= 2u; continue; }
        if (wsp) {
            // kseq's stored name: header with its first isspace() byte turned into ' ', or cut there
            // when nothing follows it (an empty comment, fqzcomp5.c:487-507)
            const uint32_t ws = wsp[r];
            const uint32_t Ls = (ws != 0xffffffffu && ws + 1 == L) ? L - 1 : L;
            copy(name + no, a + 1, Ls, 0);
            __syncwarp();
            if (lane == 0) { if (ws < Ls) name[no + ws] = ' '; name[no + Ls] = 0; }
            copy(seq + so, n0 + 1, sl, 0);
            copy(qual + so, n2 + 1, sl, 0xdfdfdfdfu);                         // - 33 (fqzcomp5.c:563-564)
            // READ2 (:512-520): the stored string ends in "/2" (name longer than one byte), or equals the
            // previous record's stored string
            const uint32_t name_l = ws == 0xffffffffu ? L : ws;
            bool f = name_l > 1 && Ls >= 2 && byte_at(a + Ls) == '2' && byte_at(a + Ls - 1) == '/';
            if (!f && r) {
                uint32_t pa, pn0;
                if (j) { pa = s_nl[4 * j - 4] + 1; pn0 = s_nl[4 * j - 3]; }
                else { pa = r > 1 ? nl[4 * r - 5] + 1 : 0; pn0 = nl[4 * r - 4]; }
                const uint32_t pL = pn0 - pa - 1, pws = wsp[r - 1];
                const uint32_t pLs = (pws != 0xffffffffu && pws + 1 == pL) ? pL - 1 : pL;
                bool same = pLs == Ls && (pws < pLs ? pws : 0xffffffffu) == (ws < Ls ? ws : 0xffffffffu);
                if (same) {
                    bool eq = true;     // separators sit at the same place: compare everything else
                    for (uint32_t i = lane; i < Ls; i += 32) eq = eq && (i == ws || text[a + 1 + i] == text[pa + 1 + i]);
                    same = __all_sync(FULL, eq);
                }
                f = same;
            }
            if (lane == 0) flag[r] = f ? FREAD2 : 0u;
            continue;
        }
        copy(name + no, a + 1, L, 0);
        if (lane == 0) name[no + L] = 0;
        copy(seq + so, n0 + 1, sl, 0);
        copy(qual + so, n2 + 1, sl, 0xdfdfdfdfu);                             // - 33 (fqzcomp5.c:375)
        // READ2: name ends in "/2" (the reference tests the running buffer offset, :320-323), or
        // repeats the previous record's name (:324-326)
        bool f = L >= 2 && no + L + 1 > 3 && byte_at(n0 - 1) == '2' && byte_at(n0 - 2) == '/';
        if (!f && r) {
            uint32_t pa, pn0;
            if (j) { pa = s_nl[4 * j - 4] + 1; pn0 = s_nl[4 * j - 3]; }
            else { pa = r > 1 ? nl[4 * r - 5] + 1 : 0; pn0 = nl[4 * r - 4]; }      // previous group's last
            bool same = pn0 - pa - 1 == L;
            if (same) {
                bool eq = true;
                if (j && staged) for (uint32_t i = lane; i < L; i += 32) eq = eq && lds_b(sa0 + a + 1 + i) == lds_b(sa0 + pa + 1 + i);
                else for (uint32_t i = lane; i < L; i += 32) eq = eq && text[a + 1 + i] == text[pa + 1 + i];
                same = __all_sync(FULL, eq);
            }
            f = same;
        }
        if (lane == 0) flag[r] = f ? FREAD2 : 0u;
    }
    if (err && lane == 0) atomicOr(&W->err, err);
}

__global__ void split_finalize(const uint32_t *nl, uint32_t cap_nl, uint32_t max_records, const uint32_t *name_off,
                               const uint32_t *seq_off, const uint32_t *nlen1, const uint32_t *len, FqWork *W,
                               FqInfo *info, int kseq) {
    uint32_t total = min(W->total, cap_nl);
    uint32_t R = kseq ? W->cut : min(total >> 2, max_records) - W->drop_last;
    // kseq mode: 1 when the block-size rule ended the block (a complete record was left for the next one),
    // 0 when the text ran out first (the caller appends more text and calls again, unless the file ended)
    info->more = kseq ? (R < min(total >> 2, max_records) ? 1u : 0u) : 0u;
    info->status = (W->err & 1u) ? 1 : (W->err & 2u) ? 2 : 0;
    info->num_records = R;
    info->name_len = R ? name_off[R - 1] + nlen1[R - 1] : 0;
    info->seq_len = info->qual_len = R ? seq_off[R - 1] + len[R - 1] : 0;
    const uint32_t mn = ~W->not_minlen, mx = W->maxlen;
    info->fixed_len = !W->have_len ? -1 : (mn == mx ? (int32_t)mx : 0);
    info->consumed = R ? nl[4 * R - 1] + 1 : 0;
    info->text_len = 0;
}

// ---------------------------------------------------------------- join
// One CTA per GREC consecutive records: their names, bases and qualities are three contiguous
// spans, staged in shared memory when they fit; each warp writes whole records of text.
constexpr uint32_t JN_CAP = 4096, JS_CAP = 10240;
__global__ void __launch_bounds__(256)
gather_kernel_fq(const uint8_t *__restrict__ name, uint32_t name_len, const uint32_t *__restrict__ nul,
                 const uint8_t *__restrict__ seq, const uint8_t *__restrict__ qual, const uint32_t *__restrict__ len,
                 const uint32_t *__restrict__ seq_off, uint32_t R, int plus_name, uint8_t *__restrict__ text,
                 uint32_t text_cap, FqWork *W, FqInfo *info) {
    __shared__ uint32_t s_nul[GREC + 1], s_so[GREC], s_len[GREC];
    __shared__ __align__(16) uint8_t s_name[JN_CAP + 32], s_seq[JS_CAP + 32], s_qual[JS_CAP + 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t r0 = blockIdx.x * GREC;
    if (r0 >= R || W->total != R) return;
    const uint32_t g = min(GREC, R - r0);
    for (uint32_t i = threadIdx.x; i < g + 1; i += blockDim.x)
        s_nul[i] = (r0 == 0 && i == 0) ? 0xffffffffu : nul[r0 - 1 + i];
    for (uint32_t i = threadIdx.x; i < g; i += blockDim.x) { s_so[i] = seq_off[r0 + i]; s_len[i] = len[r0 + i]; }
    __syncthreads();
    const uint32_t nlo = s_nul[0] + 1, nhi = s_nul[g] + 1, nbase = nlo & ~15u;
    const uint32_t slo = s_so[0], shi = s_so[g - 1] + s_len[g - 1], sbase = slo & ~15u;
    const bool staged = nhi - nbase <= JN_CAP && shi - sbase <= JS_CAP;
    if (staged) {
        // the spans end inside the buffers; the last 16-byte load may not, so bound it by the span
        stage_span(s_name, name, nbase, nhi, nhi);
        stage_span(s_seq, seq, sbase, shi, shi);
        stage_span(s_qual, qual, sbase, shi, shi);
    }
    __syncthreads();
    const uint32_t na0 = (uint32_t)__cvta_generic_to_shared(s_name) - nbase;
    const uint32_t sa0 = (uint32_t)__cvta_generic_to_shared(s_seq) - sbase;
    const uint32_t qa0 = (uint32_t)__cvta_generic_to_shared(s_qual) - sbase;
    auto copy = [&](uint8_t *dst, uint32_t sa, const uint8_t *gsrc, uint32_t len, uint32_t add4) {
        if (staged) smem_copy_add(dst, sa, len, add4 & 0xff, lane);
        else warp_copy_add(dst, gsrc, len, add4, lane);
    };
    for (uint32_t j = wid; j < g; j += blockDim.x >> 5) {
        const uint32_t r = r0 + j;
        const uint32_t no = s_nul[j] + 1, L = s_nul[j + 1] - no, sl = s_len[j], so = s_so[j];
        // bytes in front of record r: names (without NULs), two copies of the bases' count, 6 per record
        uint64_t o = (uint64_t)(no - r) * (plus_name ? 2 : 1) + 2ull * so + 6ull * r;
        const uint64_t sz = (uint64_t)L * (plus_name ? 2 : 1) + 2ull * sl + 6;
        if (r == R - 1 && lane == 0) info->text_len = (uint32_t)(o + sz);
        if (o + sz > text_cap) { if (lane == 0) atomicOr(&W->err, 2u); continue; }
        uint8_t *p = text + o;
        if (lane == 0) p[0] = '@';
        copy(p + 1, na0 + no, name + no, L, 0);
        p += 1 + L;
        if (lane == 0) p[0] = '\n';
        copy(p + 1, sa0 + so, seq + so, sl, 0);
        p += 1 + sl;
        if (lane == 0) { p[0] = '\n'; p[1] = '+'; }
        p += 2;
        if (plus_name) { copy(p, na0 + no, name + no, L, 0); p += L; }
        if (lane == 0) p[0] = '\n';
        copy(p + 1, qa0 + so, qual + so, sl, 0x21212121u);                    // + 33 (fqzcomp5.c:2532-2533)
        if (lane == 0) p[1 + sl] = '\n';
    }
}

__global__ void join_finalize(uint32_t R, FqWork *W, FqInfo *info) {
    uint32_t e = W->err;
    if (W->total != R) e |= 1u;                  // name buffer does not hold R names
    info->status = (e & 1u) ? 1 : (e & 2u) ? 2 : 0;
    info->num_records = R;
    if (!R) info->text_len = 0;
}

inline uint32_t cdivu(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

struct SplitLayout {
    size_t work, tile_cnt, tile_off, masks, nl, nlen1, sums0, sums1, toff0, toff1, rsize, roff, wsp, rerr, sums2, toff2, total;
    uint32_t ntiles, cap_nl, stiles;
    SplitLayout(uint32_t n, uint32_t max_records) {
        ntiles = cdivu(n ? n : 1, TILE);
        cap_nl = 4 * max_records + 4;
        stiles = cdivu(max_records ? max_records : 1, STILE);
        size_t o = 0;
        work = o; o += al256(sizeof(FqWork));
        tile_cnt = o; o += al256((size_t)ntiles * 4);
        tile_off = o; o += al256((size_t)ntiles * 4);
        masks = o; o += al256((size_t)ntiles * TPB * 4);
        nl = o; o += al256((size_t)cap_nl * 4);
        nlen1 = o; o += al256((size_t)max_records * 4 + 4);
        sums0 = o; o += al256((size_t)stiles * 4);
        sums1 = o; o += al256((size_t)stiles * 4);
        toff0 = o; o += al256((size_t)stiles * 4);
        toff1 = o; o += al256((size_t)stiles * 4);
        // kseq mode: record sizes, their prefix, separator positions, per-record verdicts
        rsize = o; o += al256((size_t)max_records * 4 + 4);
        roff = o; o += al256((size_t)max_records * 4 + 4);
        wsp = o; o += al256((size_t)max_records * 4 + 4);
        rerr = o; o += al256((size_t)max_records + 4);
        sums2 = o; o += al256((size_t)stiles * 4);
        toff2 = o; o += al256((size_t)stiles * 4);
        total = o;
    }
};

}  // namespace

size_t fq_split_scratch_bytes(uint32_t n, uint32_t max_records) { return SplitLayout(n, max_records).total; }

cudaError_t fq_split_launch(const uint8_t *d_text, uint32_t n, uint8_t *d_name, uint8_t *d_seq, uint8_t *d_qual,
                            uint32_t name_cap, uint32_t seq_cap, uint32_t *d_len, uint32_t *d_flag,
                            uint32_t *d_name_off, uint32_t *d_seq_off, uint32_t max_records, uint8_t *S,
                            FqInfo *d_info, cudaStream_t st, int *launches, int kseq, uint32_t blk_size) {
    SplitLayout L(n, max_records);
    FqWork *W = (FqWork *)(S + L.work);
    uint32_t *tile_cnt = (uint32_t *)(S + L.tile_cnt), *tile_off = (uint32_t *)(S + L.tile_off);
    uint32_t *nl = (uint32_t *)(S + L.nl), *nlen1 = (uint32_t *)(S + L.nlen1);
    uint32_t *sums0 = (uint32_t *)(S + L.sums0), *sums1 = (uint32_t *)(S + L.sums1);
    uint32_t *toff0 = (uint32_t *)(S + L.toff0), *toff1 = (uint32_t *)(S + L.toff1);
    cudaError_t e = cudaMemsetAsync(W, 0, sizeof(FqWork), st);
    if (e != cudaSuccess) return e;
    uint32_t *masks = (uint32_t *)(S + L.masks);
    count_kernel<true><<<L.ntiles, TPB, 0, st>>>(d_text, n, '\n', tile_cnt, masks, W);
    scan_small<<<1, 1024, 0, st>>>(tile_cnt, tile_off, &W->total, nullptr, nullptr, nullptr, L.ntiles, nullptr, 0, 1);
    mark_kernel<<<L.ntiles, TPB, 0, st>>>(masks, tile_off, nl, L.cap_nl);
    uint32_t *rsize = (uint32_t *)(S + L.rsize), *roff = (uint32_t *)(S + L.roff), *wsp = (uint32_t *)(S + L.wsp);
    uint8_t *rerr = S + L.rerr;
    if (kseq)
        records_kseq_kernel<<<cdivu(max_records + 1, 256), 256, 0, st>>>(d_text, n, nl, L.cap_nl, max_records, nlen1,
                                                                       d_len, rsize, wsp, rerr, W);
    else
        records_kernel<<<cdivu(max_records + 1, 256), 256, 0, st>>>(d_text, n, nl, L.cap_nl, max_records, nlen1, d_len, W);
    // element count of the offset scans (a held-back last record is scanned too, harmlessly)
    set_count<<<1, 1, 0, st>>>(W, L.cap_nl, max_records);
    if (kseq) {
        // the block-size rule: prefix of the record sizes, the cut, then statistics over the records taken
        uint32_t *sums2 = (uint32_t *)(S + L.sums2), *toff2 = (uint32_t *)(S + L.toff2);
        scan_reduce<<<L.stiles, 256, 0, st>>>(rsize, nullptr, &W->tot[0], max_records, sums2, nullptr);
        scan_small<<<1, 1024, 0, st>>>(sums2, toff2, nullptr, nullptr, nullptr, nullptr, 0, &W->tot[0], max_records, STILE);
        scan_apply<<<L.stiles, 256, 0, st>>>(rsize, nullptr, &W->tot[0], max_records, toff2, nullptr, roff, nullptr);
        cut_kseq_kernel<<<cdivu(max_records + 1, 256), 256, 0, st>>>(rsize, roff, L.cap_nl, max_records, blk_size, W);
        stats_kseq_kernel<<<cdivu(max_records + 1, 256), 256, 0, st>>>(d_len, rerr, L.cap_nl, max_records, W);
        if (launches) *launches += 5;
    }
    scan_reduce<<<L.stiles, 256, 0, st>>>(nlen1, d_len, &W->tot[0], max_records, sums0, sums1);
    scan_small<<<1, 1024, 0, st>>>(sums0, toff0, nullptr, sums1, toff1, nullptr, 0, &W->tot[0], max_records, STILE);
    scan_apply<<<L.stiles, 256, 0, st>>>(nlen1, d_len, &W->tot[0], max_records, toff0, toff1, d_name_off, d_seq_off);
    scatter_kernel<<<cdivu(max_records ? max_records : 1, GREC), 256, 0, st>>>(
        d_text, n, nl, L.cap_nl, max_records, d_name_off, d_seq_off, d_name, d_seq, d_qual, name_cap, seq_cap,
        d_flag, W, kseq ? wsp : nullptr);
    split_finalize<<<1, 1, 0, st>>>(nl, L.cap_nl, max_records, d_name_off, d_seq_off, nlen1, d_len, W, d_info, kseq);
    if (launches) *launches += 10;
    return cudaGetLastError();
}

namespace {
struct JoinLayout {
    size_t work, tile_cnt, tile_off, masks, nul, seq_off, sums0, toff0, cnt, total;
    uint32_t ntiles, stiles;
    JoinLayout(uint32_t name_len, uint32_t R) {
        ntiles = cdivu(name_len ? name_len : 1, TILE);
        stiles = cdivu(R ? R : 1, STILE);
        size_t o = 0;
        work = o; o += al256(sizeof(FqWork));
        tile_cnt = o; o += al256((size_t)ntiles * 4);
        tile_off = o; o += al256((size_t)ntiles * 4);
        masks = o; o += al256((size_t)ntiles * TPB * 4);
        nul = o; o += al256((size_t)R * 4 + 4);
        seq_off = o; o += al256((size_t)R * 4 + 4);
        sums0 = o; o += al256((size_t)stiles * 4);
        toff0 = o; o += al256((size_t)stiles * 4);
        cnt = o; o += 256;
        total = o;
    }
};
__global__ void set_u32(uint32_t *p, uint32_t v) { *p = v; }
}  // namespace

size_t fq_join_scratch_bytes(uint32_t name_len, uint32_t num_records) { return JoinLayout(name_len, num_records).total; }

cudaError_t fq_join_launch(const uint8_t *d_name, uint32_t name_len, const uint8_t *d_seq, const uint8_t *d_qual,
                           const uint32_t *d_len, uint32_t R, int plus_name, uint8_t *d_text, uint32_t text_cap,
                           uint8_t *S, FqInfo *d_info, cudaStream_t st, int *launches) {
    JoinLayout L(name_len, R);
    FqWork *W = (FqWork *)(S + L.work);
    uint32_t *tile_cnt = (uint32_t *)(S + L.tile_cnt), *tile_off = (uint32_t *)(S + L.tile_off);
    uint32_t *nul = (uint32_t *)(S + L.nul), *seq_off = (uint32_t *)(S + L.seq_off);
    uint32_t *sums0 = (uint32_t *)(S + L.sums0), *toff0 = (uint32_t *)(S + L.toff0), *cnt = (uint32_t *)(S + L.cnt);
    cudaError_t e = cudaMemsetAsync(W, 0, sizeof(FqWork), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(d_info, 0, sizeof(FqInfo), st);
    if (e != cudaSuccess) return e;
    uint32_t *masks = (uint32_t *)(S + L.masks);
    count_kernel<false><<<L.ntiles, TPB, 0, st>>>(d_name, name_len, 0, tile_cnt, masks, W);
    scan_small<<<1, 1024, 0, st>>>(tile_cnt, tile_off, &W->total, nullptr, nullptr, nullptr, L.ntiles, nullptr, 0, 1);
    mark_kernel<<<L.ntiles, TPB, 0, st>>>(masks, tile_off, nul, R);
    set_u32<<<1, 1, 0, st>>>(cnt, R);
    scan_reduce<<<L.stiles, 256, 0, st>>>(d_len, nullptr, cnt, R, sums0, nullptr);
    scan_small<<<1, 1024, 0, st>>>(sums0, toff0, nullptr, nullptr, nullptr, nullptr, 0, cnt, R, STILE);
    scan_apply<<<L.stiles, 256, 0, st>>>(d_len, nullptr, cnt, R, toff0, nullptr, seq_off, nullptr);
    gather_kernel_fq<<<cdivu(R ? R : 1, GREC), 256, 0, st>>>(d_name, name_len, nul, d_seq, d_qual, d_len, seq_off, R, plus_name,
                                                          d_text, text_cap, W, d_info);
    join_finalize<<<1, 1, 0, st>>>(R, W, d_info);
    if (launches) *launches += 9;
    return cudaGetLastError();
}

}  // namespace b200
